"""Exploration: the reference's own marcher (oracle/_ref/libvnr_marcher_ref.so) against this library and the CPU oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import instantvnr_b200 as vnr
import oracle as O
from oracle import marcher_ref as MR
from instantvnr_b200 import synthetic as syn

CFG = dict(n_levels=4, n_features=8, log2_hashmap=12, base_res=8, n_hidden=2)
DIMS = (48, 32, 40); SIZE = (72, 56)
gt = syn.make_volume(DIMS, seed=5)
vol = vnr.NeuralVolume(vnr.model_json(**CFG), DIMS)
vol.set_groundtruth(gt); vol.init_params(3)
rgb, alpha = syn.make_tfn(64)
vol.set_transfer_function(rgb, alpha)
vol.macrocell_from_groundtruth()
vol.train(200, batch=8192, fast_mode=True)
_, vr_ours, mo_ours = vol.get_macrocell()
ref = MR.RefMarcher(DIMS, gt)
ref.set_transfer_function(rgb, alpha, (0.0, 1.0))
mcd, vr_ref, mo_ref = ref.get_macrocell()
print("macrocell dims", mcd, "value range max|d|", np.abs(vr_ref.reshape(-1) - np.asarray(vr_ours).reshape(-1)).max(), "max opacity max|d|", np.abs(mo_ref - mo_ours).max())
colors = np.concatenate([rgb, np.ones((rgb.shape[0], 1), np.float32)], 1)

def ours(mode, gt_source, view, density=1.0):
    ren = vnr.Renderer(vol); ren.set_size(*SIZE); ren.set_camera(*syn.default_camera(DIMS, view)); ren.set_mode(mode)
    ren.set_groundtruth_source(gt_source); ren.set_density_scale(density); ren.render()
    return ren.map_frame().copy(), ren.stats()

def cmp(name, a, b):
    d = np.abs(a - b)
    print(f"{name}: max|d| {d.max():.3e} mean|d| {d.mean():.3e} psnr {syn.psnr(a, b):.1f} dB  frac>1e-4 {(d.max(-1) > 1e-4).mean():.4f}", flush=True)

for view in (2, 9):
    cam = syn.default_camera(DIMS, view)
    for mode in (5, 8, 11, 4, 7, 10, 6):
        ref.reset_accumulation()
        r, _ = ref.render(mode, SIZE, *cam, neural=False)
        o, st = ours(mode, True, view)
        cmp(f"GT source view {view} mode {mode}: reference vs ours", r, o)
    fr = O.Frame(DIMS, *SIZE, *cam)
    w, _, _ = O.render(None, None, fr, mo_ours, colors, alpha, volume=gt) if False else (None, None, None)
# neural: reference marcher + OUR decoder
ref.set_decoder(MR.function_address(vnr.lib(), "vnr_volume_decode"), vol._h)
ref.set_macrocell_value_range(np.asarray(vr_ours, np.float32))
for view in (2, 9):
    cam = syn.default_camera(DIMS, view)
    for mode in (5, 8, 11):
        ref.reset_accumulation()
        r, st = ref.render(mode, SIZE, *cam, neural=True)
        o, ost = ours(mode, False, view)
        cmp(f"network view {view} mode {mode}: reference marcher + our decode vs ours ({st['decode_coords']} vs {ost['samples_decoded']} coords)", r, o)
# path tracing on the GT volume
for mode in (14, 15, 13):
    ref.reset_accumulation()
    r, _ = ref.render(mode, SIZE, *syn.default_camera(DIMS, 3), neural=False)
    o, _ = ours(mode if mode != 13 else 15, True, 3)
    d = np.abs(r[..., :3] - o[..., :3]).max(-1)
    print(f"path tracing mode {mode} GT source: {100 * (d <= 1e-4).mean():.2f} % of pixels within 1e-4, means {r[..., :3].mean():.5f} vs {o[..., :3].mean():.5f}", flush=True)
ref.reset_accumulation()
r, st = ref.render(14, SIZE, *syn.default_camera(DIMS, 3), neural=True)
o, ost = ours(14, False, 3)
d = np.abs(r[..., :3] - o[..., :3]).max(-1)
print(f"path tracing mode 14 network (our decode): {100 * (d <= 1e-3).mean():.2f} % within 1e-3, means {r[..., :3].mean():.5f} vs {o[..., :3].mean():.5f}; {st} vs {ost}", flush=True)
