"""Golden frames from the REFERENCE'S OWN renderer (oracle/_ref/libvnr_marcher_ref.so: core/renderer/method_raymarching.cu,
method_pathtracing.cu, core/macrocell.cu compiled unmodified in place).  Run on a GPU box:
    gpurun -- 'python tools/make_golden_marcher.py'          (writes gpurun_out/marcher_ref_golden.npz)
then copy the file to tests/golden/.  tests/test_oracle_golden_marcher.py checks the CPU oracle against it without a GPU.
The scene is regenerated from seeds by the test (synthetic.make_volume / make_tfn), so only the frames are stored."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from instantvnr_b200 import synthetic as syn
from oracle import marcher_ref as MR

DIMS = (40, 28, 36); SIZE = (56, 44); SEED = 11; TFN_N = 48
gt = syn.make_volume(DIMS, seed=SEED)
rgb, alpha = syn.make_tfn(TFN_N)
ref = MR.RefMarcher(DIMS, gt)
ref.set_transfer_function(rgb, alpha, (0.0, 1.0))
mcd, vr, mo = ref.get_macrocell()
out = {"dims": np.array(DIMS), "size": np.array(SIZE), "seed": np.array(SEED), "tfn_n": np.array(TFN_N), "mc_dims": np.array(mcd),
       "mc_value_range": vr, "mc_max_opacity": mo}
cases = []
for mode in (5, 8, 11, 6, 9, 12, 14, 15):
    for view, frames, rate, density in ((1, 1, 1.0, 1.0), (6, 2, 2.0, 0.5), (13, 1, 0.7, 1.0)):
        ref.set_sampling(rate, density)
        ref.reset_accumulation()
        for _ in range(frames):
            img, _ = ref.render(mode, SIZE, *syn.default_camera(DIMS, view), neural=False)
        key = f"frame_m{mode}_v{view}_f{frames}"
        out[key] = img.astype(np.float32)
        cases.append((mode, view, frames, rate, density))
        print(key, "alpha max", float(img[..., 3].max()), "rgb mean", float(img[..., :3].mean()), flush=True)
out["cases"] = np.array(cases, dtype=np.float64)
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/marcher_ref_golden.npz", **out)
print("wrote gpurun_out/marcher_ref_golden.npz", os.path.getsize("gpurun_out/marcher_ref_golden.npz"), "bytes")
