"""Rendering mode 4 (march the progressively decoded volume), the ground-truth (SimpleVolume) renderer and the
north-star frame criterion: |PSNR(ours, GT render) - PSNR(reference arithmetic, GT render)| <= 0.1 dB."""
import numpy as np
import pytest

import instantvnr_b200 as vnr
import oracle as O
from instantvnr_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
CFG = dict(n_levels=4, n_features=8, log2_hashmap=12, base_res=8, n_hidden=2)
DIMS = (48, 32, 40)


def _trained_volume(steps=300):
    gt = syn.make_volume(DIMS, seed=5)
    vol = vnr.NeuralVolume(vnr.model_json(**CFG), DIMS)
    vol.set_groundtruth(gt)
    vol.init_params(3)
    rgb, alpha = syn.make_tfn(64)
    vol.set_transfer_function(rgb, alpha)
    vol.macrocell_from_groundtruth()
    vol.train(steps, batch=8192, fast_mode=True)
    return vol, gt, rgb, alpha


def _frame(vol, mode=5, gt_source=False, size=(72, 56), view=2):
    ren = vnr.Renderer(vol)
    ren.set_size(*size)
    ren.set_camera(*syn.default_camera(DIMS, view))
    ren.set_mode(mode)
    ren.set_groundtruth_source(gt_source)
    ren.render()
    return ren.map_frame(), ren.stats()


def test_progressive_decode_fills_the_volume_blob_by_blob():
    vol, gt, _, _ = _trained_volume(50)
    assert vol.num_blobs() == 3                              # 40 slices / 16
    dx, dy, dz = DIMS
    f = np.float32                                           # generate_coords (network.cu:51-68): (i + 0.5f) * (1.f / dims), in float
    ax = [(np.arange(n, dtype=f) + f(0.5)) * (f(1.0) / f(n)) for n in (dz, dy, dx)]
    zz, yy, xx = np.meshgrid(*ax, indexing="ij")
    want = vol.decode_host(np.stack([xx.ravel(), yy.ravel(), zz.ravel()], 1).astype(np.float32)).reshape(dz, dy, dx)
    vol.decode_progressive()
    d = vol.get_decoded()
    assert np.array_equal(d[:16], want[:16]) and not d[16:].any()      # first blob only
    vol.decode_progressive(); vol.decode_progressive()
    assert np.array_equal(vol.get_decoded(), want)
    vol.decode_progressive()                                            # the cursor wrapped: blob 0 again, same values
    assert np.array_equal(vol.get_decoded(), want)


def test_groundtruth_renderer_matches_oracle_and_mode4_marches_the_decoded_volume():
    vol, gt, rgb, alpha = _trained_volume(200)
    m = O.ModelCfg(CFG["n_levels"], CFG["n_features"], CFG["log2_hashmap"], CFG["base_res"], 2.0, CFG["n_hidden"])
    p16 = vol.get_params_f16()
    _, _, mo = vol.get_macrocell()
    colors = np.concatenate([rgb, np.ones((rgb.shape[0], 1), np.float32)], 1)
    w, h = 72, 56
    fr = O.Frame(DIMS, w, h, *syn.default_camera(DIMS, 2))
    # ground-truth renderer (SimpleVolume), sample-streaming mode: same wavefront, trilinear volume lookup
    got_gt, st = _frame(vol, mode=5, gt_source=True)
    want_gt, _, ost = O.render(m, p16, fr, mo, colors, alpha, volume=gt)
    assert want_gt[..., 3].max() > 0.3 and st["rays_hit"] == ost["rays_hit"]
    assert syn.psnr(got_gt, want_gt) >= 50.0 and np.abs(got_gt - want_gt).max() <= 4.0 / 255.0
    # ... and in mode 4 the single-kernel marcher (equal steps per macrocell)
    got_gt4, st4 = _frame(vol, mode=4, gt_source=True)
    want_gt4, _, ost4 = O.render_single_kernel(fr, mo, colors, alpha, gt)
    assert st4["rays_hit"] == ost4["rays_hit"] and abs(st4["samples_decoded"] - ost4["samples_decoded"]) <= 0.002 * ost4["samples_decoded"]
    assert syn.psnr(got_gt4, want_gt4) >= 50.0 and np.abs(got_gt4 - want_gt4).max() <= 4.0 / 255.0
    assert syn.psnr(got_gt4, got_gt) >= 35.0                             # two samplings of the same integral
    # mode 4 before any decode: a zero volume renders nothing the transfer function maps to alpha > 0
    empty, _ = _frame(vol, mode=4)
    assert not empty[..., 3].any()
    for _ in range(vol.num_blobs()):
        vol.decode_progressive()
    got4, _ = _frame(vol, mode=4)
    want4, _, _ = O.render_single_kernel(fr, mo, colors, alpha, vol.get_decoded())
    assert syn.psnr(got4, want4) >= 50.0 and np.abs(got4 - want4).max() <= 4.0 / 255.0
    # and it approximates the per-sample decode of mode 5 (trilinear reconstruction of the decoded voxels)
    got5, _ = _frame(vol, mode=5)
    assert syn.psnr(got4, got5) >= 30.0


@pytest.mark.parametrize("mode,shade", [(7, 1), (10, 2), (9, 1), (12, 2)])
def test_shaded_single_kernel_modes_match_the_oracle(mode, shade):
    """Decoding modes 7 / 10 march the decoded network, in-shader modes 9 / 12 on a SimpleVolume march the ground truth:
    both run raymarching_traceray (method_raymarching.cu:400-487) with the boundary-aware gradient and the inline shadow ray."""
    vol, gt, rgb, alpha = _trained_volume(150)
    _, _, mo = vol.get_macrocell()
    colors = np.concatenate([rgb, np.ones((rgb.shape[0], 1), np.float32)], 1)
    gt_source = mode in (9, 12)
    if not gt_source:
        for _ in range(vol.num_blobs()):
            vol.decode_progressive()
    src = gt if gt_source else vol.get_decoded()
    for view in (3, 12):
        fr = O.Frame(DIMS, 72, 56, *syn.default_camera(DIMS, view), shade_mode=shade)
        got, st = _frame(vol, mode=mode, gt_source=gt_source, view=view)
        want, _, ost = O.render_single_kernel(fr, mo, colors, alpha, src)
        assert want[..., 3].max() > 0.3 and st["rays_hit"] == ost["rays_hit"]
        assert abs(st["samples_decoded"] - ost["samples_decoded"]) <= 0.002 * ost["samples_decoded"]
        assert syn.psnr(got, want) >= 50.0 and np.abs(got - want).max() <= 4.0 / 255.0


def test_neural_frame_psnr_within_a_tenth_of_a_db_of_the_reference_arithmetic():
    """BASELINE north star: 'rendered frames within 0.1 dB PSNR of the reference', measured against the ground-truth
    render: PSNR(our neural frame, our GT frame) vs PSNR(oracle neural frame, oracle GT frame)."""
    vol, gt, rgb, alpha = _trained_volume(400)
    m = O.ModelCfg(CFG["n_levels"], CFG["n_features"], CFG["log2_hashmap"], CFG["base_res"], 2.0, CFG["n_hidden"])
    p16 = vol.get_params_f16()
    _, _, mo = vol.get_macrocell()
    colors = np.concatenate([rgb, np.ones((rgb.shape[0], 1), np.float32)], 1)
    assert vol.psnr() > 30.0                                 # the reference's quality floor (README: PSNR > 30 dB)
    deltas = []
    for view in (1, 6, 11):
        fr = O.Frame(DIMS, 72, 56, *syn.default_camera(DIMS, view))
        ours_n, _ = _frame(vol, 5, view=view); ours_gt, _ = _frame(vol, 5, gt_source=True, view=view)
        ref_n, _, _ = O.render(m, p16, fr, mo, colors, alpha, acc_mode=1)     # acc_mode 1: fp16-accumulating MLP as the reference
        ref_gt, _, _ = O.render(m, p16, fr, mo, colors, alpha, volume=gt)
        a, b = syn.psnr(ours_n, ours_gt), syn.psnr(ref_n, ref_gt)
        assert b > 25.0
        deltas.append(abs(a - b))
    print("PSNR deltas (dB):", deltas)
    assert max(deltas) <= 0.1


@pytest.mark.parametrize("mode,shade", [(8, 1), (11, 2)])
def test_shaded_sample_streaming_modes_match_the_oracle(mode, shade):
    """Modes 8 (gradient shading: 4 decodes per sample, Phong-style scivis light, method_raymarching.cu:719-726,773-788)
    and 11 (single-shade heuristic: camera pass + shadow pass, :789-795,813-833,877-900) on the same wavefront, for the
    network and for the ground-truth volume.  Tolerance: PSNR >= 50 dB and max-abs <= 4/255 vs the oracle frame."""
    vol, gt, rgb, alpha = _trained_volume(200)
    m = O.ModelCfg(CFG["n_levels"], CFG["n_features"], CFG["log2_hashmap"], CFG["base_res"], 2.0, CFG["n_hidden"])
    p16 = vol.get_params_f16()
    _, _, mo = vol.get_macrocell()
    colors = np.concatenate([rgb, np.ones((rgb.shape[0], 1), np.float32)], 1)
    plain, _ = _frame(vol, mode=5)
    for view in (2, 9):
        fr = O.Frame(DIMS, 72, 56, *syn.default_camera(DIMS, view), shade_mode=shade)
        got, st = _frame(vol, mode=mode, view=view)
        want, _, ost = O.render(m, p16, fr, mo, colors, alpha)
        assert want[..., 3].max() > 0.3
        assert st["rays_hit"] == ost["rays_hit"]
        assert abs(st["samples_decoded"] - ost["samples_decoded"]) <= 0.002 * ost["samples_decoded"]
        assert syn.psnr(got, want) >= 50.0 and np.abs(got - want).max() <= 4.0 / 255.0
        # ground-truth source through the same shaded wavefront (render_normal, renderer.cpp:143-180)
        got_gt, _ = _frame(vol, mode=mode, gt_source=True, view=view)
        want_gt, _, _ = O.render(m, p16, fr, mo, colors, alpha, volume=gt)
        assert syn.psnr(got_gt, want_gt) >= 50.0 and np.abs(got_gt - want_gt).max() <= 4.0 / 255.0
    # shading changes colours, never the alpha channel
    got, _ = _frame(vol, mode=mode, view=2)
    assert np.array_equal(got[..., 3], plain[..., 3]) and not np.allclose(got[..., :3], plain[..., :3], atol=1e-3)


def test_shaded_modes_graph_and_host_enqueued_frames_are_identical_and_accumulate():
    vol, _, _, _ = _trained_volume(100)
    for mode in (8, 11):
        frames = []
        for graph in (True, False):
            ren = vnr.Renderer(vol)
            ren.set_size(64, 48); ren.set_camera(*syn.default_camera(DIMS, 4)); ren.set_mode(mode); ren.set_graph(graph)
            ren.render(); a = ren.map_frame().copy()
            ren.render(); b = ren.map_frame().copy()         # frame_index 2: accumulation with a new jitter
            frames.append((a, b))
        assert np.array_equal(frames[0][0], frames[1][0]) and np.array_equal(frames[0][1], frames[1][1])
        assert not np.array_equal(frames[0][0], frames[0][1])
