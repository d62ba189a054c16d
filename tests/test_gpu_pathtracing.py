"""Path-tracing modes 13 / 14 / 15 (pathtrace.cuh) against the oracle's restatement of core/renderer/method_pathtracing.cu.
A path is a chain of accept / reject decisions `u * majorant < alpha(value) * density`: one ulp of difference in a
transcendental (logf / sincosf on the device vs libm) or an fp16-level difference in a decoded value can flip one decision
and send that pixel down another path, so parity is stated per pixel for the bulk and in the mean for the frame:
  * volume sources (bit-exact trilinear lookups): >= 98 % of the pixels within 1e-4 of the oracle, frame mean within 1 %;
  * network source (decode within fp16 tolerance): >= 85 % of the pixels within 1e-3, frame mean within 3 %."""
import numpy as np
import pytest

import instantvnr_b200 as vnr
import oracle as O
from instantvnr_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
CFG = dict(n_levels=4, n_features=8, log2_hashmap=12, base_res=8, n_hidden=2)
DIMS = (48, 32, 40)
SIZE = (72, 56)


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


def _renderer(vol, mode, gt_source=False, view=2, density=1.0, graph=True):
    ren = vnr.Renderer(vol)
    ren.set_size(*SIZE)
    ren.set_camera(*syn.default_camera(DIMS, view))
    ren.set_mode(mode)
    ren.set_groundtruth_source(gt_source)
    ren.set_density_scale(density)
    ren.set_graph(graph)
    return ren


def _agreement(got, want, tol):
    d = np.abs(got[..., :3] - want[..., :3]).max(-1)
    return float((d <= tol).mean()), float(got[..., :3].mean()), float(want[..., :3].mean())


@pytest.mark.parametrize("mode,streaming", [(14, True), (15, False), (13, False)])
def test_volume_sources_match_the_oracle(mode, streaming):
    """Mode 14 on a SimpleVolume streams its samples through the trilinear lookup kernel, mode 15 on a SimpleVolume and mode
    13 (the progressively decoded network) run the single-kernel tracer (render_normal / render_neural, renderer.cpp:143-225)."""
    vol, gt, rgb, alpha = _trained_volume(150)
    _, _, mo = vol.get_macrocell()
    colors = np.concatenate([rgb, np.ones((rgb.shape[0], 1), np.float32)], 1)
    gt_source = mode != 13
    if not gt_source:
        for _ in range(vol.num_blobs()):
            vol.decode_progressive()
    src = gt if gt_source else vol.get_decoded()
    for view, density in ((3, 1.0), (12, 0.35)):
        ren = _renderer(vol, mode, gt_source, view, density)
        ren.render()
        got, st = ren.map_frame(), ren.stats()
        fr = O.Frame(DIMS, *SIZE, *syn.default_camera(DIMS, view))
        want, _, ost = O.render_pathtracing(fr, mo, colors, alpha, volume=src, streaming=streaming, density_scale=density)
        assert np.all(got[..., 3] == 1.0) and want[..., :3].max() > 0.2
        assert st["rays_hit"] == ost["rays_hit"]
        assert abs(st["samples_decoded"] - ost["samples_decoded"]) <= 0.02 * ost["samples_decoded"]
        frac, gm, wm = _agreement(got, want, 1e-4)
        print(f"mode {mode} view {view}: {100 * frac:.2f} % of pixels within 1e-4, mean {gm:.5f} vs {wm:.5f}")
        assert frac >= 0.98 and abs(gm - wm) <= 0.01 * wm


@pytest.mark.parametrize("mode", [14, 15])
def test_network_source_matches_the_oracle(mode):
    """Modes 14 and 15 on a neural volume: every tentative collision is one decode of the fused hash-grid + MLP kernel."""
    vol, gt, rgb, alpha = _trained_volume(200)
    m = O.ModelCfg(CFG["n_levels"], CFG["n_features"], CFG["log2_hashmap"], CFG["base_res"], 2.0, CFG["n_hidden"])
    p16 = vol.get_params_f16()
    _, _, mo = vol.get_macrocell()
    colors = np.concatenate([rgb, np.ones((rgb.shape[0], 1), np.float32)], 1)
    ren = _renderer(vol, mode, view=5)
    ren.render()
    got, st = ren.map_frame(), ren.stats()
    fr = O.Frame(DIMS, *SIZE, *syn.default_camera(DIMS, 5))
    want, _, ost = O.render_pathtracing(fr, mo, colors, alpha, m=m, params_f16=p16, streaming=True)
    assert st["rays_hit"] == ost["rays_hit"] and st["rounds"] > 4
    assert abs(st["samples_decoded"] - ost["samples_decoded"]) <= 0.03 * ost["samples_decoded"]
    frac, gm, wm = _agreement(got, want, 1e-3)
    print(f"mode {mode}: {100 * frac:.2f} % of pixels within 1e-3, mean {gm:.5f} vs {wm:.5f}, {st['rounds']} rounds, {st['samples_decoded']} decodes")
    assert frac >= 0.85 and abs(gm - wm) <= 0.03 * wm


def test_graph_and_host_loops_are_identical_and_frames_accumulate():
    vol, _, _, _ = _trained_volume(100)
    frames = []
    for graph in (True, False):
        ren = _renderer(vol, 14, view=4, graph=graph)
        ren.render(); a = ren.map_frame().copy()
        ren.render(); b = ren.map_frame().copy()             # frame_index 2: new random sequence, running mean
        frames.append((a, b, ren.stats()))
    assert np.array_equal(frames[0][0], frames[1][0]) and np.array_equal(frames[0][1], frames[1][1])
    assert frames[0][2]["samples_decoded"] == frames[1][2]["samples_decoded"] and frames[0][2]["rounds"] == frames[1][2]["rounds"]
    assert not np.array_equal(frames[0][0], frames[0][1])
    # Monte Carlo convergence: two independent 24-frame means are closer to each other than two single frames
    ren = _renderer(vol, 14, view=4)
    def mean_of(n):
        ren.reset_accumulation()
        first = None
        for k in range(n):
            ren.render()
            if k == 0:
                first = ren.map_frame().copy()
        return first, ren.map_frame().copy()
    f1, m24 = mean_of(24)
    ren2 = _renderer(vol, 14, view=4)
    for _ in range(24):                                      # frames 1..24, then 25..48 on the same renderer: a different sequence
        ren2.render()
    acc24 = ren2.map_frame().copy()
    for _ in range(24):
        ren2.render()
    acc48 = ren2.map_frame().copy()
    second24 = 2.0 * acc48 - acc24                           # mean of frames 25..48
    assert np.array_equal(acc24, m24)                        # same frame indices, same result
    noise1 = np.abs(f1[..., :3] - m24[..., :3]).mean()
    noise24 = np.abs(second24[..., :3] - m24[..., :3]).mean()
    assert noise24 < 0.5 * noise1


def test_density_scale_and_tile_partition():
    vol, _, _, _ = _trained_volume(100)
    thin = _renderer(vol, 14, view=7, density=0.1); thin.render(); a = thin.map_frame().copy()
    thick = _renderer(vol, 14, view=7, density=2.0); thick.render(); b = thick.map_frame().copy()
    assert (a[..., :3].sum(-1) == 0).mean() > (b[..., :3].sum(-1) == 0).mean()       # denser medium: fewer unscattered rays
    # interleaved strips of two partitions reproduce the full frame (rays are independent; the generator is keyed by pixel)
    full = _renderer(vol, 14, view=7); full.render(); want = full.map_frame().copy()
    got = np.zeros_like(want)
    for rank in range(2):
        ren = _renderer(vol, 14, view=7)
        ren.set_partition(rank, 2)
        ren.render()
        img = ren.map_frame()
        rows = np.arange(SIZE[1])
        mine = ((rows // 4) % 2) == rank
        got[mine] = img[mine]
    assert np.array_equal(got, want)
