"""The CPU oracle's marcher, path tracer and macrocells against frames produced by the REFERENCE'S OWN renderer
(tests/golden/marcher_ref_golden.npz: core/renderer/method_raymarching.cu, method_pathtracing.cu and core/macrocell.cu compiled
unmodified in place and run on a B200 by tools/make_golden_marcher.py; the scene is regenerated here from its seeds).
Tolerances: macrocells bit-exact; ray-marching frames PSNR >= 70 dB and max-abs <= 4/255 (fma contraction, and the texture
unit's filtering restated with 1.8 fixed-point weights; single-shade modes: <= 0.2 % of the pixels may exceed it, an argmax); path-traced frames >= 97 % of the pixels within 1e-3 and frame means
within 3 % (one ulp in logf / sincosf flips an accept / reject decision now and then and sends that pixel down another path)."""
import os

import numpy as np
import pytest

import oracle as O
from instantvnr_b200 import synthetic as syn

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "marcher_ref_golden.npz"))
DIMS = tuple(int(x) for x in G["dims"]); SIZE = tuple(int(x) for x in G["size"])
CASES = [tuple(c) for c in G["cases"]]
DUMMY_MODEL = O.ModelCfg(2, 2, 4, 2, 2.0, 1)          # the volume source never touches the network


@pytest.fixture(scope="module")
def scene():
    gt = syn.make_volume(DIMS, seed=int(G["seed"]))
    rgb, alpha = syn.make_tfn(int(G["tfn_n"]))
    mc = O.macrocell_update_implicit(gt, DIMS)
    mo = O.macrocell_max_opacity(mc, alpha, 0.0, 1.0)
    return gt, rgb, alpha, mc, mo


def test_macrocells_equal_the_reference(scene):
    gt, rgb, alpha, mc, mo = scene
    assert tuple(int(x) for x in G["mc_dims"]) == O.macrocell_dims(DIMS)
    assert np.array_equal(np.asarray(mc, np.float32).reshape(-1), G["mc_value_range"].reshape(-1))
    assert np.array_equal(np.asarray(mo, np.float32).reshape(-1), G["mc_max_opacity"].reshape(-1))


@pytest.mark.parametrize("case", CASES, ids=[f"mode{int(c[0])}-view{int(c[1])}" for c in CASES])
def test_frames_equal_the_reference(scene, case):
    gt, rgb, alpha, mc, mo = scene
    mode, view, frames, rate, density = int(case[0]), int(case[1]), int(case[2]), float(case[3]), float(case[4])
    want = G[f"frame_m{mode}_v{view}_f{frames}"]
    colors = np.concatenate([rgb, np.ones((rgb.shape[0], 1), np.float32)], 1)
    shade = {5: 0, 6: 0, 8: 1, 9: 1, 11: 2, 12: 2}.get(mode, 0)
    accum = None; light = None
    for f in range(1, frames + 1):
        fr = O.Frame(DIMS, *SIZE, *syn.default_camera(DIMS, view), sampling_rate=rate, frame_index=f, shade_mode=shade, light_dir=light)
        light = fr.light_dir.copy()                      # the sign flip of the light persists in the renderer (renderer.cpp:98-101)
        if mode in (5, 8, 11):
            got, accum, _ = O.render(DUMMY_MODEL, None, fr, mo, colors, alpha, volume=gt, accum=accum)
        elif mode in (6, 9, 12):
            got, accum, _ = O.render_single_kernel(fr, mo, colors, alpha, gt, accum=accum)
        else:
            got, accum, _ = O.render_pathtracing(fr, mo, colors, alpha, volume=gt, streaming=(mode == 14), density_scale=density, accum=accum)
    if mode >= 13:
        d = np.abs(got[..., :3] - want[..., :3]).max(-1)
        assert np.all(got[..., 3] == 1.0)
        assert (d <= 1e-3).mean() >= 0.97, (d <= 1e-3).mean()
        assert abs(got[..., :3].mean() - want[..., :3].mean()) <= 0.03 * want[..., :3].mean()
    else:
        assert want[..., 3].max() > 0.3
        d = np.abs(got - want).max(-1)
        assert syn.psnr(got, want) >= 70.0, syn.psnr(got, want)
        if shade == 2:      # the single shade's point is an argmax over the samples of a ray: a near-tie may resolve differently
            assert (d > 4.0 / 255.0).mean() <= 0.002 and d.max() <= 0.1, ((d > 4.0 / 255.0).mean(), d.max())
        else:
            assert d.max() <= 4.0 / 255.0, d.max()
