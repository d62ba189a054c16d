"""Closed-form checks of the oracle's path tracer (oracle/vnr_oracle.cpp, restating core/renderer/method_pathtracing.cu).
The reference ships no golden frames and its RandomTEA generator is un-vendored, so the random sequence is unpinned; what
IS checkable without it are the estimator's expectations: an empty transfer function gives black frames with alpha 1, and
in a homogeneous medium the fraction of primary rays that cross without a collision is exp(-sigma * length), whatever the
majorant (delta tracking is unbiased under loose majorants)."""
import numpy as np
import pytest

import oracle as O
from instantvnr_b200 import synthetic as syn

DIMS = (32, 32, 32)


def _setup(alpha_value, loose=1.0, size=(96, 72), view=3, frame_index=1):
    fr = O.Frame(DIMS, size[0], size[1], *syn.default_camera(DIMS, view), frame_index=frame_index)
    vol = np.full(DIMS[::-1], 0.5, np.float32)
    colors = np.array([[0.9, 0.6, 0.3, 1.0]] * 2, np.float32)
    alphas = np.full(4, alpha_value, np.float32)
    cells = int(np.prod(O.macrocell_dims(DIMS)))
    mo = np.full(cells, alpha_value * loose, np.float32)
    return fr, vol, colors, alphas, mo


def test_empty_transfer_function_gives_black_opaque_pixels_and_no_samples():
    fr, vol, colors, alphas, mo = _setup(0.0)
    for streaming in (True, False):
        img, accum, st = O.render_pathtracing(fr, mo, colors, alphas, volume=vol, streaming=streaming)
        assert st["samples_decoded"] == 0 and st["rays_hit"] > 0
        assert np.all(img[..., :3] == 0.0) and np.all(img[..., 3] == 1.0)      # writePixelColor(vec4f(L, 1)) for every pixel


@pytest.mark.parametrize("streaming", [True, False])
@pytest.mark.parametrize("loose", [1.0, 3.0])
def test_unscattered_fraction_of_a_homogeneous_medium(streaming, loose):
    """P(primary ray crosses without a real collision) = exp(-alpha * density_scale * length in world units).  A ray that
    collides at least once always picks up light afterwards (directional or ambient, both > 0) unless Russian roulette
    ends it after 5+ scatters, so `L == 0` marks the unscattered rays up to that (here < 0.5 %) tail."""
    a, density = 0.02, 1.5
    fr, vol, colors, alphas, mo = _setup(a, loose)
    rays = O.rays(fr)
    hit = rays[:, 7] >= 0
    length = np.where(hit, rays[:, 7] - rays[:, 6], 0.0)
    expect = np.exp(-a * density * length[hit]).mean()
    img, _, st = O.render_pathtracing(fr, mo, colors, alphas, volume=vol, streaming=streaming, density_scale=density)
    L = img[..., :3].reshape(-1, 3)
    assert st["rays_hit"] == int(hit.sum())
    assert np.all(L[~hit] == 0.0)
    got = float((L[hit].sum(1) == 0.0).mean())
    n = int(hit.sum())
    sigma = np.sqrt(expect * (1 - expect) / n)
    assert abs(got - expect) < 4 * sigma + 0.005, (got, expect, sigma)
    # a looser majorant costs more tentative collisions, never fewer
    if loose > 1.0:
        _, _, tight = O.render_pathtracing(fr, mo / loose, colors, alphas, volume=vol, streaming=streaming, density_scale=density)
        assert st["samples_decoded"] > 1.5 * tight["samples_decoded"]


def test_frames_accumulate_and_the_two_variants_agree_in_the_mean():
    a = 0.03
    means = {}
    for streaming in (True, False):
        accum = None
        for k in range(1, 7):
            fr, vol, colors, alphas, mo = _setup(a, 2.0, size=(64, 48), frame_index=k)
            img, accum, _ = O.render_pathtracing(fr, mo, colors, alphas, volume=vol, streaming=streaming, accum=accum)
        assert np.allclose(img[..., 3], 1.0) and np.allclose(img.reshape(-1, 4) * 6, accum, rtol=1e-5, atol=1e-6)
        means[streaming] = img[..., :3].mean()
    # the streaming variant drops the ambient term of paths that leave right after a shadow ray (iterative_take_sample
    # returns without it, method_pathtracing.cu:613-626): it is darker, but by less than the ambient share
    assert 0.5 * means[False] < means[True] <= means[False] * 1.02
