"""CPU tests of the oracle: known-answer vectors, golden fixtures generated from the reference's own
tiny-cuda-nn build (tools/make_golden_tcnn.py), and internal consistency.  No GPU needed."""
import json
import os

import numpy as np
import pytest

import oracle as O
from instantvnr_b200 import synthetic as syn

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_pcg32_known_answer_vector():
    # pcg32 demo (pcg-c-basic): pcg32_srandom(42, 54) -> first six outputs
    want = [0xA15C02B7, 0x7B47F409, 0xBA1D3330, 0x83D2F293, 0xBFA4784B, 0xCBED606E]
    assert [int(x) for x in O.pcg32_uints(42, 54, 6)] == want


def test_pcg32_advance_equals_stepping():
    a = O.pcg32_uints(1337, 1, 64)
    for k in (1, 4, 17, 63):
        assert np.array_equal(O.pcg32_uints(1337, 1, 64 - k, advance=k), a[k:])
    f = O.pcg32_floats(1337, 1, 1000)
    assert f.min() >= 0.0 and f.max() < 1.0


def test_device_order_uniform_is_a_permutation_of_the_stream():
    # generate_random_uniform: thread i takes stream elements 4i..4i+3 and writes i + n_threads*j
    n = 1000
    r = O.Rng(1337)
    out = r.uniform(n)
    stream = O.pcg32_floats(1337, 1, 4 * 256)
    n_threads = 256
    for i in (0, 1, 5, 249):
        for j in range(4):
            idx = i + n_threads * j
            if idx < n:
                assert out[idx] == stream[4 * i + j]
    # the host generator advanced by exactly n
    nxt = r.uniform(4)
    assert nxt[0] == O.pcg32_floats(1337, 1, 1, advance=n)[0]


def test_level_offset_table_example_model():
    m = O.ModelCfg()
    assert list(m.offsets) == [0, 4096, 36864, 299008, 823296, 1347584, 1871872, 2396160, 2920448]
    assert list(m.res) == [16, 32, 64, 128, 256, 512, 1024, 2048]
    assert m.n_grid == 23363584 and m.n_mlp == 17408 and m.enc_pad == 64
    mb = O.ModelCfg(16, 2, 19, 16, 2.0, 2)
    assert mb.offsets[-1] == 7114752 and mb.n_grid == 14229504 and mb.n_mlp == 7168 and mb.enc_pad == 32
    m22 = O.ModelCfg(8, 8, 22)
    assert m22.offsets[-1] == 19173376


def test_grid_index_dense_and_hash():
    # dense: x + y*res + z*res^2 ; hashed: x ^ y*2654435761 ^ z*805459861 (mod size)
    assert O.grid_index(4096, 16, 3, 5, 7) == 3 + 5 * 16 + 7 * 256
    assert O.grid_index(262144, 64, 63, 63, 63) == 63 + 63 * 64 + 63 * 4096
    want = (9 ^ ((11 * 2654435761) & 0xFFFFFFFF) ^ ((13 * 805459861) & 0xFFFFFFFF)) % 524288
    assert O.grid_index(524288, 128, 9, 11, 13) == want
    # corner + 1 at the upper border wraps through the modulo (faithful to the reference)
    assert O.grid_index(4096, 16, 16, 15, 15) == (16 + 15 * 16 + 15 * 256) % 4096


def test_f16_conversions_exhaustive():
    assert O.lib().orc_f16_selftest() == 0
    x = (np.random.default_rng(0).standard_normal(200000) * 10.0 ** np.random.default_rng(1).integers(-8, 5, 200000)).astype(np.float32)
    assert np.array_equal(O.f32_to_f16(x), x.astype(np.float16).view(np.uint16))
    h = np.arange(65536, dtype=np.uint16)
    f = O.f16_to_f32(h)
    ref = h.view(np.float16).astype(np.float32)
    ok = np.isnan(ref) | (f == ref)
    assert ok.all()


def test_hadd_is_correctly_rounded():
    rng = np.random.default_rng(3)
    a = rng.integers(0, 0x7C00, 20000).astype(np.uint16) | (rng.integers(0, 2, 20000).astype(np.uint16) << 15)
    b = rng.integers(0, 0x7C00, 20000).astype(np.uint16) | (rng.integers(0, 2, 20000).astype(np.uint16) << 15)
    for x, y in zip(a[:5000], b[:5000]):
        assert O.lib().orc_hadd(int(x), int(y)) == O.lib().orc_hadd_exact(int(x), int(y))


def test_init_params_ranges_and_determinism():
    m = O.ModelCfg(4, 4, 12, 8, 2.0, 2)
    p32, p16 = O.init_params(m, 1337)
    q32, _ = O.init_params(m, 1337)
    r32, _ = O.init_params(m, 1338)
    assert np.array_equal(p32, q32) and not np.array_equal(p32, r32)
    g = p32[m.n_mlp:]
    assert g.min() >= -1e-4 and g.max() <= 1e-4 and abs(g.mean()) < 2e-6      # grid U(-1e-4, 1e-4)  grid.h:807
    w0 = p32[:64 * m.enc_pad]
    lim = np.sqrt(6.0 / (64 + m.enc_pad))
    assert np.abs(w0).max() <= lim and np.abs(w0).max() > 0.9 * lim           # Xavier uniform gpu_matrix.h:203
    assert np.array_equal(p16, O.f32_to_f16(p32))


@pytest.mark.parametrize("name", ["example", "variant", "small"])
def test_decode_against_reference_tcnn_golden(name):
    """Golden vectors produced by the reference's own tcnn build on a B200 (NetworkWithInputEncoding::inference)."""
    g = np.load(os.path.join(GOLD, f"tcnn_ref_{name}.npz"))
    assert bool(g["init_identical"])       # tcnn Trainer::initialize_params == oracle.init_params on the same seed
    cfg = json.loads(str(g["cfg"]))
    m = O.ModelCfg(cfg["n_levels"], cfg["n_features"], cfg["log2_hashmap"], cfg["base_res"], 2.0, cfg["n_hidden"])
    p32, _ = O.init_params(m, int(g["seed"]))
    p32 = p32.copy(); p32[m.n_mlp:] *= float(g["grid_scale"])
    p16 = O.f32_to_f16(p32)
    ref = g["decode"]
    d1 = O.decode(m, p16, g["xyz"], acc_mode=1)      # fp16-accumulating emulation of the wmma path
    d0 = O.decode(m, p16, g["xyz"], acc_mode=0)      # fp32 accumulation
    # tensor-core internal summation order is not reproducible on a CPU: 2 fp16 ulps at |y| <= 0.125
    assert np.abs(d1 - ref).max() <= 2 * 2.0 ** -14
    assert (d1 == ref).mean() > 0.97
    assert np.abs(d0 - ref).max() <= 2.0 ** -9


@pytest.mark.parametrize("name", ["small", "variant"])
def test_training_against_reference_tcnn_golden(name):
    """Loss trajectory of Trainer::training_step (reference tcnn on B200) vs the oracle's training step."""
    g = np.load(os.path.join(GOLD, f"tcnn_ref_{name}.npz"))
    cfg = json.loads(str(g["cfg"]))
    m = O.ModelCfg(cfg["n_levels"], cfg["n_features"], cfg["log2_hashmap"], cfg["base_res"], 2.0, cfg["n_hidden"])
    dims = tuple(int(x) for x in g["train_dims"])
    vol = syn.make_volume(dims, seed=int(g["train_vol_seed"]))
    p32, _ = O.init_params(m, int(g["seed"]))
    tr = O.Trainer(m, p32)
    srng = O.Rng(1337)
    losses = []
    for _ in range(len(g["train_losses"])):
        c, t = O.sample_batch(srng, int(g["train_batch"]), vol, dims)
        losses.append(tr.step(c, t, acc_mode=1, grad_mode=1))
    losses = np.array(losses)
    # atomics / split-K order in the reference and Adam's sign-like first steps make parameters diverge
    # at the noise level; the loss curve agrees to well under the 2 % stated in SURVEY 8d
    assert np.abs(losses - g["train_losses"]).max() <= 0.02 * g["train_losses"].max()
    assert losses[-1] < 0.75 * losses[0]


def test_mlp_accumulation_modes_agree_to_fp16():
    m = O.ModelCfg(4, 4, 12, 8, 2.0, 3)
    p32, _ = O.init_params(m, 5)
    p32 = p32.copy(); p32[m.n_mlp:] *= 3000
    p16 = O.f32_to_f16(p32)
    x = np.random.default_rng(0).random((2000, 3), dtype=np.float32)
    a, b = O.decode(m, p16, x, 0), O.decode(m, p16, x, 1)
    assert np.abs(a).max() > 1e-2
    assert np.abs(a - b).max() <= 2.0 ** -9 * max(1.0, np.abs(a).max())
    enc = O.encode(m, p16, x)
    assert np.array_equal(O.mlp(m, p16, enc, 0), a)


def test_tex3d_voxel_centres_and_linearity():
    dims = (8, 6, 5)
    vol = np.random.default_rng(0).random(dims[::-1]).astype(np.float32)
    zz, yy, xx = np.meshgrid(*[(np.arange(d) + 0.5) / d for d in dims[::-1]], indexing="ij")
    c = np.stack([xx.ravel(), yy.ravel(), zz.ravel()], 1).astype(np.float32)
    assert np.allclose(O.tex3d(vol, dims, c), vol.ravel(), atol=0)
    # a linear ramp along x is reproduced up to the 8-bit weight quantisation
    ramp = np.broadcast_to(np.arange(dims[0], dtype=np.float32)[None, None, :], dims[::-1]).copy()
    q = np.array([[0.3, 0.5, 0.5], [0.61, 0.2, 0.9]], np.float32)
    want = q[:, 0] * dims[0] - 0.5
    assert np.abs(O.tex3d(ramp, dims, q) - want).max() <= 1.0 / 256 + 1e-6


def test_macrocell_ranges_and_max_opacity():
    dims = (40, 32, 20)
    vol = syn.make_volume(dims, seed=1)
    mc = O.macrocell_update_implicit(vol, dims)
    md = O.macrocell_dims(dims)
    assert md == (3, 2, 2)
    mcr = mc.reshape(md[2], md[1], md[0], 2)
    # cell (0,0,0) covers voxels [0,16]^3 including the +1 apron; stored with -1 / +1 offsets
    blk = vol[:17, :17, :17]
    assert np.isclose(mcr[0, 0, 0, 0] + 1, blk.min()) and np.isclose(mcr[0, 0, 0, 1] - 1, blk.max())
    # explicit update from samples approaches the same bounds (a border sample also feeds the neighbour
    # cell with a value interpolated one voxel further, hence the small slack)
    rng = np.random.default_rng(0)
    c = rng.random((5000, 3), dtype=np.float32)
    v = O.tex3d(vol, dims, c)
    mc2 = np.zeros_like(mc)
    O.macrocell_update_explicit(c, v, dims, mc2)
    touched = mc2[1::2] != 0
    assert touched.all()
    lo2, hi2 = mc2[0::2] + 1, mc2[1::2] - 1
    assert np.all(lo2 <= hi2) and lo2.min() >= vol.min() - 1e-6 and hi2.max() <= vol.max() + 1e-6
    # online ranges approach the offline ones (border samples also feed the neighbour cell with a value
    # interpolated one voxel further, so they are not strict subsets)
    assert np.abs(lo2 - (mc[0::2] + 1)).mean() < 0.02 and np.abs(hi2 - (mc[1::2] - 1)).mean() < 0.06
    _, alpha = syn.make_tfn(64)
    mo = O.macrocell_max_opacity(mc, alpha)
    assert mo.shape == (12,) and mo.min() >= 0 and mo.max() <= alpha.max() + 1e-7
    # a cell whose range lies entirely below the opacity threshold is transparent
    flat = np.zeros(2, np.float32); flat[0] = 0.05 - 1; flat[1] = 0.1 + 1
    assert O.macrocell_max_opacity(flat, alpha)[0] == 0.0


def test_jitter_lcg_range():
    vals = [O.lcg_tea16_first(1, i) for i in range(2000)]
    assert 0 <= min(vals) and max(vals) < 1 and 0.45 < np.mean(vals) < 0.55


def test_marcher_on_ground_truth_constant_volume():
    """Closed form: constant volume v, constant alpha a: pixel alpha = 1-(1-a)^(path length / step)."""
    dims = (32, 32, 32)
    vol = np.full(dims[::-1], 0.5, np.float32)
    n = 16
    rgb = np.tile(np.array([[1.0, 0.5, 0.25]], np.float32), (n, 1))
    alpha = np.full(n, 0.02, np.float32)
    colors = np.concatenate([rgb, np.ones((n, 1), np.float32)], 1)
    mc = O.macrocell_update_implicit(vol, dims)
    mo = O.macrocell_max_opacity(mc, alpha)
    cam = (np.array([0, 0, -100], np.float32), np.zeros(3, np.float32), np.array([0, 1, 0], np.float32))
    fr = O.Frame(dims, 33, 33, *cam, fovy=10.0)
    img, _, st = O.render(O.ModelCfg(2, 2, 8, 4, 2.0, 1), None, fr, mo, colors, alpha, volume=vol, jitter_mode=1)
    centre = img[16, 16]
    want_alpha = 1 - (1 - 0.02) ** 32.0            # straight through 32 voxels, Sum(dt) = 32, step = 1
    assert abs(centre[3] - want_alpha) < 2e-3
    assert np.allclose(centre[:3], want_alpha * np.array([1.0, 0.5, 0.25]), atol=2e-3)
    assert st["rays_hit"] > 0 and st["samples_composited"] == st["samples_decoded"]
    fr2 = O.Frame(dims, 33, 33, *cam, fovy=60.0)
    img2, _, _ = O.render(O.ModelCfg(2, 2, 8, 4, 2.0, 1), None, fr2, mo, colors, alpha, volume=vol, jitter_mode=1)
    assert img2[0, 0, 3] == 0.0 and img2[16, 16, 3] > 0    # wide field of view: corner rays miss the box


def test_marcher_skips_empty_macrocells_and_terminates_early():
    dims = (64, 64, 64)
    vol = syn.make_volume(dims, seed=2)
    rgb, alpha = syn.make_tfn(64)
    colors = np.concatenate([rgb, np.ones((64, 1), np.float32)], 1)
    mc = O.macrocell_update_implicit(vol, dims)
    mo = O.macrocell_max_opacity(mc, alpha)
    cam = syn.default_camera(dims, 2)
    fr = O.Frame(dims, 48, 48, *cam)
    img, _, st = O.render(O.ModelCfg(2, 2, 8, 4, 2.0, 1), None, fr, mo, colors, alpha, volume=vol)
    all_on = np.ones_like(mo)
    img2, _, st2 = O.render(O.ModelCfg(2, 2, 8, 4, 2.0, 1), None, fr, all_on, colors, alpha, volume=vol)
    assert st["samples_decoded"] < st2["samples_decoded"]        # space skipping + adaptive step reduce work
    assert syn.psnr(img, img2) > 30.0                              # ... without changing the picture much
    assert img[..., 3].max() > 0.5


def test_shaded_marcher_closed_forms():
    """Gradient shading and the single-shade heuristic (method_raymarching.cu:773-833) on cases with a known answer."""
    dims = (32, 32, 32)
    n = 16
    rgb = np.tile(np.array([[1.0, 0.5, 0.25]], np.float32), (n, 1))
    alpha = np.full(n, 0.02, np.float32)
    colors = np.concatenate([rgb, np.ones((n, 1), np.float32)], 1)
    cam = (np.array([0, 0, -100], np.float32), np.zeros(3, np.float32), np.array([0, 1, 0], np.float32))
    m = O.ModelCfg(2, 2, 8, 4, 2.0, 1)
    # (1) constant volume: zero gradient -> shade_scivis_light returns 0 -> colour = (1 - 0.95) * unshaded colour
    vol = np.full(dims[::-1], 0.5, np.float32)
    mo = O.macrocell_max_opacity(O.macrocell_update_implicit(vol, dims), alpha)
    plain, _, st0 = O.render(m, None, O.Frame(dims, 33, 33, *cam, fovy=10.0), mo, colors, alpha, volume=vol, jitter_mode=1)
    shaded, _, st1 = O.render(m, None, O.Frame(dims, 33, 33, *cam, fovy=10.0, shade_mode=1), mo, colors, alpha, volume=vol, jitter_mode=1)
    assert np.array_equal(shaded[..., 3], plain[..., 3]) and st1["samples_decoded"] == 4 * st0["samples_decoded"]
    assert np.allclose(shaded[..., :3], 0.05 * plain[..., :3], atol=1e-6)
    # (2) a linear ramp along x lit head-on: the forward difference is exact, so every sample gets the same factor
    #     0.05 + 0.95 * 0.5 * (simple + scivis) with N = -x in world space
    x = ((np.arange(32, dtype=np.float32) + 0.5) / 32)
    ramp = np.broadcast_to(x[None, None, :], dims[::-1]).astype(np.float32).copy()
    mo = O.macrocell_max_opacity(O.macrocell_update_implicit(ramp, dims), alpha)
    fr = O.Frame(dims, 9, 9, *cam, fovy=2.0, shade_mode=1)
    plain, _, _ = O.render(m, None, O.Frame(dims, 9, 9, *cam, fovy=2.0), mo, colors, alpha, volume=ramp, jitter_mode=1)
    shaded, _, _ = O.render(m, None, fr, mo, colors, alpha, volume=ramp, jitter_mode=1)
    L = fr.light_dir / np.linalg.norm(fr.light_dir)
    assert np.dot(L, [0, 0, 1]) < 0                                   # flipped to face the camera (renderer.cpp:98-101)
    N, V = np.array([-1.0, 0, 0]), np.array([0, 0, -1.0])
    cosNL = max(float(N @ L), 0.0)
    H = (L + V) / np.linalg.norm(L + V)
    scivis = 0.6 + (0.9 * cosNL if cosNL > 0 else 0.0)
    spec = 0.4 * max(float(N @ H), 0.0) ** 40 if cosNL > 0 else 0.0
    simple = 0.2 + 0.8 * abs(float(-V @ N))
    c = plain[4, 4, :3]
    want = 0.05 * c + 0.95 * 0.5 * (simple * c + scivis * c + spec * plain[4, 4, 3])
    assert np.allclose(shaded[4, 4, :3], want, rtol=2e-3, atol=1e-5)
    # (3) single-shade heuristic: alpha untouched; colour = 0.05 c + 0.95 * highest_colour * alpha * T with 0 <= T <= 1;
    #     rays that miss the box stay exactly zero
    vol = syn.make_volume(dims, seed=3)
    rgb2, alpha2 = syn.make_tfn(64)
    colors2 = np.concatenate([rgb2, np.ones((64, 1), np.float32)], 1)
    mo = O.macrocell_max_opacity(O.macrocell_update_implicit(vol, dims), alpha2)
    cam2 = syn.default_camera(dims, 3)
    plain, _, _ = O.render(m, None, O.Frame(dims, 40, 40, *cam2), mo, colors2, alpha2, volume=vol)
    ssh, _, st = O.render(m, None, O.Frame(dims, 40, 40, *cam2, shade_mode=2), mo, colors2, alpha2, volume=vol)
    assert np.array_equal(ssh[..., 3], plain[..., 3]) and plain[..., 3].max() > 0.3
    lo = 0.05 * plain[..., :3]
    hi = lo + 0.95 * rgb2.max() * plain[..., 3:4]
    assert (ssh[..., :3] >= lo - 1e-6).all() and (ssh[..., :3] <= hi + 1e-6).all()
    assert (ssh[..., :3] > lo + 1e-4).any()                            # some light gets through
    assert not ssh[plain[..., 3] == 0].any()
    # the light never shines from behind the camera: a second frame with the already-corrected direction keeps it
    fr3 = O.Frame(dims, 8, 8, *cam2, shade_mode=2, light_dir=fr.light_dir)
    assert float(np.dot(fr3.light_dir, fr3.f[3:6])) <= 0


def test_single_kernel_marcher_closed_forms():
    """raymarching_traceray (method_raymarching.cu:400-487): equal steps per macrocell (sample_size_scaler) integrate the
    same opacity as the streaming marcher; constant volumes shade to 5 % under gradient shading; the single shade is bounded."""
    dims = (32, 32, 32)
    n = 16
    rgb = np.tile(np.array([[1.0, 0.5, 0.25]], np.float32), (n, 1))
    alpha = np.full(n, 0.02, np.float32)
    colors = np.concatenate([rgb, np.ones((n, 1), np.float32)], 1)
    cam = (np.array([0, 0, -100], np.float32), np.zeros(3, np.float32), np.array([0, 1, 0], np.float32))
    vol = np.full(dims[::-1], 0.5, np.float32)
    mo = O.macrocell_max_opacity(O.macrocell_update_implicit(vol, dims), alpha)
    img, _, st = O.render_single_kernel(O.Frame(dims, 33, 33, *cam, fovy=10.0), mo, colors, alpha, vol, jitter_mode=1)
    want_alpha = 1 - (1 - 0.02) ** 32.0            # opacity correction makes the result independent of the step count
    assert abs(img[16, 16, 3] - want_alpha) < 2e-3
    assert np.allclose(img[16, 16, :3], want_alpha * np.array([1.0, 0.5, 0.25]), atol=2e-3)
    # equal division: a 16-voxel macrocell crossed head-on with max-opacity 0.02 -> adaptive step 1 + 15 * 0.9^2 = 13.15
    # -> N = int(16 / 13.15 + 1) = 2 steps of 8 per cell (the streaming marcher takes 13.15 + 2.85)
    assert st["samples_decoded"] // st["rays_hit"] == 4
    shaded, _, st1 = O.render_single_kernel(O.Frame(dims, 33, 33, *cam, fovy=10.0, shade_mode=1), mo, colors, alpha, vol, jitter_mode=1)
    assert np.array_equal(shaded[..., 3], img[..., 3]) and st1["samples_decoded"] == 4 * st["samples_decoded"]
    assert np.allclose(shaded[..., :3], 0.05 * img[..., :3], atol=1e-6)
    vol = syn.make_volume(dims, seed=3)
    rgb2, alpha2 = syn.make_tfn(64)
    colors2 = np.concatenate([rgb2, np.ones((64, 1), np.float32)], 1)
    mo = O.macrocell_max_opacity(O.macrocell_update_implicit(vol, dims), alpha2)
    cam2 = syn.default_camera(dims, 3)
    plain, _, _ = O.render_single_kernel(O.Frame(dims, 40, 40, *cam2), mo, colors2, alpha2, vol)
    stream, _, _ = O.render(O.ModelCfg(2, 2, 8, 4, 2.0, 1), None, O.Frame(dims, 40, 40, *cam2), mo, colors2, alpha2, volume=vol)
    assert syn.psnr(plain, stream) > 35.0
    ssh, _, _ = O.render_single_kernel(O.Frame(dims, 40, 40, *cam2, shade_mode=2), mo, colors2, alpha2, vol)
    lo = 0.05 * plain[..., :3]
    assert np.array_equal(ssh[..., 3], plain[..., 3])
    assert (ssh[..., :3] >= lo - 1e-6).all() and (ssh[..., :3] <= lo + 0.95 * rgb2.max() * plain[..., 3:4] + 1e-6).all()
    assert (ssh[..., :3] > lo + 1e-4).any()


def test_outofcore_sampler_restatement_against_numpy():
    """OutOfCoreSampler::sample (neural_sampler.cpp:1065-1120): the drawn point lies in the chosen voxel's cell, the value
    is trilinear_vkl of the normalised file there, and one ghost row / slice per side covers every voxel it touches."""
    dims = (20, 11, 7)
    rng = np.random.default_rng(2)
    raw = rng.integers(0, 255, dims[0] * dims[1] * dims[2]).astype(np.float32)
    block_rows = 4
    nby = -(-dims[1] // block_rows)
    blocks = [(by, bz) for bz in range(dims[2]) for by in range(nby)]
    first = np.array([(bz * dims[1] + by * block_rows) * dims[0] for by, bz in blocks], dtype=np.uint64)
    length = np.array([dims[0] * (min(by * block_rows + block_rows, dims[1]) - by * block_rows) for by, bz in blocks], dtype=np.uint32)
    r = O.Rng(1337)
    s0 = r.state.copy()
    n = 5000
    xyz, v, bad = O.ooc_sample(r.state, n, first, length, block_rows, raw, dims, 10.0, 200.0)
    assert bad == 0
    assert np.array_equal(r.state, O.Rng(1337).state) is False and not np.array_equal(r.state, s0)     # stream advanced (by 5 n)
    u = O.pcg32_floats(1337, 1, 5 * n).reshape(n, 5)
    bidx = np.minimum((u[:, 3] * np.float32(len(blocks))).astype(np.int64), len(blocks) - 1)
    vidx = np.minimum((u[:, 4] * length[bidx].astype(np.float32)).astype(np.int64), length[bidx] - 1)
    lin = first[bidx].astype(np.int64) + vidx
    vox = np.stack([lin % dims[0], (lin // dims[0]) % dims[1], lin // (dims[0] * dims[1])], 1)
    p = u[:, :3] + vox.astype(np.float32)
    assert np.array_equal(xyz, (p * (np.float32(1) / np.array(dims, np.float32))).astype(np.float32))
    norm = np.clip((raw - np.float32(10.0)) * (np.float32(1) / np.float32(190.0)), 0, 1).reshape(dims[2], dims[1], dims[0]).astype(np.float64)
    q = np.clip(p.astype(np.float64), 0.5, np.array(dims) - 0.5) - 0.5
    i0 = np.floor(q).astype(int); w = q - i0; i1 = np.minimum(i0 + 1, np.array(dims) - 1)
    want = np.zeros(n)
    for cx in (0, 1):
        for cy in (0, 1):
            for cz in (0, 1):
                ix = np.where(cx, i1[:, 0], i0[:, 0]); iy = np.where(cy, i1[:, 1], i0[:, 1]); iz = np.where(cz, i1[:, 2], i0[:, 2])
                want += norm[iz, iy, ix] * np.where(cx, w[:, 0], 1 - w[:, 0]) * np.where(cy, w[:, 1], 1 - w[:, 1]) * np.where(cz, w[:, 2], 1 - w[:, 2])
    assert np.abs(v - want).max() < 1e-5
    # a slab without its ghost rows would not do: shrink the declared slab height and the bound check fires
    _, _, bad2 = O.ooc_sample(O.Rng(1337).state, n, first, length, 1, raw, dims, 10.0, 200.0)
    assert bad2 > 0


def test_training_reduces_loss_small_model():
    m = O.ModelCfg(4, 2, 10, 4, 2.0, 2)
    dims = (16, 16, 16)
    vol = syn.make_volume(dims, seed=3)
    p32, _ = O.init_params(m, 1)
    tr = O.Trainer(m, p32)
    rng = O.Rng(1337)
    losses = [tr.step(*O.sample_batch(rng, 2048, vol, dims)) for _ in range(30)]
    assert losses[-1] < 0.6 * losses[0]


def test_ssim_oracle_against_the_float64_closed_form():
    """orc_ssim (compute_ssim core/network.cu:70-125) vs the Wang et al. formula with a 7^3 uniform window and sample
    covariance evaluated in float64 with scipy (what skimage.metrics.structural_similarity computes for win_size=7)."""
    from scipy.ndimage import uniform_filter
    rng = np.random.default_rng(0)
    a = rng.random((20, 18, 16), dtype=np.float32)
    b = np.clip(a + 0.05 * rng.standard_normal(a.shape).astype(np.float32), 0, 1)
    v, m = O.ssim(a, b, return_map=True)
    A, B = a.astype(np.float64), b.astype(np.float64)
    f = lambda x: uniform_filter(x, 7, mode="constant")[3:-3, 3:-3, 3:-3]
    ux, uy, uxx, uyy, uxy = f(A), f(B), f(A * A), f(B * B), f(A * B)
    cn = 343.0 / 342.0
    vx, vy, vxy = cn * (uxx - ux * ux), cn * (uyy - uy * uy), cn * (uxy - ux * uy)
    S = ((2 * ux * uy + 1e-4) * (2 * vxy + 9e-4)) / ((ux * ux + uy * uy + 1e-4) * (vx + vy + 9e-4))
    assert m.shape == S.shape == (14, 12, 10)
    assert np.abs(m - S).max() < 2e-5 and abs(v - S.mean()) < 1e-6
    assert abs(O.ssim(a, a) - 1.0) < 1e-6                      # identical volumes
    assert O.ssim(a, 1.0 - a) < 0.0                            # anti-correlated structure
    assert O.ssim(np.zeros((7, 7, 7), np.float32), np.zeros((7, 7, 7), np.float32)) == 1.0   # a single window


def test_oracle_takes_the_host_cores_back(monkeypatch):
    """torchrun exports OMP_NUM_THREADS=1; the bench legs that run the oracle ask for the cores the process may use"""
    import os
    n = O.use_host_cores(2)
    assert n == 2
    want = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    assert O.use_host_cores() == want
