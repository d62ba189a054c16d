"""GPU parity: sample-streaming marcher (macrocell DDA, compaction, compositing, ERT) vs the oracle."""
import numpy as np
import pytest

import instantvnr_b200 as vnr
import oracle as O
from instantvnr_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

# frame tolerance: PSNR >= 50 dB and max-abs <= 4/255 against the oracle frame
PSNR_MIN, MAXABS = 50.0, 4.0 / 255.0


def _scene(dims, cfg, seed=7, grid_scale=3000.0):
    m = O.ModelCfg(cfg.get("n_levels", 8), cfg.get("n_features", 8), cfg.get("log2_hashmap", 19), cfg.get("base_res", 16), 2.0, cfg.get("n_hidden", 4))
    p32, _ = O.init_params(m, seed)
    p32 = p32.copy(); p32[m.n_mlp:] *= grid_scale
    p16 = O.f32_to_f16(p32)
    # a "decoded volume" on the voxel grid gives consistent macrocell value ranges
    zz, yy, xx = np.meshgrid(*[(np.arange(d, dtype=np.float32) + 0.5) / d for d in dims[::-1]], indexing="ij")
    grid_xyz = np.stack([xx.ravel(), yy.ravel(), zz.ravel()], axis=1)
    dec = O.decode(m, p16, grid_xyz)
    lo, hi = float(dec.min()), float(dec.max())
    return m, p16, dec, (lo, hi)


def _render_both(dims, cfg, size, view=1, jitter_mode=0, sampling_rate=1.0, frames=1, partition=None, n_iters=16):
    m, p16, dec, (lo, hi) = _scene(dims, cfg)
    rgb, alpha = syn.make_tfn(64)
    # map the tfn range onto the decoded value range so that part of the volume is transparent
    tr = (max(lo, 0.0), min(hi, 1.0)) if hi > lo else (0.0, 1.0)
    mc = O.macrocell_update_implicit(np.clip(dec, 0, 1), dims)
    mo = O.macrocell_max_opacity(mc, alpha, tr[0], tr[1])
    cam = syn.default_camera(dims, view)
    w, h = size
    vol = vnr.NeuralVolume(vnr.model_json(**cfg), dims)
    vol.set_params_f16(p16)
    vol.set_transfer_function(rgb, alpha, tr)
    vol.set_macrocell(mc)
    md, vr, gmo = vol.get_macrocell()
    assert md == O.macrocell_dims(dims)
    assert np.array_equal(gmo, mo)                     # max-opacity kernel: exact
    ren = vnr.Renderer(vol)
    ren.set_size(w, h)
    ren.set_camera(*cam)
    ren.set_mode(vnr.VNR_RAYMARCHING_NO_SHADING_SAMPLE_STREAMING)
    ren.set_sampling_rate(sampling_rate)
    ren.set_jitter_mode(jitter_mode)
    if partition:
        ren.set_partition(*partition)
    colors_rgba = np.concatenate([rgb, np.ones((rgb.shape[0], 1), np.float32)], axis=1)
    accum = None
    got = want = None
    for f in range(1, frames + 1):
        ren.render()
        got = ren.map_frame()
        fr = O.Frame(dims, w, h, *cam, sampling_rate=sampling_rate, tfn_range=tr, frame_index=f, n_iters=n_iters)
        want, accum, ostats = O.render(m, p16, fr, mo, colors_rgba, alpha, acc_mode=0, jitter_mode=jitter_mode, accum=accum)
    return got, want, ren.stats(), ostats


@pytest.mark.parametrize("cfg,dims,size", [
    (dict(), (64, 64, 64), (96, 64)),
    (dict(n_levels=16, n_features=2, n_hidden=2), (48, 64, 32), (64, 64)),
])
def test_frame_matches_oracle(cfg, dims, size):
    got, want, gstats, ostats = _render_both(dims, cfg, size)
    assert want[..., 3].max() > 0.3, "scene is not vacuous"
    assert gstats["rays_hit"] == ostats["rays_hit"]
    assert abs(gstats["samples_decoded"] - ostats["samples_decoded"]) <= 0.001 * ostats["samples_decoded"] + 16
    assert syn.psnr(got, want) >= PSNR_MIN
    assert np.abs(got - want).max() <= MAXABS


def test_frame_fixed_jitter_and_high_sampling_rate():
    got, want, gstats, ostats = _render_both((64, 64, 64), dict(), (64, 48), view=3, jitter_mode=1, sampling_rate=2.5)
    assert gstats["rays_hit"] == ostats["rays_hit"]
    assert syn.psnr(got, want) >= PSNR_MIN
    assert np.abs(got - want).max() <= MAXABS


def test_accumulation_over_frames():
    """frame_index > 1: accum += rgba; frame = accum / frame_index (writePixelColor)."""
    got, want, _, _ = _render_both((32, 32, 32), dict(log2_hashmap=14), (48, 32), frames=3)
    assert syn.psnr(got, want) >= PSNR_MIN
    assert np.abs(got - want).max() <= MAXABS


def test_camera_outside_misses_everything():
    dims = (32, 32, 32)
    vol = vnr.NeuralVolume(vnr.model_json(log2_hashmap=14), dims)
    vol.init_params(3)
    rgb, alpha = syn.make_tfn(32)
    vol.set_transfer_function(rgb, alpha)
    ren = vnr.Renderer(vol)
    ren.set_size(32, 32)
    ren.set_camera([0, 0, -100], [0, 0, -200], [0, 1, 0])     # looking away from the volume
    ren.render()
    img = ren.map_frame()
    assert np.all(img == 0)
    assert ren.stats()["rays_hit"] == 0


def test_partition_tiles_reassemble():
    """Tile-parallel rendering: the union of the ranks' pixel strips equals the single-GPU frame."""
    dims, cfg, size = (64, 64, 64), dict(log2_hashmap=15), (64, 50)   # 50 rows: last strip is partial
    full, _, _, _ = _render_both(dims, cfg, size)
    world = 3
    acc = np.zeros_like(full)
    for rank in range(world):
        part, _, _, _ = _render_both(dims, cfg, size, partition=(rank, world))
        rows = [y for y in range(size[1]) if (y // 4) % world == rank]
        other = [y for y in range(size[1]) if (y // 4) % world != rank]
        assert np.all(part[other] == 0)
        acc[rows] = part[rows]
    assert np.array_equal(acc, full)


def test_graph_loop_equals_host_enqueued_rounds():
    """The CUDA-graph WHILE loop and the bounded host-enqueued rounds run the same kernels: identical frames."""
    dims, cfg = (64, 64, 64), dict(log2_hashmap=15)
    m, p16, dec, (lo, hi) = _scene(dims, cfg)
    rgb, alpha = syn.make_tfn(64)
    tr = (max(lo, 0.0), min(hi, 1.0))
    vol = vnr.NeuralVolume(vnr.model_json(**cfg), dims)
    vol.set_params_f16(p16)
    vol.set_transfer_function(rgb, alpha, tr)
    vol.set_macrocell(O.macrocell_update_implicit(np.clip(dec, 0, 1), dims))
    out = []
    for graph in (True, False, True):
        ren = vnr.Renderer(vol)
        ren.set_size(80, 60)
        ren.set_graph(graph)
        frames = []
        for view in (1, 5):                       # second frame replays the captured graph with a new camera
            ren.set_camera(*syn.default_camera(dims, view))
            ren.render()
            frames.append(ren.map_frame())
        ren.set_size(64, 64)                      # resize: the graph is rebuilt
        ren.render()
        frames.append(ren.map_frame())
        out.append((frames, ren.stats()))
    for k in range(3):
        assert np.array_equal(out[0][0][k], out[1][0][k])
        assert np.array_equal(out[0][0][k], out[2][0][k])
    assert out[0][1] == out[1][1]
    assert out[0][0][0][..., 3].max() > 0.3


def test_full_size_frame_is_invariant_under_rescheduling():
    """BASELINE configs[1] size (256^3 volume, 1024^2 frame, example model): the frame is a per-ray function of the
    samples, so it must not depend on how the wavefront is scheduled: device-driven graph loop vs host-enqueued
    rounds and one renderer vs the union of three pixel partitions give the SAME bits; a different number of samples
    per ray per round (n_iters) resumes the DDA at round boundaries through next_cell_begin = t - t_min
    (dda.h:20-138), which re-rounds t by an ulp, so those frames agree to rounding only."""
    dims = (256, 256, 256)
    m = O.ModelCfg()
    p32, _ = O.init_params(m, 7)
    p32 = p32.copy(); p32[m.n_mlp:] *= 3000.0
    vol = vnr.NeuralVolume(vnr.example_model_json(), dims)
    vol.set_params_f16(O.f32_to_f16(p32))
    dec = np.clip(vol.decode_host(np.random.default_rng(1).random((100000, 3), dtype=np.float32)), 0.0, 1.0)
    assert dec.max() > dec.min()
    rgb, alpha = syn.make_tfn(256)
    vol.set_transfer_function(rgb, alpha, (float(dec.min()), float(dec.max())))
    vol.set_macrocell(np.tile(np.array([-1.0, 2.0], np.float32), 16 ** 3))        # value range [0,1] everywhere

    def frame(n_iters=16, graph=True, partition=None):
        ren = vnr.Renderer(vol)
        ren.set_size(1024, 1024)
        ren.set_camera(*syn.default_camera(dims, 3))
        ren.set_n_iters(n_iters); ren.set_graph(graph)
        if partition:
            ren.set_partition(*partition)
        ren.render()
        return ren.map_frame(), ren.stats()

    ref, st = frame()
    assert ref[..., 3].max() > 0.9 and st["rays_hit"] > 500000 and st["samples_composited"] > 5e6
    assert np.all(ref[..., 3] <= 1.0) and np.all(ref >= 0.0)
    img, st2 = frame(16, False)
    assert np.array_equal(img, ref) and st2 == st
    for n_iters, graph in ((8, True), (5, False)):
        img, st2 = frame(n_iters, graph)
        assert np.abs(img - ref).max() <= 2e-3 and syn.psnr(img, ref) >= 70.0
        assert abs(st2["samples_composited"] - st["samples_composited"]) <= 1e-4 * st["samples_composited"] and st2["rays_hit"] == st["rays_hit"]
    acc = np.zeros_like(ref)
    for rank in range(3):
        part, _ = frame(partition=(rank, 3))
        rows = [y for y in range(1024) if (y // 4) % 3 == rank]
        acc[rows] = part[rows]
    assert np.array_equal(acc, ref)


@pytest.mark.parametrize("size", [(96, 64), (75, 49), (80, 50)])
def test_ray_order_and_slot_layout_do_not_change_the_frame(size):
    """Warps own 8 x 4 pixel tiles and lay their sample slots out depth-major (march_round_kernel); sizes that are not multiples
    of 8 x 4 fall back to scanline order.  Per-ray arithmetic is independent of both: all four combinations, the graph loop,
    the host-enqueued rounds and a 3-way partition (partial last strip) give the same bits, for plain and shaded modes."""
    dims, cfg = (64, 64, 64), dict(log2_hashmap=15)
    m, p16, dec, (lo, hi) = _scene(dims, cfg)
    rgb, alpha = syn.make_tfn(64)
    vol = vnr.NeuralVolume(vnr.model_json(**cfg), dims)
    vol.set_params_f16(p16)
    vol.set_transfer_function(rgb, alpha, (max(lo, 0.0), min(hi, 1.0)))
    vol.set_macrocell(O.macrocell_update_implicit(np.clip(dec, 0, 1), dims))

    def frame(mode, tiled, transpose, graph=True, partition=None):
        ren = vnr.Renderer(vol)
        ren.set_size(*size); ren.set_camera(*syn.default_camera(dims, 3)); ren.set_mode(mode)
        ren.set_layout(tiled, transpose); ren.set_graph(graph)
        if partition:
            ren.set_partition(*partition)
        ren.render()
        return ren.map_frame().copy(), ren.stats()

    for mode in (5, 8, 11):
        ref, st = frame(mode, False, False)
        assert ref[..., 3].max() > 0.3
        for tiled, transpose, graph in ((True, True, True), (True, False, True), (False, True, False), (True, True, False)):
            img, st2 = frame(mode, tiled, transpose, graph)
            assert np.array_equal(img, ref), (mode, tiled, transpose, graph)
            assert st2["samples_decoded"] == st["samples_decoded"] and st2["rays_hit"] == st["rays_hit"]
        acc = np.zeros_like(ref)
        for rank in range(3):
            part, _ = frame(mode, True, True, partition=(rank, 3))
            rows = [y for y in range(size[1]) if (y // 4) % 3 == rank]
            acc[rows] = part[rows]
        assert np.array_equal(acc, ref)


def _ring_renderer(dims=(48, 48, 48), size=(80, 64)):
    cfg = dict(log2_hashmap=14)
    m, p16, dec, (lo, hi) = _scene(dims, cfg)
    rgb, alpha = syn.make_tfn(64)
    tr = (max(lo, 0.0), min(hi, 1.0))
    vol = vnr.NeuralVolume(vnr.model_json(**cfg), dims)
    vol.set_params_f16(p16)
    vol.set_transfer_function(rgb, alpha, tr)
    vol.set_macrocell(O.macrocell_update_implicit(np.clip(dec, 0, 1), dims))
    ren = vnr.Renderer(vol)
    ren.set_size(*size)
    ren.set_mode(vnr.VNR_RAYMARCHING_NO_SHADING_SAMPLE_STREAMING)
    return vol, ren


def test_frames_in_flight_ring_is_fifo_and_bit_identical():
    """vnr_renderer_set_frames_in_flight: consecutive vnr_render calls land in consecutive slots (own stream, buffers, graph);
    vnr_map_frame returns the oldest unmapped frame; every frame equals the one-at-a-time frame bit for bit."""
    dims = (48, 48, 48)
    vol, ren = _ring_renderer(dims)
    cams = [syn.default_camera(dims, v, 7) for v in range(7)]
    single = []
    for cam in cams:
        ren.set_camera(*cam); ren.render(); single.append(ren.map_frame())
    assert len({f.tobytes() for f in single}) == len(single)
    for depth in (2, 3):
        ren.set_frames_in_flight(depth)
        assert len(ren.streams()) == depth and len(set(ren.streams())) == depth
        got = []
        for i, cam in enumerate(cams):               # keep `depth` frames in flight: map frame i - depth + 1 after launching frame i
            ren.set_camera(*cam); ren.render()
            if i >= depth - 1:
                got.append(ren.map_frame())
        while len(got) < len(cams):
            got.append(ren.map_frame())
        for a, b in zip(got, single):
            assert np.array_equal(a, b)
        with pytest.raises(vnr.VnrError):            # nothing left to map
            ren.map_frame()
    # a full ring drops the oldest frame
    ren.set_frames_in_flight(2)
    for cam in cams[:3]:
        ren.set_camera(*cam); ren.render()
    assert np.array_equal(ren.map_frame(), single[1]) and np.array_equal(ren.map_frame(), single[2])


def test_accumulation_across_frame_slots():
    """progressive accumulation (frame_index > 1) reads the previous frame's sums from the previous slot"""
    dims = (48, 48, 48)
    vol, ren = _ring_renderer(dims)
    cam = syn.default_camera(dims, 2)
    ren.set_camera(*cam)
    want = []
    for _ in range(4):
        ren.render(); want.append(ren.map_frame())
    assert not np.array_equal(want[0], want[3])      # jitter differs per frame index
    ren.set_frames_in_flight(3)
    ren.set_camera(*cam)
    for _ in range(4):
        ren.render()
    # the ring holds the last three frames
    got = [ren.map_frame() for _ in range(3)]
    for a, b in zip(got, want[1:]):
        assert np.array_equal(a, b)


def test_second_map_without_render_raises():
    vol, ren = _ring_renderer()
    ren.set_camera(*syn.default_camera((48, 48, 48), 1))
    with pytest.raises(vnr.VnrError):
        ren.map_frame()
    ren.render(); ren.map_frame()
    with pytest.raises(vnr.VnrError):
        ren.map_frame()


def test_training_is_ordered_after_frames_in_flight():
    """vnr_volume_train right after vnr_render (no map in between) must not change the frame: the optimizer waits on the
    device for the frames that still decode the current parameters"""
    dims = (48, 48, 48)
    vol, ren = _ring_renderer(dims, size=(256, 256))
    vol.set_groundtruth(syn.make_volume(dims, seed=3))
    cam = syn.default_camera(dims, 1)
    ren.set_camera(*cam)
    ren.render(); want = ren.map_frame()
    p0 = vol.get_params_f16()
    ren.reset_accumulation()
    ren.render()
    vol.train(3, batch=4096, fast_mode=True)         # rewrites the parameters and (fast_mode, online macrocell) the value ranges
    got = ren.map_frame()
    assert np.array_equal(got, want)
    assert not np.array_equal(vol.get_params_f16(), p0)


def test_zero_copy_host_frame_skips_unchanged_zero_pixels_and_stays_exact():
    """Zero-copy download (the default on one GPU): the compositing kernels store finished pixels straight into the pinned host
    frame and do not store a zero pixel over a zero pixel again (FrameParams::host_nonzero, march.cuh).  Over a camera path that
    turns hits into misses and back -- including views that miss the volume altogether -- and with the download mode switched in
    between, every mapped frame equals the frame of the copy-after-the-frame path bit for bit."""
    dims = (48, 48, 48)
    vol, ren = _ring_renderer(dims)
    away = (np.array([0, 0, -600], np.float32), np.array([0, 0, -1200], np.float32), np.array([0, 1, 0], np.float32))
    near = syn.default_camera(dims, 3, 9)
    close = (near[0] * 0.6, near[1], near[2])                    # closer: covers more of the image
    cams = [syn.default_camera(dims, v, 9) for v in range(9)]
    path = [cams[0], cams[1], away, cams[2], close, cams[3], away, away, close, cams[4], cams[5], cams[0], cams[8]]
    ren.set_zero_copy(False)
    want = []
    for cam in path:
        ren.set_camera(*cam); ren.render(); want.append(ren.map_frame())
    assert not want[2].any() and want[4].any()                   # the path really has empty and covered frames
    bg = [float((f[..., 3] == 0).mean()) for f in want]
    assert min(bg) < 0.6 < max(bg)
    for depth in (1, 2, 3):
        ren.set_frames_in_flight(depth)
        ren.set_zero_copy(True)
        got = []
        for i, cam in enumerate(path):
            if i == 7:
                ren.set_zero_copy(False)                         # one frame arrives by copy: the masks of that host buffer are stale afterwards
            if i == 8:
                ren.set_zero_copy(True)
            ren.set_camera(*cam); ren.render()
            if i >= depth - 1:
                got.append(ren.map_frame())
        while len(got) < len(path):
            got.append(ren.map_frame())
        for k, (a, b) in enumerate(zip(got, want)):
            assert np.array_equal(a, b), (depth, k)
