"""The out-of-core training sampler (vnr_volume_set_groundtruth_outofcore): OutOfCoreSampler::sample + RandomBuffer
(core/samplers/neural_sampler.cpp:488-668,1040-1120) with the slab pool resident in HBM.  Parity: every batch is restated
bit for bit by the oracle from the slab table and the sampler's pcg32 stream; size-independent properties: each value is
the trilinear interpolation of the normalised file at its coordinate, and the pool turns over."""
import numpy as np
import pytest
import torch

import instantvnr_b200 as vnr
import oracle as O

pytestmark = pytest.mark.gpu
CFG = vnr.model_json(n_levels=4, n_features=2, log2_hashmap=12, base_res=4, n_hidden=2)


def _write(tmp_path, dims, dtype, offset=0, seed=1):
    rng = np.random.default_rng(seed)
    n = dims[0] * dims[1] * dims[2]
    if np.dtype(dtype).kind == "f":
        raw = (rng.standard_normal(n) * 40 + 7).astype(dtype)
    else:
        info = np.iinfo(dtype)
        raw = rng.integers(max(info.min, -20000), min(info.max, 50000), n, dtype=dtype)
    path = tmp_path / f"vol_{np.dtype(dtype).name}.raw"
    with open(path, "wb") as f:
        f.write(b"\x55" * offset)
        f.write(raw.tobytes())
    return path, raw


def _draw(vol, n):
    xyz = torch.empty(n, 3, device="cuda"); tgt = torch.empty(n, device="cuda")
    vol.sample(xyz, tgt, n)
    torch.cuda.synchronize()
    return xyz.cpu().numpy(), tgt.cpu().numpy()


@pytest.mark.parametrize("dtype,dims,offset", [("float32", (40, 24, 17), 0), ("uint8", (300, 50, 9), 96), ("int16", (64, 64, 64), 0), ("float64", (33, 7, 5), 8)])
def test_batches_equal_the_oracle_restatement_bit_for_bit(tmp_path, dtype, dims, offset):
    path, raw = _write(tmp_path, dims, dtype, offset)
    f32 = raw.astype(np.float32)
    lo, hi = float(f32.min()) + 1.0, float(f32.max()) - 1.0          # a clamping range
    vol = vnr.NeuralVolume(CFG, dims)
    vol.set_groundtruth_outofcore(path, dtype, (lo, hi), offset=offset, num_concurrent_blocks=8, num_blocks=40)
    info = vol.outofcore_info()
    row = dims[0] * np.dtype(dtype).itemsize
    block_rows = min(-(-32768 // row), dims[1])
    gy, gz = min(block_rows + 2, dims[1]), min(3, dims[2])
    assert info["n_slots"] == 40 and info["n_refresh"] == 8
    assert info["slot_bytes"] == -(-(dims[0] * gy * gz * np.dtype(dtype).itemsize) // 512) * 512     # block_size_aligned :560
    rng = O.Rng(1337)
    tables = []
    for step in range(4):
        t = vol.outofcore_info(table=True)
        tables.append(t["first_voxel"].copy())
        # every slot is a whole slab: starts on a block boundary, covers block_rows rows of one slice
        sy, sz = dims[0], dims[0] * dims[1]
        assert ((t["first_voxel"] % sy) == 0).all() and ((((t["first_voxel"] % sz) // sy) % block_rows) == 0).all()
        n = 4096
        got_xyz, got_v = _draw(vol, n)
        want_xyz, want_v, bad = O.ooc_sample(rng.state, n, t["first_voxel"], t["length"], block_rows, f32, dims, lo, hi)
        assert bad == 0                                       # one ghost row / slice suffices for every access
        assert np.array_equal(got_xyz, want_xyz) and np.array_equal(got_v, want_v)
        assert got_v.min() >= 0.0 and got_v.max() <= 1.0 and (got_xyz >= 0).all() and (got_xyz < 1).all()
    # the pool turns over: 8 of 40 slots are replaced per step
    changed = [(a != b).sum() for a, b in zip(tables[:-1], tables[1:])]
    assert all(0 < c <= 8 for c in changed)
    assert vol.outofcore_info()["bytes_uploaded"] >= (40 + 3 * 8) * info["slot_bytes"]


def test_values_are_the_trilinear_interpolation_of_the_file_and_training_runs(tmp_path):
    """Size-independent property on a smooth volume: value == trilinear(normalised file) at the returned coordinate, checked
    with an independent numpy interpolation; then the training loop runs from the pool and learns the volume."""
    from instantvnr_b200 import synthetic as syn
    dims = (96, 80, 64)
    gt = syn.make_volume(dims, seed=4)
    raw = (gt * 1000.0).astype(np.uint16)
    path = tmp_path / "smooth.raw"
    raw.tofile(path)
    vol = vnr.NeuralVolume(CFG, dims)
    vol.set_groundtruth_outofcore(path, "uint16", (0.0, 1000.0), num_concurrent_blocks=64, num_blocks=1024)
    xyz, v = _draw(vol, 20000)
    norm = np.clip(raw.astype(np.float32) / np.float32(1000.0), 0, 1)
    p = xyz.astype(np.float64) * np.array(dims) - 0.5
    p = np.clip(p, 0, np.array(dims) - 1.0)
    i0 = np.floor(p).astype(int); w = p - i0
    i1 = np.minimum(i0 + 1, np.array(dims) - 1)
    def at(ix, iy, iz): return norm[iz, iy, ix].astype(np.float64)
    want = sum(at(*(np.where(c[d], i1[:, d], i0[:, d]) for d in range(3))) * np.prod([np.where(c[d], w[:, d], 1 - w[:, d]) for d in range(3)], axis=0)
               for c in [(a, b, e) for a in (0, 1) for b in (0, 1) for e in (0, 1)])
    assert np.abs(v - want).max() <= 2e-4                      # float32 coordinate rounding x the volume's gradient
    # samples land in every z-slice region over time (random slabs, random voxels)
    assert len(np.unique((xyz[:, 2] * dims[2]).astype(int))) > dims[2] // 2
    vol.init_params(5)
    rgb, alpha = syn.make_tfn(64)
    vol.set_transfer_function(rgb, alpha)
    vol.train(10, batch=8192, fast_mode=False)                # online macrocell update from the drawn samples
    _, loss0 = vol.stats()
    vol.train(290, batch=8192, fast_mode=False)
    step, _ = vol.stats()
    loss = vol.last_loss()
    assert step == 300 and loss < 0.5 * loss0
    # the decoded network approximates the file
    zz, yy, xx = np.meshgrid(*[(np.arange(d, dtype=np.float32) + 0.5) / d for d in dims[::-1]], indexing="ij")
    idx = np.random.default_rng(0).integers(0, norm.size, 20000)
    pts = np.stack([xx.ravel()[idx], yy.ravel()[idx], zz.ravel()[idx]], 1)
    assert np.abs(vol.decode_host(pts) - norm.ravel()[idx]).mean() < 0.5 * np.abs(norm.ravel()[idx]).mean()


def test_outofcore_errors(tmp_path):
    vol = vnr.NeuralVolume(CFG, (16, 16, 16))
    short = tmp_path / "short.raw"; short.write_bytes(b"\x01" * 100)
    with pytest.raises(vnr.VnrError) as e:
        vol.set_groundtruth_outofcore(short, "uint8", (0, 255))
    assert e.value.code == -1 and "shorter" in str(e.value)
    ok = tmp_path / "ok.raw"; ok.write_bytes(b"\x01" * 4096)
    with pytest.raises(vnr.VnrError) as e:
        vol.set_groundtruth_outofcore(ok, "uint8", (5, 5))
    assert "valid value range" in str(e.value)               # neural_sampler.cpp:1068-1070
    with pytest.raises(vnr.VnrError) as e:
        vol.set_groundtruth_outofcore(ok, "uint8", (0, 255), num_concurrent_blocks=16, num_blocks=8)
    assert e.value.code == -1
    with pytest.raises(vnr.VnrError):
        vol.outofcore_info()


def test_data_parallel_ranks_keep_their_own_slabs_and_the_step_equals_accumulated_batches(tmp_path):
    """BASELINE configs[3] in small: out-of-core ground truth + data-parallel training behind vnr_comm.  Every rank refreshes its
    OWN random slabs (pcg32 stream 2 + rank of the slab selector), draws the r-th of `world` consecutive batches of the one
    sample stream from them, and the fused reduce-scatter + Adam + all-gather step equals one process accumulating the same
    batches (restated by the oracle from each rank's slab table) into one optimizer step."""
    import os
    os.environ.setdefault("VNR_COMM_SHARE_DEVICES", "1")       # a one-GPU box runs both ranks on the device
    from instantvnr_b200 import synthetic as syn
    dims, dtype, world, n, steps = (64, 48, 40), "float32", 2, 2048, 3
    gt = syn.make_volume(dims, seed=9)
    raw = (gt * 100.0 - 20.0).astype(np.float32)
    path = tmp_path / "dp.raw"; raw.tofile(path)
    lo, hi = -20.0, 80.0
    row = dims[0] * 4
    block_rows = min(-(-32768 // row), dims[1])

    def make(init):
        v = vnr.NeuralVolume(CFG, dims)
        v.set_groundtruth_outofcore(path, dtype, (lo, hi), num_concurrent_blocks=8, num_blocks=40)
        if init:
            v.init_params(21)
        return v

    solo = make(True)
    p0 = solo.get_params_f16()
    comms = vnr.Comm.init_local(world)
    vols = []
    for r, c in enumerate(comms):
        c.set_device()
        vols.append(make(r == 0))
    for v, c in zip(vols, comms):
        v.attach_comm(c)
    tabs = [v.outofcore_info(table=True) for v in vols]
    assert np.array_equal(tabs[0]["first_voxel"], solo.outofcore_info(table=True)["first_voxel"])    # rank 0 == a single process
    assert (tabs[0]["first_voxel"] != tabs[1]["first_voxel"]).mean() > 0.5                            # rank 1 drew its own slabs
    # the accumulated single-process run on the oracle's restatement of every rank's batch
    ref = vnr.NeuralVolume(CFG, dims)
    ref.set_params_f16(p0)
    rng = O.Rng(1337)
    f32 = raw.ravel()
    ref_losses, dp_losses = [], []
    for step in range(steps):
        tabs = [v.outofcore_info(table=True) for v in vols]            # the tables this step's batches are drawn from
        loss = 0.0
        for r in range(world):
            xyz, tgt, bad = O.ooc_sample(rng.state, n, tabs[r]["first_voxel"], tabs[r]["length"], block_rows, f32, dims, lo, hi)
            assert bad == 0
            ref.train_grads(torch.from_numpy(xyz).cuda(), torch.from_numpy(tgt).cuda(), n, n * world)
            loss += ref.last_loss()
        ref.optimizer_step()
        ref_losses.append(loss)
        for v in vols:
            v.train(1, batch=n, fast_mode=True)
        dp_losses.append(vols[0].last_loss())
    ps = [v.get_params_f16() for v in vols]
    assert np.array_equal(ps[0], ps[1])                                # replicas bit-identical
    assert np.allclose(dp_losses, ref_losses, rtol=2e-3), (dp_losses, ref_losses)
    a = ps[0].view(np.float16).astype(np.float32); b = ref.get_params_f16().view(np.float16).astype(np.float32)
    i0 = p0.view(np.float16).astype(np.float32)
    # same batches, same arithmetic up to where the fp16 hash-grid sums are rounded (per rank, then fp32, against one fp16 buffer)
    assert np.linalg.norm(a - b) < 0.1 * np.linalg.norm(b - i0) and np.abs(a - b).max() <= 2 * 5e-3
    # the ranks' pools keep diverging: different slots refreshed with different slabs
    t0, t1 = [v.outofcore_info(table=True)["first_voxel"] for v in vols]
    assert (t0 != t1).mean() > 0.5
    for v in vols:
        v.detach_comm()
    for c in comms:
        c.close()
