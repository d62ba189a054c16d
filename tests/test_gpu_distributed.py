"""Two-GPU (NCCL) tests of the sharded paths on the product library: data-parallel training with the
gradient all-reduce, tile-parallel rendering with the frame gathered on rank 0 (peer stores and NCCL
gather).  Skipped on a single-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_distributed.py`."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import instantvnr_b200 as vnr
import oracle as O
from instantvnr_b200 import synthetic as syn
from instantvnr_b200.distributed import DataParallelTrainer, GpuTrainBackend, TileParallelRenderer, broadcast_params

pytestmark = pytest.mark.gpu

CFG = dict(n_levels=4, n_features=8, log2_hashmap=12, base_res=8, n_hidden=2)
DIMS = (32, 32, 32)
N, STEPS = 2048, 12


def _need_two_gpus():
    if vnr.device_count() < 2:
        pytest.skip("needs 2 GPUs")


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _init(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))


def _dp_worker(rank, world, port, out_dir, mode):
    _init(rank, world, port)
    try:
        vol = vnr.NeuralVolume(vnr.model_json(**CFG), DIMS)
        vol.set_groundtruth(syn.make_volume(DIMS, seed=3))
        vol.init_params(21 + rank)                 # deliberately different: the broadcast must replicate rank 0
        broadcast_params(vol)
        rgb, alpha = syn.make_tfn(32)
        vol.set_transfer_function(rgb, alpha)
        dp = DataParallelTrainer(GpuTrainBackend(vol), mode=mode)
        assert dp.mode == mode
        losses = [dp.step(N, fast_mode=False, want_loss=True) for _ in range(STEPS)]
        torch.cuda.synchronize()
        md, vr, mo = vol.get_macrocell()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), p16=vol.get_params_f16(), losses=np.array(losses), vr=vr, step=vol.stats()[0])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["allreduce", "sharded"])
def test_data_parallel_training_two_gpus(tmp_path, mode):
    _need_two_gpus()
    world = 2
    mp.spawn(_dp_worker, args=(world, _free_port(), str(tmp_path), mode), nprocs=world, join=True)
    r = [np.load(tmp_path / f"rank{k}.npz") for k in range(world)]
    assert np.array_equal(r[0]["p16"], r[1]["p16"])                 # replicas stay bit-identical
    assert np.array_equal(r[0]["vr"], r[1]["vr"])                   # merged macrocell value ranges
    assert int(r[0]["step"]) == STEPS
    # oracle: gradients of `world` consecutive batches accumulated, one step (tests/test_distributed_cpu.py)
    m = O.ModelCfg(CFG["n_levels"], CFG["n_features"], CFG["log2_hashmap"], CFG["base_res"], 2.0, CFG["n_hidden"])
    p32, _ = O.init_params(m, 21)
    gt = syn.make_volume(DIMS, seed=3)
    tr, rng = O.Trainer(m, p32), O.Rng(1337)
    want = []
    mc = np.zeros(2 * int(np.prod(O.macrocell_dims(DIMS))), np.float32)
    for _ in range(STEPS):
        g, loss = np.zeros(m.n_params, np.float32), 0.0
        for _k in range(world):
            c, t = O.sample_batch(rng, N, gt, DIMS)
            loss += tr.grads_only(c, t, world * N)
            g = g + tr.grads()
            O.macrocell_update_explicit(c, t, DIMS, mc)
        tr.apply(O.f16_to_f32(O.f32_to_f16(g)))
        want.append(loss)
    got = r[0]["losses"]
    assert got[0] == pytest.approx(want[0], rel=1e-4)
    assert np.abs(got - np.array(want)).max() <= 0.02 * max(want)   # loss within 2 % of the oracle
    assert np.array_equal(r[0]["vr"], mc)                           # min/max merge is exact


def _render_worker(rank, world, port, out_dir, mode):
    _init(rank, world, port)
    try:
        m = O.ModelCfg(8, 8, 15, 16, 2.0, 4)
        p32, _ = O.init_params(m, 7)
        p32 = p32.copy(); p32[m.n_mlp:] *= 3000.0
        vol = vnr.NeuralVolume(vnr.model_json(log2_hashmap=15), (64, 64, 64))
        vol.set_params_f16(O.f32_to_f16(p32))
        rgb, alpha = syn.make_tfn(64)
        dec = vol.decode_host(np.random.default_rng(1).random((20000, 3), dtype=np.float32))
        vol.set_transfer_function(rgb, alpha, (max(float(dec.min()), 0.0), min(float(dec.max()), 1.0)))   # part of the volume opaque
        vol.set_macrocell(np.tile(np.array([-1.0, 2.0], np.float32), 64))     # every cell [0,1]: nothing skipped
        w, h = 96, 50                                                          # 50 rows: partial last strip
        ren = vnr.Renderer(vol)
        ren.set_size(w, h)
        tp = TileParallelRenderer(ren, mode=mode)
        frames = []
        for view in (1, 4):
            ren.set_camera(*syn.default_camera((64, 64, 64), view))
            tp.render()
            f = tp.map_frame()
            if rank == 0:
                frames.append(f)
        if rank == 0:
            single = vnr.Renderer(vol)
            single.set_size(w, h)
            ref = []
            for view in (1, 4):
                single.set_camera(*syn.default_camera((64, 64, 64), view))
                single.render()
                ref.append(single.map_frame())
            np.savez(os.path.join(out_dir, f"frames_{mode}.npz"), got=np.stack(frames), ref=np.stack(ref))
        dist.barrier()
        tp.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["p2p", "nccl"])
def test_tile_parallel_frame_equals_single_gpu(tmp_path, mode):
    _need_two_gpus()
    world = 2
    mp.spawn(_render_worker, args=(world, _free_port(), str(tmp_path), mode), nprocs=world, join=True)
    d = np.load(tmp_path / f"frames_{mode}.npz")
    assert d["ref"][..., 3].max() > 0.3
    assert np.array_equal(d["got"], d["ref"])
