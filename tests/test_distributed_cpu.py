"""World-size-2 `gloo` tests (CPU) of the multi-GPU host logic in instantvnr_b200/distributed.py: the
data-parallel training schedule (sampler sub-streams, global-batch normalisation, gradient all-reduce,
identical optimizer step) and the tile-parallel pixel partition.  The per-rank compute is the CPU oracle
standing in for the CUDA library behind the same backend interface (the oracle is the checker here; the
product path is exercised by tests/test_gpu_distributed.py on the GPU box)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as O
from instantvnr_b200 import synthetic as syn
from instantvnr_b200.distributed import DataParallelTrainer, gather_strips, sampler_schedule, strip_rows

CFG = dict(n_levels=4, n_features=4, log2_hashmap=10, base_res=4, n_hidden=2)
DIMS = (16, 16, 16)
N, STEPS = 512, 4


def _model():
    return O.ModelCfg(CFG["n_levels"], CFG["n_features"], CFG["log2_hashmap"], CFG["base_res"], 2.0, CFG["n_hidden"])


class OracleBackend:
    """same interface as distributed.GpuTrainBackend, computed by the CPU oracle"""

    def __init__(self, m, p32, gt):
        self.m, self.gt = m, gt
        self.tr = O.Trainer(m, p32)
        self.rng = O.Rng(1337)
        self.g = None
        self._loss = 0.0

    def sampler_skip(self, n_floats):
        if n_floats:
            self.rng.uniform(n_floats)

    def sample(self, n):
        return O.sample_batch(self.rng, n, self.gt, DIMS)

    def grads(self, xyz, tgt, n, n_global):
        self._loss = self.tr.grads_only(xyz, tgt, n_global)
        self.g = torch.from_numpy(self.tr.grads().copy())
        return [self.g]

    def apply(self):
        self.tr.apply(self.g.numpy())

    def local_loss(self):
        return self._loss

    def macrocell(self, xyz, tgt, n):
        raise NotImplementedError


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _dp_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = _model()
        p32, _ = O.init_params(m, 9)
        gt = syn.make_volume(DIMS, seed=3)
        dp = DataParallelTrainer(OracleBackend(m, p32, gt))
        assert (dp.rank, dp.world) == (rank, world)
        losses = [dp.step(N, want_loss=True) for _ in range(STEPS)]
        p16, master = dp.b.tr.params()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), p16=p16, master=master, losses=np.array(losses))
    finally:
        dist.destroy_process_group()


def test_data_parallel_step_equals_accumulating_consecutive_batches(tmp_path):
    world = 2
    mp.spawn(_dp_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"rank{k}.npz") for k in range(world)]
    # replicas stay bit-identical
    assert np.array_equal(r[0]["p16"], r[1]["p16"]) and np.array_equal(r[0]["master"], r[1]["master"])
    assert np.array_equal(r[0]["losses"], r[1]["losses"])
    # single process: `world` consecutive sampler calls per step, gradients accumulated, one optimizer step
    m = _model()
    p32, _ = O.init_params(m, 9)
    gt = syn.make_volume(DIMS, seed=3)
    tr, rng = O.Trainer(m, p32), O.Rng(1337)
    losses = []
    for _ in range(STEPS):
        g, loss = np.zeros(m.n_params, np.float32), 0.0
        for _k in range(world):
            c, t = O.sample_batch(rng, N, gt, DIMS)
            loss += tr.grads_only(c, t, world * N)
            g = g + tr.grads()
        tr.apply(g)
        losses.append(loss)
    p16, master = tr.params()
    assert np.array_equal(p16, r[0]["p16"])
    assert np.array_equal(master, r[0]["master"])
    assert np.allclose(losses, r[0]["losses"], rtol=1e-12)
    assert losses[-1] < losses[0]


def test_sampler_schedule_tiles_the_global_stream():
    for world in (1, 2, 3, 8):
        n = 640
        total = 0
        for rank in range(world):
            before, after = sampler_schedule(rank, world, n)
            assert before == rank * 3 * n and before + 3 * n + after == world * 3 * n
            total += 3 * n
        assert total == world * 3 * n
    # rank r's draw equals the r-th of `world` consecutive single-process draws
    gt = syn.make_volume(DIMS, seed=3)
    ref = O.Rng(1337)
    consecutive = [O.sample_batch(ref, 256, gt, DIMS)[0] for _ in range(3)]
    for rank in range(3):
        rng = O.Rng(1337)
        before, _ = sampler_schedule(rank, 3, 256)
        if before:
            rng.uniform(before)
        assert np.array_equal(O.sample_batch(rng, 256, gt, DIMS)[0], consecutive[rank])


def _gather_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        h, w = 50, 24                                   # 50 rows: the last strip is partial
        full = torch.arange(h * w * 4, dtype=torch.float32).view(h, w, 4)
        rows = [torch.tensor(strip_rows(h, r, world), dtype=torch.long) for r in range(world)]
        frame = torch.zeros(h, w, 4)
        frame[rows[rank]] = full[rows[rank]]            # this rank "rendered" only its strips
        gather_strips(frame, rows, rank, world)
        if rank == 0:
            np.save(os.path.join(out_dir, "frame.npy"), frame.numpy())
    finally:
        dist.destroy_process_group()


def test_tile_partition_gathers_to_the_full_frame(tmp_path):
    for world in (1, 2, 5):
        rows = [strip_rows(50, r, world) for r in range(world)]
        assert sorted(sum(rows, [])) == list(range(50))                # disjoint cover
        assert max(len(x) for x in rows) - min(len(x) for x in rows) <= 4
    world = 3
    mp.spawn(_gather_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    frame = np.load(tmp_path / "frame.npy")
    assert np.array_equal(frame.ravel(), np.arange(50 * 24 * 4, dtype=np.float32))
