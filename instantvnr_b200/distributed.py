"""Multi-GPU drivers for the two places where the path shards: data-parallel training with a
gradient all-reduce, and tile-parallel rendering with the frame gathered on rank 0.

One process per GPU; `torch.distributed` provides the process group (NCCL over NVLink on the GPU box, gloo
in the CPU tests).  Nothing here computes: the per-rank work is done by the C-ABI library through the
backend object (`GpuTrainBackend` wraps a `NeuralVolume`; the CPU tests plug an oracle-backed stand-in with
the same five methods), this module only orders the calls and the collectives.

Training schedule (the reference trains on one GPU; `Trainer::training_step` tcnn trainer.h:211-247):
every step each rank draws its own batch of `n` samples from the ONE global pcg32 sampler stream -- rank r
skips r*3n uniforms, draws 3n, skips (world-1-r)*3n, so the union over ranks is exactly what `world`
consecutive single-GPU sampler calls would draw -- runs forward + L1 loss + backward with the loss
normalised by the GLOBAL batch `world*n`, all-reduces (sum) the loss-scaled gradients (hash-grid: fp16 on
the GPU, MLP: fp32) and applies the identical optimizer step on every rank.  The result equals a single
process that accumulates the gradients of `world` consecutive batches and steps once.
"""
import numpy as np
import torch
import torch.distributed as dist


def strip_rows(height, rank, world, strip=4):
    """image rows owned by `rank`: strips of `strip` rows dealt round-robin (march.cuh ray_to_pixel)"""
    return [y for y in range(height) if (y // strip) % world == rank]


def sampler_schedule(rank, world, n, uniforms_per_sample=3):
    """(uniforms to skip before, uniforms to skip after) this rank's draw of n samples in one DP step"""
    return rank * uniforms_per_sample * n, (world - 1 - rank) * uniforms_per_sample * n


def make_peer_barrier(group=None):
    """a vnr.PeerBarrier attached to every rank of `group` (None when the group has one rank)"""
    import instantvnr_b200 as vnr
    world = dist.get_world_size(group)
    if world == 1:
        return None
    b = vnr.PeerBarrier()
    handles = [None] * world
    dist.all_gather_object(handles, b.handle, group=group)
    b.attach(dist.get_rank(group), world, b"".join(handles))
    dist.barrier(group=group)          # every rank has mapped every flag array before the first sync
    return b


def _wrap_device(ptr, n, dtype):
    """zero-copy torch view of a device buffer owned by the library (CUDA array interface)"""
    typestr = {torch.float32: "<f4", torch.float16: "<f2"}[dtype]

    class _A:
        __cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}
    return torch.as_tensor(_A(), device="cuda")


class GpuTrainBackend:
    """The five per-rank operations of a DP step on a `NeuralVolume` (all enqueued on the volume's stream)."""

    def __init__(self, vol):
        self.vol = vol
        self.stream = torch.cuda.ExternalStream(vol.stream())
        pm, nm, _ = vol.grad_buffer(0)
        pg, ng, _ = vol.grad_buffer(1)
        self.g_mlp = _wrap_device(pm, nm, torch.float32)
        self.g_grid = _wrap_device(pg, ng, torch.float16)
        self._xyz = self._tgt = None
        pmc, nmc = vol.macrocell_buffer()
        self.mc = _wrap_device(pmc, nmc, torch.float32).view(-1, 2)

    def sampler_skip(self, n_floats):
        self.vol.sampler_skip(n_floats)

    def sample(self, n):
        if self._xyz is None or self._xyz.shape[0] != n:
            self._xyz = torch.empty(n, 3, device="cuda"); self._tgt = torch.empty(n, device="cuda")
        self.vol.sample(self._xyz, self._tgt, n)
        return self._xyz, self._tgt

    def grads(self, xyz, tgt, n, n_global):
        self.vol.train_grads(xyz, tgt, n, n_global)
        return [self.g_mlp, self.g_grid]

    def apply(self):
        self.vol.optimizer_step()

    def local_loss(self):
        return self.vol.last_loss()

    def macrocell(self, xyz, tgt, n):
        self.vol.macrocell_update(xyz, tgt, n)
        return self.mc

    # -- peer-memory optimizer (vnr_volume_dp_*): reduce-scatter + Adam + all-gather in one kernel
    def attach_peers(self, group=None):
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        handles = [None] * world
        dist.all_gather_object(handles, self.vol.dp_export(), group=group)
        self.vol.dp_attach(rank, world, b"".join(handles))
        self.barrier = make_peer_barrier(group)

    def apply_sharded(self, group=None):
        """every rank: barrier (all gradients complete) -> fused kernel -> barrier (all parameters complete) -> clear"""
        s = self.vol.stream()
        self.barrier.sync(s)
        self.vol.dp_optimizer_step()
        self.barrier.sync(s)
        self.vol.dp_finish_step()


class DataParallelTrainer:
    """vnrNeuralVolumeTrain across ranks: per-rank batches, one gradient all-reduce per step."""

    def __init__(self, backend, group=None, mode="allreduce"):
        """mode "allreduce": NCCL all-reduce of the gradient buffers + replicated optimizer; mode "sharded": the
        optimizer runs over peer memory (NVLink P2P), each rank owning 1/world of the hash-grid parameters."""
        self.b = backend
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.last_loss = None
        self.mode = mode if self.world > 1 else "allreduce"
        if self.mode == "sharded":
            self.b.attach_peers(group)

    def _stream_ctx(self):
        st = getattr(self.b, "stream", None)
        return torch.cuda.stream(st) if st is not None else _Null()

    def step(self, n_per_rank, fast_mode=True, want_loss=False):
        before, after = sampler_schedule(self.rank, self.world, n_per_rank, getattr(getattr(self.b, "vol", None), "uniforms_per_sample", 3))
        with self._stream_ctx():
            self.b.sampler_skip(before)
            xyz, tgt = self.b.sample(n_per_rank)
            self.b.sampler_skip(after)
            bufs = self.b.grads(xyz, tgt, n_per_rank, n_per_rank * self.world)
            if self.mode == "sharded":
                self.b.apply_sharded(self.group)
            else:
                if self.world > 1:
                    for g in bufs:
                        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
                self.b.apply()
            if not fast_mode:
                # online macrocell value ranges (core/network.cu:249-257): merge the ranks' (min-1, max+1) pairs
                mc = self.b.macrocell(xyz, tgt, n_per_rank)
                if self.world > 1:
                    lo = mc[:, 0].contiguous(); hi = mc[:, 1].contiguous()
                    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=self.group)
                    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=self.group)
                    mc[:, 0] = lo; mc[:, 1] = hi
            if want_loss:
                bar = getattr(self.b, "barrier", None)
                if bar is not None and bar.timed_out():       # synchronises: a rank that never arrived leaves incomplete gradients behind
                    raise RuntimeError("data-parallel step: a rank did not reach the peer barrier within 5 s; the parameters are not valid")
                loss = torch.tensor([self.b.local_loss()], dtype=torch.float64)
                if self.world > 1:
                    dev = bufs[0].device
                    loss = loss.to(dev)
                    dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=self.group)
                self.last_loss = float(loss.item())
        return self.last_loss

    def train(self, steps, n_per_rank, fast_mode=True):
        for _ in range(steps):
            self.step(n_per_rank, fast_mode)


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def broadcast_params(vol, group=None, src=0):
    """replicate the fp16 parameter blob of rank `src` (after loading / before DP training)"""
    p = torch.from_numpy(vol.get_params_f16().view(np.float16)).cuda()      # bit pattern travels unchanged
    dist.broadcast(p, src=src, group=group)
    vol.set_params_f16(p.cpu().numpy().view(np.uint16))


def gather_strips(frame, rows, rank, world, group=None):
    """Collect every rank's strips of `frame` (h, w, 4) into rank 0's copy.  Strip counts can differ by one
    between ranks (partial last strip), so every rank sends a block padded to the largest count."""
    if world == 1:
        return
    n_max = max(len(r) for r in rows)
    mine = frame.new_zeros((n_max,) + tuple(frame.shape[1:]))
    mine[: len(rows[rank])] = frame[rows[rank]]
    if rank == 0:
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.gather(mine, parts, dst=0, group=group)
        for r in range(1, world):
            frame[rows[r]] = parts[r][: len(rows[r])]
    else:
        dist.gather(mine, None, dst=0, group=group)


class TileParallelRenderer:
    """vnrRender across ranks: every rank marches its interleaved pixel strips; finished pixels land in rank
    0's frame buffer.  mode "p2p": the compositing kernel of every rank stores straight into rank 0's buffer
    over NVLink (CUDA IPC mapping, vnr_renderer_set_frame_target) and a barrier closes the frame; mode
    "nccl": ranks send their strips with a gather."""

    def __init__(self, renderer, group=None, mode="p2p"):
        import instantvnr_b200 as vnr
        self.ren, self.group, self.mode = renderer, group, mode
        self.rank = dist.get_rank(group); self.world = dist.get_world_size(group)
        self.w, self.h = renderer.size
        renderer.set_partition(self.rank, self.world)
        self.stream = torch.cuda.ExternalStream(renderer.stream())
        self.frame = _wrap_device(renderer.device_frame(), self.w * self.h * 4, torch.float32).view(self.h, self.w, 4)
        self._peer = None
        if self.world > 1 and mode == "p2p":
            handles = [None] * self.world
            dist.all_gather_object(handles, vnr.ipc_export(renderer.device_frame()) if self.rank == 0 else b"", group=group)
            if self.rank != 0:
                self._peer = vnr.ipc_open(handles[0])
                renderer.set_frame_target(self._peer)
        elif self.world > 1:
            self.rows = [torch.tensor(strip_rows(self.h, r, self.world), device="cuda", dtype=torch.long) for r in range(self.world)]
        self.download = True
        self.barrier = None
        if self.world > 1:
            renderer.set_download(False)          # rank 0 downloads after the peers' pixels have arrived
            if mode == "p2p":
                self.barrier = make_peer_barrier(group)

    def render(self):
        """one frame; on return (all ranks) the frame in rank 0's device buffer is complete on rank 0's stream"""
        if self.barrier is not None and self.download:
            # rank 0's download of the previous frame (enqueued on its stream) must finish before a peer's
            # ray-generation kernel starts storing pixels of this frame into the same buffer
            self.barrier.sync(self.ren.stream())
        self.ren.render()
        if self.world == 1:
            return
        with torch.cuda.stream(self.stream):
            if self.mode == "p2p":
                # every rank's stores must have completed before rank 0 reads / downloads the frame, and rank 0 must
                # have consumed the previous frame before anybody overwrites it: one peer-memory barrier kernel on
                # the renderer streams (stream-ordered, no host sync, no collective call)
                self.barrier.sync(self.ren.stream())
            else:
                gather_strips(self.frame, self.rows, self.rank, self.world, self.group)
        if self.rank == 0 and self.download:
            self.ren.download()

    def map_frame(self, copy=True):
        if self.rank != 0:
            return None
        img = self.ren.map_frame(copy)                         # synchronises the frame
        if self.barrier is not None and self.barrier.timed_out():
            raise RuntimeError("tile-parallel frame: a rank did not reach the peer barrier within 5 s; the frame is incomplete")
        return img

    def close(self):
        import instantvnr_b200 as vnr
        if self._peer:
            self.ren.set_frame_target(None)
            vnr.ipc_close(self._peer)
            self._peer = None
