"""Per-phase CUDA-event timing of one data-parallel training step (run under torchrun, N >= 2):
grads (fwd+loss+bwd), flag all-reduce (barrier), peer-memory optimizer kernel, NCCL all-reduce of the
gradient buffers, replicated optimizer.  Diagnostic only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import instantvnr_b200 as vnr
from instantvnr_b200.distributed import GpuTrainBackend
import bench

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dims = (256,) * 3
vol = vnr.NeuralVolume(vnr.model_json(), dims)
vol.set_groundtruth_device(bench.synth_volume_device(dims))
vol.init_params(1337)
b = GpuTrainBackend(vol)
b.attach_peers()
st = b.stream
n = 1 << 18
flag = torch.zeros(1, device="cuda")


def timed(name, fn, reps=30):
    with torch.cuda.stream(st):
        for _ in range(3):
            fn()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps):
            fn()
        e1.record(st); st.synchronize()
    if rank == 0:
        print(f"{name:40s} {e0.elapsed_time(e1) / reps * 1e3:9.1f} us", flush=True)


xyz, tgt = b.sample(n)
st.synchronize()
timed("sample", lambda: b.sample(n))
timed("grads (fwd+loss+bwd)", lambda: vol.train_grads(xyz, tgt, n, n * world))
timed("flag all-reduce (barrier)", lambda: dist.all_reduce(flag))
def sharded():
    vol.train_grads(xyz, tgt, n, n * world)
    vol.dp_optimizer_step()
timed("grads + peer-memory optimizer kernel", sharded)
def sharded_full():
    vol.train_grads(xyz, tgt, n, n * world)
    b.apply_sharded()
timed("grads + barriers + peer optimizer + clear", sharded_full)
def allreduce_only():
    dist.all_reduce(b.g_grid); dist.all_reduce(b.g_mlp)
timed("NCCL all-reduce of grid+MLP gradients", allreduce_only)
def plain():
    vol.train_grads(xyz, tgt, n, n * world)
    vol.optimizer_step()
timed("grads + replicated optimizer", plain)
vol.dp_detach()
dist.destroy_process_group()
