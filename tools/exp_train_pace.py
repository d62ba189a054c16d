"""Training kernel: pacing of the scatter reductions and gather depth (GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import instantvnr_b200 as vnr
import bench
torch.cuda.set_device(0)
n = 1 << 18
dims = (256,) * 3
gt = bench.synth_volume_device(dims)
vol = vnr.NeuralVolume(vnr.model_json(), dims)
vol.set_groundtruth_device(gt); vol.init_params(1337)
vol.train(50, batch=1 << 16, fast_mode=True)
st = torch.cuda.ExternalStream(vol.stream())
xyz = torch.empty(n, 3, device="cuda"); tgt = torch.empty(n, device="cuda")
vol.sample(xyz, tgt, n); torch.cuda.synchronize()


def time_kernel(reps=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        vol.train_grads(xyz, tgt, n, n)
    e0.record(st)
    for _ in range(reps):
        vol.train_grads(xyz, tgt, n, n)
    e1.record(st); st.synchronize()
    vol.optimizer_step(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for name, flags in (("baseline", 0), ("gather only (no chain, no reductions)", 5), ("scatter only (no chain, no loads)", 6), ("gather + scatter, no chain", 4),
                    ("chain + gather", 1), ("chain + scatter", 2), ("chain only", 3), ("all, two levels in flight", 128), ("gather only, two levels in flight", 128 | 5)):
    vol.train_debug(1, flags, False)
    print(f"{name:45s} {time_kernel():7.1f} us", flush=True)
# role timers with the chain off
NAMES = ["total", "wait_x0_full", "wait_mma", "wait_dx_empty", "bar", "wait_wgrad", "g_wait_empty", "g_work", "s_wait_full", "s_work", "tiles"]
for flags in (5, 6, 4):
    vol.train_debug(1, flags, True)
    vol.train_grads(xyz, tgt, n, n)
    prof = vol.train_profile().astype(np.float64).mean(0)
    vol.optimizer_step()
    print(flags, {k: round(float(prof[i])) for i, k in enumerate(NAMES)}, flush=True)
