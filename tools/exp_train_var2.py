"""Training kernel variants side by side: kernel time with roles switched off, role timers, full step (GPU box)."""
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


variants = [int(v) for v in os.environ.get("VARIANTS", "1,2").split(",")]
NAMES = {0: "total", 1: "wait_x0_full", 2: "wait_D", 3: "wait_dx_empty", 4: "wait_epilogues", 5: "wait_wgrad", 6: "g_wait_empty", 7: "g_work", 8: "s_wait_full", 9: "s_work",
         10: "tiles", 12: "wait_side_stores"}
for var in variants:
    for name, flags in (("all", 0), ("chain + gather", 1), ("chain + scatter", 2), ("chain only", 3)):
        vol.train_debug(var, flags, False)
        print(f"variant {var} {name:20s} {time_kernel():7.1f} us", flush=True)
    for flags in (0, 3):
        vol.train_debug(var, flags, True)
        vol.train_grads(xyz, tgt, n, n)
        prof = vol.train_profile().astype(np.float64).mean(0)
        vol.optimizer_step()
        print(f"variant {var} flags {flags} role timers:", {v: round(float(prof[k])) for k, v in NAMES.items()}, flush=True)
    vol.train_debug(var, 0, False)
    vol.init_params(1337)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    vol.train(20, batch=n, fast_mode=True)
    e0.record(st); vol.train(200, batch=n, fast_mode=True); e1.record(st); st.synchronize()
    us = e0.elapsed_time(e1) / 200 * 1e3
    print(f"variant {var} full step {us:.1f} us = {1e6 / us:.0f} steps/s; loss {vol.stats()} psnr {vol.psnr():.2f}", flush=True)
