"""Where does the training step's time go?  (GPU box; writes gpurun_out/exp_train_roles.json)
  * memory-system probes: random 16-byte loads over 46.7 MB / 306.8 MB, random fp16x8 reductions over 46.7 MB
  * the training kernel alone (train_grads on a fixed batch), both chain variants, with roles switched off
  * per-CTA role timers (what each group waits for)
  * full steps/s through vnr_volume_train
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import instantvnr_b200 as vnr          # noqa: E402
import bench                           # noqa: E402

out = {}
torch.cuda.set_device(0)
n = 1 << 18

T0 = time.time()


def log(*a):
    print(f"[{time.time() - T0:6.1f}s]", *a, flush=True)


# ---- probes
TAB19, TAB22 = 2920448 * 16, 19173376 * 16
for name, kind, tab, ops in (("loads_46MB", "loads", TAB19, (1 << 24) * 64), ("loads_307MB", "loads", TAB22, (1 << 24) * 64),
                             ("reds_46MB_2p18", "reds", TAB19, n * 64), ("reds_46MB_2p22", "reds", TAB19, (1 << 22) * 64),
                             ("reds_307MB_2p22", "reds", TAB22, (1 << 22) * 64), ("copy_1GiB", "copy", 1 << 30, 0)):
    best, mean = vnr.probe_memory(kind, tab, ops, 5)
    rec = {"ms_best": best, "ms_mean": mean}
    if kind == "copy":
        rec["GBps"] = 2 * tab / (best * 1e-3) / 1e9
    else:
        rec["Gops"] = ops / (best * 1e-3) / 1e9
        rec["GBps_16B"] = ops * 16 / (best * 1e-3) / 1e9
    out[name] = rec
    log(name, rec)

# ---- training kernel
dims = (256,) * 3
gt = bench.synth_volume_device(dims)
vol = vnr.NeuralVolume(vnr.model_json(), dims)
vol.set_groundtruth_device(gt)
vol.init_params(1337)
vol.train(50, batch=1 << 16, fast_mode=True)
st = torch.cuda.ExternalStream(vol.stream())
xyz = torch.empty(n, 3, device="cuda"); tgt = torch.empty(n, device="cuda")
vol.sample(xyz, tgt, n)
torch.cuda.synchronize()


def time_kernel(reps=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        vol.train_grads(xyz, tgt, n, n)
    e0.record(st)
    for _ in range(reps):
        vol.train_grads(xyz, tgt, n, n)
    e1.record(st)
    st.synchronize()
    vol.optimizer_step()           # consumes / clears the accumulated gradients
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3      # us (includes the small MLP-partial reduction kernel + memset)


def time_steps(reps=50):
    vol.train(5, batch=n, fast_mode=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    vol.train(reps, batch=n, fast_mode=True)
    e1.record(st)
    st.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


NAMES = ["total", "wait_x0_full", "wait_mma", "wait_dx_empty", "bar", "wait_wgrad", "g_wait_empty", "g_work", "s_wait_full", "s_work", "tiles",
         "fwd_ldtm", "fwd_cvt_sts", "fwd_fence"]
for variant in (1,):
    for flags in (0, 32, 3, 35):
        vol.train_debug(variant, flags, False)
        us = time_kernel()
        out[f"kernel_us_v{variant}_f{flags}"] = us
        log(f"variant {variant} flags {flags}: train_grads {us:.1f} us")
    vol.init_params(1337)          # the switched-off roles left meaningless gradients behind
    vol.train(30, batch=1 << 16, fast_mode=True)
    vol.train_debug(variant, 0, True)
    vol.train_grads(xyz, tgt, n, n)
    prof = vol.train_profile().astype(np.float64)
    vol.optimizer_step()
    mean = prof.mean(0)
    rec = {k: float(mean[i]) for i, k in enumerate(NAMES)}
    out[f"roles_v{variant}"] = rec
    tot = rec["total"]
    log(f"variant {variant} role timers (cycles, mean over CTAs):", {k: (round(v), round(v / tot, 3)) for k, v in rec.items()})
    log("trace of CTA 0, tile 5 (cycles since tile start; all roles on):", [int(x) for x in prof[0, 16:52]])
    vol.train_debug(variant, 3, True)          # the chain alone (no gather loads, no reductions)
    vol.train_grads(xyz, tgt, n, n)
    prof2 = vol.train_profile().astype(np.float64)
    prof = prof2.mean(0)
    vol.optimizer_step()
    log(f"variant {variant} chain alone:", {k: round(float(prof[i])) for i, k in enumerate(NAMES)})
    log("trace of CTA 0, tile 5 (chain alone):", [int(x) for x in prof2[0, 16:52]])
    vol.init_params(1337)
    vol.train_debug(variant, 0, False)
    us = time_steps()
    out[f"step_us_v{variant}"] = us
    log(f"variant {variant}: full step {us:.1f} us = {1e6 / us:.0f} steps/s")
step, loss = vol.stats()
out["final_loss"] = loss
print("mean loss", loss, "psnr", vol.psnr())
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "exp_train_roles.json"), "w"), indent=1)
