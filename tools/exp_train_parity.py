"""Training parity against the reference's own tiny-cuda-nn build as a function of the batch size (GPU box).
Same initial blob, same pre-drawn batches through vnr_volume_train_on and Trainer::training_step.  Prints the loss curves,
and after the FIRST step (Adam moves every parameter with a non-zero gradient by exactly +-lr) how the two gradient sign /
zero patterns compare.  flags=8 re-runs ours with activation gradients below the fp16 normal range flushed to zero."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import instantvnr_b200 as vnr          # noqa: E402
import bench                           # noqa: E402
from oracle import tcnn_ref            # noqa: E402

torch.cuda.set_device(0)
DIMS = (256,) * 3
gt = bench.synth_volume_device(DIMS)
st = torch.cuda.Stream()


def run(n, steps=8, flags=0, psnr_steps=0):
    vol = vnr.NeuralVolume(vnr.example_model_json(), DIMS)
    vol.set_groundtruth_device(gt)
    vol.init_params(1337)
    vol.train_debug(1, flags, False)
    ref = tcnn_ref.RefNetwork(vnr.example_model_json(), 1337)
    p0 = vol.get_params_f16().view(np.float16).astype(np.float32)
    n_mlp = vol.n_mlp_params
    ring = []
    for _ in range(8):
        xyz = torch.empty(n, 3, device="cuda"); tgt = torch.empty(n, device="cuda")
        vol.sample(xyz, tgt, n)
        ring.append((xyz, tgt))
    torch.cuda.synchronize()
    lo, lr_ = [], []
    first = None
    for i in range(max(steps, psnr_steps)):
        xyz, tgt = ring[i % 8]
        vol.train_on(xyz, tgt, n)
        with torch.cuda.stream(st):
            l = ref.training_step(xyz, tgt, n, st.cuda_stream, want_loss=i < steps)
        st.synchronize()
        if i < steps:
            lo.append(vol.last_loss()); lr_.append(l)
        if i == 0:
            a = vol.get_params_f16().view(np.float16).astype(np.float32) - p0
            b = ref.get_params_f16().view(np.float16).astype(np.float32) - p0
            first = (a, b)
    print(f"--- batch {n} flags {flags}")
    print("  loss ours  ", np.round(lo, 5))
    print("  loss theirs", np.round(lr_, 5))
    a, b = first
    for name, sl in (("mlp", slice(0, n_mlp)), ("grid", slice(n_mlp, None))):
        x, y = a[sl], b[sl]
        both = (x != 0) & (y != 0)
        print(f"  first step {name}: moved ours {np.mean(x != 0):.4f} theirs {np.mean(y != 0):.4f} | only ours {np.mean((x != 0) & (y == 0)):.4f} "
              f"only theirs {np.mean((x == 0) & (y != 0)):.4f} | same direction among both-moved {np.mean(np.sign(x[both]) == np.sign(y[both])):.4f}")
    if psnr_steps:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from test_gpu_fullsize import _ref_psnr
        print(f"  after {psnr_steps} steps: PSNR ours {vol.psnr():.3f} dB, theirs {_ref_psnr(ref, gt, st):.3f} dB")


if len(sys.argv) > 1:          # n:flags:psnr_steps ...
    for spec in sys.argv[1:]:
        n_, f_, p_ = (int(x) for x in spec.split(":"))
        run(n_, flags=f_, psnr_steps=p_)
else:
    for n in (1 << 12, 1 << 14, 1 << 16, 1 << 18):
        run(n)
    run(1 << 18, flags=8)
    run(1 << 16, psnr_steps=300)
    run(1 << 18, flags=8, psnr_steps=300)
