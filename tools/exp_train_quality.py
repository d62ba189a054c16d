"""Volume PSNR over training with FRESH batches every step (the product's sampler feeds both arms): ours with fp32 / fp16
weight-gradient accumulators against the reference's own tiny-cuda-nn trainer.  GPU box."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import instantvnr_b200 as vnr
import bench
from oracle import tcnn_ref
from test_gpu_fullsize import _ref_psnr

torch.cuda.set_device(0)
DIMS = (256,) * 3
gt = bench.synth_volume_device(DIMS)
st = torch.cuda.Stream()
rng = float(gt.max() - gt.min())
print("constant-predictor PSNR (mean of the volume):", 10 * np.log10(rng * rng / float(((gt - gt.mean()) ** 2).mean())))


def run(n, steps, marks, seed=1337):
    vols = {}
    for name, flags in (("fp32", 0), ("fp16", 64)):
        v = vnr.NeuralVolume(vnr.example_model_json(), DIMS); v.set_groundtruth_device(gt); v.init_params(seed); v.train_debug(1, flags, False)
        vols[name] = v
    ref = tcnn_ref.RefNetwork(vnr.example_model_json(), seed)
    src = vols["fp32"]
    xyz = torch.empty(n, 3, device="cuda"); tgt = torch.empty(n, device="cuda")
    for i in range(1, steps + 1):
        src.sample(xyz, tgt, n); torch.cuda.synchronize()
        for v in vols.values():
            v.train_on(xyz, tgt, n)
        with torch.cuda.stream(st):
            ref.training_step(xyz, tgt, n, st.cuda_stream, want_loss=False)
        st.synchronize(); torch.cuda.synchronize()
        if i in marks:
            print(f"  batch {n} step {i}: PSNR ours fp32-wgrad {vols['fp32'].psnr():.2f}  ours fp16-wgrad {vols['fp16'].psnr():.2f}  reference {_ref_psnr(ref, gt, st):.2f}   "
                  f"(last loss {vols['fp32'].last_loss():.5f} / {vols['fp16'].last_loss():.5f})", flush=True)


for seed in (1337, 7):
    print("seed", seed)
    run(1 << 16, 1000, (100, 300, 1000), seed)
    run(1 << 18, 600, (100, 300, 600), seed)
