"""VAR 2 with fp32 weight-gradient accumulators at full size: error per matrix against the oracle's exact sums (GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import instantvnr_b200 as vnr
import oracle as O
import test_gpu_fullsize as T
O.use_host_cores()
N, DIMS = T.N, T.DIMS
vol = vnr.NeuralVolume(vnr.example_model_json(), DIMS)
vol.set_groundtruth_device(T._synth_device(DIMS))
vol.init_params(1337)
vol.train(40, batch=1 << 16, fast_mode=True)
p16 = vol.get_params_f16()
m = O.ModelCfg()
xyz = torch.empty(N, 3, device="cuda"); tgt = torch.empty(N, device="cuda")
vol.sample(xyz, tgt, N); torch.cuda.synchronize()
c, t = xyz.cpu().numpy(), tgt.cpu().numpy()
tr = O.Trainer(m, O.f16_to_f32(p16)); tr.set_wgrad_slices(148)
tr.grads_only(c, t, N, 0, 0)
exact = tr.grads().copy()[:m.n_mlp]
sc = np.abs(exact).max()
W, E, NH = 64, m.enc_pad, m.n_hidden
offs = [0, W * E] + [W * E + k * W * W for k in range(1, NH)] + [W * E + (NH - 1) * W * W + 16 * W]
reps = int(os.environ.get("REPS", "12"))
for variant in (2, 1):
    worst = np.zeros(NH + 1); bad = 0
    for it in range(reps):
        vol.set_params_f16(p16)
        vol.train_debug(variant, 0, False)                 # the test's sequence: a half-mode launch, the optimizer step, then fp32 mode
        vol.train_grads(xyz, tgt, N, N); torch.cuda.synchronize()
        vol.optimizer_step(); vol.set_params_f16(p16)
        vol.train_debug(variant, 64, False)
        vol.train_grads(xyz, tgt, N, N); torch.cuda.synchronize()
        gm, _ = vol.get_grads()
        vol.optimizer_step()
        errs = np.array([float(np.abs(gm[offs[k]:offs[k + 1]] - exact[offs[k]:offs[k + 1]]).max() / sc) for k in range(NH + 1)])
        worst = np.maximum(worst, errs)
        if errs.max() > 1e-4:
            bad += 1
            k = int(errs.argmax()); d = np.abs(gm[offs[k]:offs[k + 1]] - exact[offs[k]:offs[k + 1]])
            print(f"variant {variant} iteration {it}: per-matrix err {np.round(errs, 6)}; matrix {k}: {int((d > 1e-4 * sc).sum())} entries off, first at {np.argwhere(d > 1e-4 * sc)[:6].ravel()}", flush=True)
    print(f"variant {variant}: {bad} of {reps} runs off; worst per-matrix err {np.round(worst, 6)}", flush=True)
