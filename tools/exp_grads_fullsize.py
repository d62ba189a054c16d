"""Gradients of ONE training batch at BASELINE size against the CPU oracle (fp32 accumulation), by batch size: does anything go
wrong once a CTA processes more than one tile (n > 148 * 128)?  GPU box."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import instantvnr_b200 as vnr
import oracle as O
import bench

O.use_host_cores()
DIMS = (256,) * 3
gt = bench.synth_volume_device(DIMS)
m = O.ModelCfg()                        # example-model.json
vol = vnr.NeuralVolume(vnr.example_model_json(), DIMS)
vol.set_groundtruth_device(gt)
vol.init_params(1337)
vol.train(40, batch=1 << 16, fast_mode=True)          # a table that matters
p16 = vol.get_params_f16()
p32 = O.f16_to_f32(p16)
for n in (1 << 14, 1 << 16, 1 << 18):
    xyz = torch.empty(n, 3, device="cuda"); tgt = torch.empty(n, device="cuda")
    vol.sample(xyz, tgt, n); torch.cuda.synchronize()
    for flags in (0, 64):
        vol.train_debug(1, flags, False)
        vol.train_grads(xyz, tgt, n, n); torch.cuda.synchronize()
        gm, gg16 = vol.get_grads(); gg = O.f16_to_f32(gg16)
        loss = vol.last_loss()
        if flags == 0:
            t0 = time.time()
            tr = O.Trainer(m, p32)
            want_loss = tr.grads_only(xyz.cpu().numpy(), tgt.cpu().numpy(), n, acc_mode=0, grad_mode=0)
            want = tr.grads(); wm, wg = want[:m.n_mlp], want[m.n_mlp:]
            print(f"n = {n}: oracle {time.time() - t0:.1f} s; loss ours {loss:.6f} oracle {want_loss:.6f}", flush=True)
        e = np.abs(gm - wm)
        print(f"  flags {flags}: MLP grads: scale {np.abs(wm).max():.3e}, max err/scale {e.max() / np.abs(wm).max():.3e}, rel L2 {np.linalg.norm(gm - wm) / np.linalg.norm(wm):.3e}, "
              f"sign agreement {np.mean(np.sign(gm) == np.sign(wm)):.4f}")
        if flags == 0:
            eg = np.abs(gg - wg)
            print(f"  grid grads: rel L2 {np.linalg.norm(gg - wg) / np.linalg.norm(wg):.3e}, nonzero pattern match {np.mean((gg != 0) == (wg != 0)):.5f}, sign agreement among nonzero "
                  f"{np.mean(np.sign(gg[wg != 0]) == np.sign(wg[wg != 0])):.4f}")
            for l in range(m.L):
                a, b = int(m.offsets[l]) * m.F, int(m.offsets[l + 1]) * m.F
                print(f"    level {l}: |want| max {np.abs(wg[a:b]).max():.3e}  rel L2 err {np.linalg.norm(gg[a:b] - wg[a:b]) / max(np.linalg.norm(wg[a:b]), 1e-30):.3e}  sum got {gg[a:b].sum():+.4e} want {wg[a:b].sum():+.4e}")
        vol.optimizer_step(); torch.cuda.synchronize()      # consumes the gradients (state advances a little; the oracle gets the new blob)
        p16 = vol.get_params_f16(); p32 = O.f16_to_f32(p16)
