"""MLP weight gradients with half accumulators (the default) against the oracle's restatement (grad_mode 2) and against exact sums."""
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
m = O.ModelCfg()
vol = vnr.NeuralVolume(vnr.example_model_json(), DIMS)
vol.set_groundtruth_device(gt); vol.init_params(1337)
vol.train(40, batch=1 << 16, fast_mode=True)
p32 = O.f16_to_f32(vol.get_params_f16())
for n in (2560, 1 << 14, 1 << 16, 1 << 18):
    xyz = torch.empty(n, 3, device="cuda"); tgt = torch.empty(n, device="cuda")
    vol.sample(xyz, tgt, n); torch.cuda.synchronize()
    got = {}
    for flags in (0, 64):
        vol.train_debug(1, flags, False)
        vol.train_grads(xyz, tgt, n, n); torch.cuda.synchronize()
        got[flags] = vol.get_grads()[0].copy()
        # clear the accumulated grid gradients without moving the parameters: not possible through the API -> re-create the state
        vol.optimizer_step(); vol.set_params_f16(O.f32_to_f16(p32)); torch.cuda.synchronize()
    tr = O.Trainer(m, p32); tr.set_wgrad_slices(148)
    c, t = xyz.cpu().numpy(), tgt.cpu().numpy()
    tr.grads_only(c, t, n, 0, 0); exact = tr.grads()[:m.n_mlp].copy()
    t0 = time.time(); tr.grads_only(c, t, n, 0, 2); half = tr.grads()[:m.n_mlp].copy()
    sc = np.abs(exact).max()
    h = got[0]
    print(f"n = {n} (oracle half mode {time.time() - t0:.1f} s): |g| max {sc:.3e}")
    print(f"   GPU half vs oracle half: identical {np.mean(h == half):.4f}, max err/scale {np.abs(h - half).max() / sc:.2e}, rel L2 {np.linalg.norm(h - half) / np.linalg.norm(half):.2e}")
    print(f"   GPU half vs exact:       max err/scale {np.abs(h - exact).max() / sc:.2e}, rel L2 {np.linalg.norm(h - exact) / np.linalg.norm(exact):.2e}, sign agreement {np.mean(np.sign(h) == np.sign(exact)):.4f}")
    print(f"   GPU fp32 vs exact:       max err/scale {np.abs(got[64] - exact).max() / sc:.2e}, rel L2 {np.linalg.norm(got[64] - exact) / np.linalg.norm(exact):.2e}")
    print(f"   oracle half vs exact:    rel L2 {np.linalg.norm(half - exact) / np.linalg.norm(exact):.2e}, sign agreement {np.mean(np.sign(half) == np.sign(exact)):.4f}")
