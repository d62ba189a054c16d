import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import instantvnr_b200 as vnr
import oracle as O
from instantvnr_b200 import synthetic as syn
CFGS = [
    dict(n_levels=4, n_features=8, log2_hashmap=12, base_res=8, n_hidden=2),
    dict(n_levels=8, n_features=8, log2_hashmap=14, base_res=16, n_hidden=4),
    dict(n_levels=16, n_features=2, log2_hashmap=12, base_res=4, n_hidden=2),
    dict(n_levels=8, n_features=4, log2_hashmap=12, base_res=4, n_hidden=3),
    dict(n_levels=16, n_features=1, log2_hashmap=12, base_res=4, n_hidden=1),
]
for cfg in CFGS:
    m = O.ModelCfg(cfg["n_levels"], cfg["n_features"], cfg["log2_hashmap"], cfg["base_res"], 2.0, cfg["n_hidden"])
    dims = (16, 16, 16); gt = syn.make_volume(dims, seed=3)
    p32, _ = O.init_params(m, 5); p32 = p32.copy(); p32[m.n_mlp:] *= 1000.0; p16 = O.f32_to_f16(p32)
    vol = vnr.NeuralVolume(vnr.model_json(**cfg), dims); vol.set_groundtruth(gt); vol.set_params_f16(p16)
    n = 128 * 20; rng = O.Rng(77); c, t = O.sample_batch(rng, n, gt, dims)
    dc, dt = torch.from_numpy(c).cuda(), torch.from_numpy(t).cuda()
    vol.train_grads(dc, dt, n, n, torch.cuda.current_stream().cuda_stream); torch.cuda.synchronize()
    gm, gg16 = vol.get_grads(); gg = O.f16_to_f32(gg16)
    tr = O.Trainer(m, O.f16_to_f32(p16)); loss = tr.step(c, t, acc_mode=0, grad_mode=0, do_step=False)
    want = tr.grads(); wm, wg = want[:m.n_mlp], want[m.n_mlp:]
    e = np.abs(gg - wg); i = int(e.argmax()); gs = np.abs(wg).max()
    lvl = np.searchsorted(m.offsets, i // m.F, side="right") - 1
    print(cfg, "loss", vol.last_loss(), loss)
    print("  mlp: max err/scale", np.abs(gm - wm).max() / np.abs(wm).max())
    print("  grid: scale", gs, "max err", e.max(), "at", i, "level", lvl, "want", wg[i], "got", gg[i], "| mean err/scale", e.mean() / gs, "| frac>1%", (e > 0.01 * gs).mean(), "| nonzero match", np.mean((gg != 0) == (wg != 0)))
    for l in range(m.L):
        a, b = int(m.offsets[l]) * m.F, int(m.offsets[l + 1]) * m.F
        print(f"    level {l}: scale {np.abs(wg[a:b]).max():.3e} max err {e[a:b].max():.3e} sum got {gg[a:b].sum():.4e} want {wg[a:b].sum():.4e}")
