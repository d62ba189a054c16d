"""Where do data-parallel and accumulated training differ?  (GPU box; the set-up of tests/test_gpu_comm.py)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import instantvnr_b200 as vnr
from instantvnr_b200 import synthetic as syn
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_comm as T
DIMS, CFG = T.DIMS, T.CFG
world, n, steps = 2, 2048, 4
gt = syn.make_volume(DIMS, seed=3)


def make_volume(init):
    vol = vnr.NeuralVolume(vnr.model_json(**CFG), DIMS)
    vol.set_groundtruth(gt)
    if init:
        vol.init_params(11)
    return vol


def accumulated():
    ref = make_volume(True)
    xyz = torch.empty(n, 3, device="cuda"); tgt = torch.empty(n, device="cuda")
    for _ in range(steps):
        for _ in range(world):
            ref.sample(xyz, tgt, n)
            ref.train_grads(xyz, tgt, n, n * world)
        ref.optimizer_step()
    return ref.get_params_f16().view(np.float16).astype(np.float32), ref.n_mlp_params


a1, n_mlp = accumulated()
a2, _ = accumulated()
print("accumulated run vs itself: max", np.abs(a1 - a2).max(), "frac > 5e-3", np.mean(np.abs(a1 - a2) > 5e-3))
comms = vnr.Comm.init_local(world)
vols = []
for r, c in enumerate(comms):
    c.set_device(); vols.append(make_volume(r == 0))
for v, c in zip(vols, comms):
    v.attach_comm(c)
for _ in range(steps):
    for v in vols:
        v.train(1, batch=n, fast_mode=False)
p = vols[0].get_params_f16().view(np.float16).astype(np.float32)
d = np.abs(p - a1)
print("n_mlp", n_mlp, "of", p.size)
print("MLP : max", d[:n_mlp].max(), "frac > 5e-3", np.mean(d[:n_mlp] > 5e-3), "frac != ", np.mean(d[:n_mlp] != 0))
print("grid: max", d[n_mlp:].max(), "frac > 5e-3", np.mean(d[n_mlp:] > 5e-3), "frac != ", np.mean(d[n_mlp:] != 0))
for v in vols:
    v.detach_comm()
for c in comms:
    c.close()
