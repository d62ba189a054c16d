"""Why does the mc-merge barrier of the shared-device DP test time out?  (GPU box)"""
import os, sys, time
os.environ.setdefault("VNR_COMM_SHARE_DEVICES", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import instantvnr_b200 as vnr
from instantvnr_b200 import synthetic as syn

CFG = dict(n_levels=8, n_features=8, log2_hashmap=14, base_res=16, n_hidden=4)
DIMS = (48, 48, 48)
gt = syn.make_volume(DIMS, seed=3)


def run(fast):
    comms = vnr.Comm.init_local(2)
    vols = []
    for r, c in enumerate(comms):
        c.set_device()
        v = vnr.NeuralVolume(vnr.model_json(**CFG), DIMS); v.set_groundtruth(gt)
        if r == 0:
            v.init_params(11)
        vols.append(v)
    for v, c in zip(vols, comms):
        v.attach_comm(c)
    for step in range(3):
        for r, v in enumerate(vols):
            t0 = time.time(); v.train(1, batch=2048, fast_mode=fast); print(f"  fast={fast} step {step} rank {r}: train() returned after {time.time() - t0:.3f} s", flush=True)
        t0 = time.time()
        try:
            print("  loss", vols[0].last_loss(), vols[1].last_loss(), f"({time.time() - t0:.3f} s)", flush=True)
        except Exception as e:
            print("  ERR", e, f"({time.time() - t0:.3f} s)", flush=True)
    ps = [v.get_params_f16() for v in vols]
    print("  replicas identical:", np.array_equal(ps[0], ps[1]))
    for v in vols:
        v.detach_comm()
    for c in comms:
        c.close()


run(True)
run(False)
