"""Quick decode micro-benchmark (device-resident coords), CUDA-event timed."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import instantvnr_b200 as vnr

def main():
    log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    n = 1 << log2n
    for name, cfg in [("example 8x8 T19 h4", dict()), ("variant 16x2 T19 h2", dict(n_levels=16, n_features=2, n_hidden=2)),
                      ("example T22", dict(log2_hashmap=22))]:
        vol = vnr.NeuralVolume(vnr.model_json(**cfg), (256, 256, 256))
        vol.init_params(1337)
        torch.manual_seed(0)
        xyz = torch.rand(n, 3, device="cuda", dtype=torch.float32)
        out = torch.empty(n, device="cuda", dtype=torch.float32)
        s = torch.cuda.current_stream().cuda_stream
        for _ in range(3):
            vol.decode(xyz, out, n, s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            vol.decode(xyz, out, n, s)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        # sorted (coherent) coordinates: Morton-ish via sorting on a coarse cell key
        key = (xyz[:, 2] * 64).long() * 4096 + (xyz[:, 1] * 64).long() * 64 + (xyz[:, 0] * 64).long()
        xs = xyz[torch.argsort(key)].contiguous()
        for _ in range(2):
            vol.decode(xs, out, n, s)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            vol.decode(xs, out, n, s)
        e1.record(); torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / reps
        fold = torch.empty(n, device="cuda", dtype=torch.int32)
        res = []
        for c in (xyz, xs):
            for _ in range(2):
                vol.gather_probe(c, fold, n, s)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                vol.gather_probe(c, fold, n, s)
            e1.record(); torch.cuda.synchronize()
            res.append(n / (e0.elapsed_time(e1) / reps) / 1e6)
        print(f"{name}: gather-only probe random {res[0]:.3f} | cell-sorted {res[1]:.3f} Gsamples/s")
        print(f"{name}: random {n/ms/1e6:.3f} Gsamples/s ({ms:.3f} ms) | cell-sorted {n/ms2/1e6:.3f} Gsamples/s ({ms2:.3f} ms)", flush=True)

main()
