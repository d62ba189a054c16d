"""Generates tests/golden/tcnn_ref_*.npz on a GPU box from the reference's OWN tiny-cuda-nn build
(oracle/_ref/libvnr_tcnn_ref.so, compiled in place from /root/reference/tcnn): decode outputs and
training-step losses / parameters for seeded inputs.  The fixtures pin the CPU oracle (and through it
the product kernels) to the reference's arithmetic.  Run:  gpurun -- python tools/make_golden_tcnn.py
(writes into gpurun_out/golden/, copied to tests/golden/ by hand).  Inputs are regenerated from seeds
by the tests: parameters = oracle.init_params(model, seed) (verified identical to the tcnn Trainer's
own initialisation) with the grid part multiplied by `grid_scale`.
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import instantvnr_b200 as vnr
import oracle as O
from oracle import tcnn_ref
from instantvnr_b200 import synthetic as syn

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "golden")
os.makedirs(OUT, exist_ok=True)

CONFIGS = {
    "example": dict(n_levels=8, n_features=8, log2_hashmap=19, base_res=16, n_hidden=4),
    "variant": dict(n_levels=16, n_features=2, log2_hashmap=19, base_res=16, n_hidden=2),
    "small": dict(n_levels=4, n_features=4, log2_hashmap=12, base_res=8, n_hidden=2),
}


def params_for(m, seed, grid_scale):
    p32, _ = O.init_params(m, seed)
    p32 = p32.copy(); p32[m.n_mlp:] *= grid_scale
    return p32, O.f32_to_f16(p32)


def main():
    for name, cfg in CONFIGS.items():
        txt = vnr.model_json(**cfg)
        m = O.ModelCfg(cfg["n_levels"], cfg["n_features"], cfg["log2_hashmap"], cfg["base_res"], 2.0, cfg["n_hidden"])
        seed, grid_scale, n = 11, 2000.0, 4096
        _, p16 = params_for(m, seed, grid_scale)
        ref = tcnn_ref.RefNetwork(txt, seed)
        init_ok = bool(np.array_equal(ref.get_params_f16(), O.init_params(m, seed)[1]))
        ref.set_params_f16(p16)
        rng = np.random.default_rng(123)
        xyz = rng.random((n, 3), dtype=np.float32)
        d_xyz = torch.from_numpy(xyz).cuda(); d_out = torch.empty(n, device="cuda")
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            ref.inference(d_xyz, d_out, n, st.cuda_stream)
        st.synchronize()
        dec = d_out.cpu().numpy()
        # ---- training: 8 steps of Trainer::training_step on seeded batches of a synthetic volume
        dims = (32, 32, 32)
        vol = syn.make_volume(dims, seed=5)
        batch, steps = 1 << 12, 8
        reft = tcnn_ref.RefNetwork(txt, seed)          # fresh trainer, its own init (== oracle init)
        srng = O.Rng(1337)
        losses = []
        with torch.cuda.stream(st):
            for s in range(steps):
                c, t = O.sample_batch(srng, batch, vol, dims)
                dc = torch.from_numpy(c).cuda(); dt = torch.from_numpy(t).cuda()
                losses.append(reft.training_step(dc, dt, batch, st.cuda_stream, want_loss=True))
        st.synchronize()
        pt = reft.get_params_f16()
        idx = np.concatenate([np.arange(min(m.n_mlp, 4096)), m.n_mlp + rng.choice(m.n_grid, 8192, replace=False)]).astype(np.int64)
        np.savez_compressed(os.path.join(OUT, f"tcnn_ref_{name}.npz"), cfg=json.dumps(cfg), seed=seed, grid_scale=grid_scale, xyz=xyz, decode=dec,
                            init_identical=init_ok, train_dims=np.array(dims), train_vol_seed=5, train_batch=batch, train_losses=np.array(losses, np.float32),
                            train_param_idx=idx, train_params_f16=pt[idx], train_mlp_f16=pt[:m.n_mlp])
        print(name, "init identical:", init_ok, "decode range", float(dec.min()), float(dec.max()), "losses", losses)

main()
