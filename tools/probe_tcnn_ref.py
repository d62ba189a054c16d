"""Does the reference's tcnn build run on this GPU?  parity vs our decode + throughput."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import instantvnr_b200 as vnr
import oracle as O
from oracle import tcnn_ref

def main():
    txt = vnr.example_model_json()
    ref = tcnn_ref.RefNetwork(txt, 1337)
    m = O.ModelCfg()
    print("ref n_params", ref.n_params, "ours", m.n_params)
    p32, _ = O.init_params(m, 7); p32[m.n_mlp:] *= 2000; p16 = O.f32_to_f16(p32)
    ref.set_params_f16(p16)
    assert np.array_equal(ref.get_params_f16(), p16)
    vol = vnr.NeuralVolume(txt, (256, 256, 256)); vol.set_params_f16(p16)
    n = 1 << 20
    torch.manual_seed(0)
    xyz = torch.rand(n, 3, device="cuda")
    a = torch.empty(n, device="cuda"); b = torch.empty(n, device="cuda")
    ref.inference(xyz, a, n); vol.decode(xyz, b, n, 0); torch.cuda.synchronize()
    a_h, b_h = a.cpu().numpy(), b.cpu().numpy()
    o0 = O.decode(m, p16, xyz[:8192].cpu().numpy(), 0); o1 = O.decode(m, p16, xyz[:8192].cpu().numpy(), 1)
    print("max|ours - tcnn|", np.abs(a_h - b_h).max(), " max|oracle(fp32acc) - tcnn|", np.abs(o0 - a_h[:8192]).max(),
          " max|oracle(fp16acc) - tcnn|", np.abs(o1 - a_h[:8192]).max(), " exact matches oracle1/tcnn", (o1 == a_h[:8192]).mean(), " value scale", np.abs(a_h).max())
    # init params parity: fresh trainer with the same seed
    ref2 = tcnn_ref.RefNetwork(txt, 1337)
    _, q16 = O.init_params(m, 1337)
    r16 = ref2.get_params_f16()
    print("init params identical to oracle:", np.array_equal(r16, q16), " mismatches:", int((r16 != q16).sum()))
    for log2n in (16, 20, 24):
        n = 1 << log2n
        xyz = torch.rand(n, 3, device="cuda"); out = torch.empty(n, device="cuda")
        for f, name in ((lambda: ref.inference(xyz, out, n), "tcnn"), (lambda: vol.decode(xyz, out, n, 0), "ours")):
            for _ in range(3): f()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): f()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"  n=2^{log2n} {name}: {n/ms/1e6:.3f} Gsamples/s ({ms:.3f} ms)")
    # training step throughput of the reference
    for log2n in (16, 18):
        n = 1 << log2n
        xyz = torch.rand(n, 3, device="cuda"); tgt = torch.rand(n, device="cuda")
        for _ in range(5): ref2.training_step(xyz, tgt, n, 0, want_loss=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): ref2.training_step(xyz, tgt, n, 0, want_loss=False)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"  tcnn training_step batch 2^{log2n}: {ms:.3f} ms/step = {1000/ms:.1f} steps/s, loss {ref2.training_step(xyz, tgt, n, 0):.5f}")
main()
