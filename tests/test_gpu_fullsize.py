"""Parity at BASELINE size, directly against the reference's own tiny-cuda-nn build on the same GPU
(oracle/_ref/libvnr_tcnn_ref.so, compiled unmodified from the reference tree by oracle/ref_driver/Makefile):

  * BASELINE configs[2]: training steps of 2^18 samples, example-model.json (T = 2^19), 256^3 volume -- the same
    pre-drawn batches go through vnr_volume_train_on and through the reference's Trainer::training_step
    (tcnn trainer.h:211-247, fully_fused_mlp.cu:819-943, encodings/grid.h:288-411) side by side;
  * decode of 2^20 coordinates: vnr_volume_decode against NetworkWithInputEncoding::inference (ref_inference).

Tolerances are written where they are asserted.  What can and cannot agree at this size:
  * one batch's gradients are deterministic up to the order of the fp16 hash-grid reductions and are checked against the CPU
    oracle at full size (MLP: the half-accumulated sums the reference's split-K GEMMs produce, restated by the oracle; grid:
    statistically);
  * the reference accumulates its MLP weight gradients in HALF (cutlass_matmul.h:83): at 2^18 samples that decides the sign of
    about a fifth of them, so the product accumulates in half too (train.cu) -- first-step update directions then agree > 99 %;
  * training itself is chaotic at this batch size: the per-sample gradient 128 / 2^18 puts every activation gradient into the
    fp16 subnormal range, Adam turns rounding noise into +-lr steps, and two runs of the REFERENCE on the same batches differ
    by several dB of volume PSNR after 300 steps (measured here, printed).  The long-run check is therefore a band, next to the
    reference's own run-to-run spread, not a 0.1 dB identity.
"""
import numpy as np
import pytest
import torch

import instantvnr_b200 as vnr
from instantvnr_b200 import synthetic as syn
from oracle import tcnn_ref

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not tcnn_ref.available(), reason="oracle/_ref/libvnr_tcnn_ref.so not built")]

DIMS = (256, 256, 256)
N = 1 << 18


def _synth_device(dims):
    import bench
    return bench.synth_volume_device(dims)


def _pair(seed=1337):
    vol = vnr.NeuralVolume(vnr.example_model_json(), DIMS)
    vol.set_groundtruth_device(_synth_device(DIMS))
    vol.init_params(seed)
    ref = tcnn_ref.RefNetwork(vnr.example_model_json(), seed)
    # Trainer::initialize_params is restated bit-exactly (golden fixtures): both start from the same blob
    assert np.array_equal(vol.get_params_f16(), ref.get_params_f16())
    return vol, ref


def _ref_psnr(ref, gt, st):
    """10 log10(range^2 / mse) of the reference network over all voxel centres (network.cu:410-472)"""
    dz, dy, dx = gt.shape
    x = (torch.arange(dx, device="cuda", dtype=torch.float32) + 0.5) / dx
    y = (torch.arange(dy, device="cuda", dtype=torch.float32) + 0.5) / dy
    se = torch.zeros((), device="cuda", dtype=torch.float64)
    slab = 16
    out = torch.empty(slab * dy * dx, device="cuda")
    for z0 in range(0, dz, slab):
        z = (torch.arange(z0, z0 + slab, device="cuda", dtype=torch.float32) + 0.5) / dz
        zz, yy, xx = torch.meshgrid(z, y, x, indexing="ij")
        xyz = torch.stack([xx, yy, zz], -1).reshape(-1, 3).contiguous()
        with torch.cuda.stream(st):
            ref.inference(xyz, out, xyz.shape[0], st.cuda_stream)
        st.synchronize()
        se += ((out.view(slab, dy, dx) - gt[z0:z0 + slab]).double() ** 2).sum()
    rng = float(gt.max() - gt.min())
    return 10.0 * np.log10(rng * rng / (float(se) / gt.numel()))


def test_gradients_of_one_baseline_batch_match_the_oracle():
    """2^18 samples, T = 2^19, 256^3: every CTA runs 13-14 tiles through the rings of the fused kernel (the small-config tests
    give each CTA one tile), gradients against the CPU oracle."""
    import oracle as O
    O.use_host_cores()
    vol = vnr.NeuralVolume(vnr.example_model_json(), DIMS)
    vol.set_groundtruth_device(_synth_device(DIMS))
    vol.init_params(1337)
    vol.train(40, batch=1 << 16, fast_mode=True)                     # a table that matters
    p16 = vol.get_params_f16()
    m = O.ModelCfg()
    xyz = torch.empty(N, 3, device="cuda"); tgt = torch.empty(N, device="cuda")
    vol.sample(xyz, tgt, N); torch.cuda.synchronize()
    c, t = xyz.cpu().numpy(), tgt.cpu().numpy()
    tr = O.Trainer(m, O.f16_to_f32(p16)); tr.set_wgrad_slices(148)
    want_loss = tr.grads_only(c, t, N, 0, 0)
    exact = tr.grads().copy()
    tr.grads_only(c, t, N, 0, 2)
    half = tr.grads()[:m.n_mlp].copy()
    # default: half-accumulated weight gradients
    vol.train_grads(xyz, tgt, N, N); torch.cuda.synchronize()
    gm, gg16 = vol.get_grads(); gg = O.f16_to_f32(gg16)
    assert abs(vol.last_loss() - want_loss) <= 1e-5 * want_loss
    sc = np.abs(exact[:m.n_mlp]).max()
    print(f"MLP (half accumulators) vs oracle restatement: identical {np.mean(gm == half):.4f}, rel L2 {np.linalg.norm(gm - half) / np.linalg.norm(half):.2e}; "
          f"vs exact sums rel L2 {np.linalg.norm(gm - exact[:m.n_mlp]) / np.linalg.norm(exact[:m.n_mlp]):.2e}")
    assert np.mean(gm == half) >= 0.75 and np.linalg.norm(gm - half) <= 1e-3 * np.linalg.norm(half)
    assert np.abs(gm - exact[:m.n_mlp]).max() <= 1e-2 * sc
    # hash-grid gradients: fp16 reductions in arrival order vs the sequential float sum of the same fp16 addends
    wg = exact[m.n_mlp:]
    print(f"grid: rel L2 {np.linalg.norm(gg - wg) / np.linalg.norm(wg):.2e}, touched-set match {np.mean((gg != 0) == (wg != 0)):.5f}")
    assert np.linalg.norm(gg - wg) <= 2e-2 * np.linalg.norm(wg)
    assert np.mean((gg != 0) == (wg != 0)) >= 0.999
    nz = wg != 0
    assert np.mean(np.sign(gg[nz]) == np.sign(wg[nz])) >= 0.995
    for l in range(m.L):
        a, b = int(m.offsets[l]) * m.F, int(m.offsets[l + 1]) * m.F
        assert abs(gg[a:b].sum() - wg[a:b].sum()) <= 1e-2 * np.abs(wg[a:b]).sum()
    # fp32 accumulators (train flag 64): the exact sums
    vol.optimizer_step(); vol.set_params_f16(p16)
    vol.train_debug(2, 64, False)
    vol.train_grads(xyz, tgt, N, N); torch.cuda.synchronize()
    gm32, _ = vol.get_grads()
    # fp32 tensor-core accumulation of 2^18 terms against double sums: 1e-6 .. 2e-4 of the scale, depending on how much the trained
    # state makes the terms of the output row cancel (tools/exp_var2_fp32.py); the half-accumulated default is 1e-3 .. 1e-2
    assert np.abs(gm32 - exact[:m.n_mlp]).max() <= 1e-3 * sc


def test_training_steps_match_reference_tcnn_at_baseline_size():
    vol, ref = _pair()
    ref2 = tcnn_ref.RefNetwork(vnr.example_model_json(), 1337)      # the reference against itself: its run-to-run spread
    st = torch.cuda.Stream()
    xyz = torch.empty(N, 3, device="cuda"); tgt = torch.empty(N, device="cuda")
    p0 = vol.get_params_f16().view(np.float16).astype(np.float32)
    n_mlp = vol.n_mlp_params
    ours, theirs, theirs2 = [], [], []
    first = None
    steps = 8
    for i in range(steps):
        vol.sample(xyz, tgt, N)                      # StaticSampler stream of the product; all arms consume the same batch
        torch.cuda.synchronize()
        vol.train_on(xyz, tgt, N)
        ours.append(vol.last_loss())
        with torch.cuda.stream(st):
            theirs.append(ref.training_step(xyz, tgt, N, st.cuda_stream, want_loss=True))
            theirs2.append(ref2.training_step(xyz, tgt, N, st.cuda_stream, want_loss=True))
        st.synchronize()
        if i == 0:
            first = (vol.get_params_f16().view(np.float16).astype(np.float32) - p0, ref.get_params_f16().view(np.float16).astype(np.float32) - p0)
    ours, theirs, theirs2 = np.array(ours), np.array(theirs), np.array(theirs2)
    print("loss ours       ", np.round(ours, 5))
    print("loss reference  ", np.round(theirs, 5))
    print("loss reference#2", np.round(theirs2, 5), " (same batches: the reference's own run-to-run spread)")
    # the first step: Adam moves every parameter with a non-zero gradient by exactly +-lr, so the update compares the SIGN and the
    # zero pattern of every single gradient
    a, b = first
    for name, sl, same_dir in (("mlp", slice(0, n_mlp), 0.99), ("grid", slice(n_mlp, None), 0.98)):
        x, y = a[sl], b[sl]
        both = (x != 0) & (y != 0)
        agree = float(np.mean(np.sign(x[both]) == np.sign(y[both])))
        same_set = float(np.mean((x != 0) == (y != 0)))
        print(f"first step {name}: moved ours {np.mean(x != 0):.4f} reference {np.mean(y != 0):.4f}, same set {same_set:.4f}, same direction {agree:.4f}")
        assert same_set >= 0.99 and agree >= same_dir
    # per-step loss: identical start, then within 5 % over 8 steps (measured <= 2.7 %; the reference against itself on the same
    # batches differs by up to ~1 % through the order of its atomics alone)
    assert abs(ours[0] - theirs[0]) <= 1e-5 * theirs[0]
    assert np.all(np.abs(ours - theirs) <= 0.05 * theirs), (ours, theirs)


def test_volume_psnr_after_300_steps_is_in_the_references_band():
    vol, ref = _pair()
    ref2 = tcnn_ref.RefNetwork(vnr.example_model_json(), 1337)
    gt = _synth_device(DIMS)
    st = torch.cuda.Stream()
    xyz = torch.empty(N, 3, device="cuda"); tgt = torch.empty(N, device="cuda")
    lo, lr_ = [], []
    p_ours, p_ref, p_ref2 = [], [], []
    for i in range(300):
        vol.sample(xyz, tgt, N)                      # fresh batches (a short ring of batches over-fits: the PSNR then measures chaos)
        torch.cuda.synchronize()
        vol.train_on(xyz, tgt, N)
        with torch.cuda.stream(st):
            l = ref.training_step(xyz, tgt, N, st.cuda_stream, want_loss=i >= 200)
            ref2.training_step(xyz, tgt, N, st.cuda_stream, want_loss=False)
        st.synchronize()
        if i >= 200:
            lo.append(vol.last_loss()); lr_.append(l)
        if i >= 250 and i % 10 == 9:                 # the whole-volume PSNR at five checkpoints: its median is what is compared
            torch.cuda.synchronize()
            p_ours.append(vol.psnr()); p_ref.append(_ref_psnr(ref, gt, st)); p_ref2.append(_ref_psnr(ref2, gt, st))
    torch.cuda.synchronize()
    psnr_ours, psnr_ref, psnr_ref2 = float(np.median(p_ours)), float(np.median(p_ref)), float(np.median(p_ref2))
    print(f"volume PSNR over steps 260-300 of 2^18 fresh samples (median of 5): ours {psnr_ours:.2f} dB {np.round(p_ours, 1)}, reference {psnr_ref:.2f} dB "
          f"{np.round(p_ref, 1)}, reference again {psnr_ref2:.2f} dB; mean L1 loss of steps 200-300: ours {np.mean(lo):.5f}, reference {np.mean(lr_):.5f}")
    # the whole-volume PSNR of either trainer swings by several dB from step to step at this batch size (see the module note; the
    # reference against itself differs by 3 - 10 dB at a single step), hence the median over five checkpoints.  The band: both have
    # learnt the volume (a constant predictor scores 17 dB) and the smoothed training loss agrees
    assert min(psnr_ours, psnr_ref) >= 38.0
    assert abs(psnr_ours - psnr_ref) <= 12.0
    assert 0.6 <= np.mean(lo) / np.mean(lr_) <= 1.6


def test_decode_2p20_matches_reference_inference():
    vol, ref = _pair()
    # make the table matter: a few training steps on both would diverge the blobs, so train ours and copy the blob over
    vol.train(60, batch=1 << 16, fast_mode=True)
    p16 = vol.get_params_f16()
    ref.set_params_f16(p16)
    n = 1 << 20
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    xyz = torch.rand(n, 3, device="cuda", generator=g)
    out = torch.empty(n, device="cuda"); want = torch.empty(n, device="cuda")
    vol.decode(xyz, out, n)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        ref.inference(xyz, want, n, st.cuda_stream)
    st.synchronize(); torch.cuda.synchronize()
    d = (out - want).abs()
    tol = 4 * 2.0 ** -9 * torch.clamp(want.abs(), min=1.0)       # fp16-accumulating reference vs fp32 accumulators: 4 half ulps at 1
    print(f"decode 2^20 vs reference inference: max |d| {float(d.max()):.3e}, mean {float(d.mean()):.3e}, value range [{float(want.min()):.3f}, {float(want.max()):.3f}]")
    assert bool((d <= tol).all())
    assert float(d.mean()) <= 2.0 ** -11
