"""Experiment: consecutive frames on alternating renderers (own streams and ray buffers) of one volume: does the tail of frame i
(small latency-bound rounds) overlap the head of frame i+1?  Device-resident frames, 16-view orbit."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import instantvnr_b200 as vnr
from instantvnr_b200 import synthetic as syn
import bench
dims = (256, 256, 256); W = H = 1024
vol, gt, (rgb, alpha) = bench.build_scene(vnr, dims, 300, 1 << 16)
cams = [syn.default_camera(dims, v, 16) for v in range(16)]
for K in (1, 2, 3):
    rens = []
    for _ in range(K):
        r = vnr.Renderer(vol); r.set_size(W, H); r.set_download(False); rens.append(r)
    for i in range(32):
        r = rens[i % K]; r.set_camera(*cams[i % 16]); r.render()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    steps = 512
    for i in range(steps):
        r = rens[i % K]; r.set_camera(*cams[i % 16]); r.render()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{K} alternating renderer(s): {steps / dt:.1f} fps ({dt / steps * 1e3:.4f} ms/frame)", flush=True)
    del rens

# end to end with one frame of lookahead: frame i+1 is launched on the other renderer before frame i is mapped; every frame is
# mapped (synchronised, host-visible) exactly once inside the timed region
import numpy as np
for K, zc in ((1, True), (2, True), (3, True), (1, False), (2, False), (3, False)):
    rens = []
    for _ in range(K):
        r = vnr.Renderer(vol); r.set_size(W, H); r.set_download(True); r.set_zero_copy(zc); rens.append(r)
    def run(steps):
        chk = 0.0
        for i in range(steps + K - 1):
            if i < steps:
                r = rens[i % K]; r.set_camera(*cams[i % 16]); r.render()
            j = i - (K - 1)
            if j >= 0:
                img = rens[j % K].map_frame(copy=False)
                chk += float(img[H // 2, W // 2, 3])
        return chk
    run(32)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    steps = 512
    run(steps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"e2e, {K} renderer(s), lookahead {K - 1}, {'zero-copy stores' if zc else 'DMA copy after the frame'}: {steps / dt:.1f} fps ({dt / steps * 1e3:.4f} ms/frame)", flush=True)
    # correctness: the mapped frames equal the single-renderer frames
    single = vnr.Renderer(vol); single.set_size(W, H)
    single.set_camera(*cams[5]); single.render(); want = single.map_frame().copy()
    r = rens[0]; r.set_camera(*cams[5]); r.render()
    if K > 1:
        r2 = rens[1]; r2.set_camera(*cams[6]); r2.render()
    got = r.map_frame().copy()
    print("   frame identical to the single renderer:", np.array_equal(got, want), flush=True)
    del rens
