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
