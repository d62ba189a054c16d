"""Experiment: one frame rendered as Q interleaved-strip sub-frames on Q streams (Q renderers on one GPU), all storing finished
pixels into ONE pinned host frame.  Measures the end-to-end frame rate (render + host-visible frame) against the single renderer."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import instantvnr_b200 as vnr
from instantvnr_b200 import synthetic as syn
import bench

dims = (256, 256, 256); W = H = 1024
vol, gt, (rgb, alpha) = bench.build_scene(vnr, dims, 300, 1 << 16)
cams = [syn.default_camera(dims, v, 16) for v in range(16)]
steps = 128

base = vnr.Renderer(vol); base.set_size(W, H)
def run_base(download):
    base.set_download(download)
    for i in range(8):
        base.set_camera(*cams[i % 16]); base.render()
        if download: base.map_frame(copy=False)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(steps):
        base.set_camera(*cams[i % 16]); base.render()
        if download: base.map_frame(copy=False)
    torch.cuda.synchronize()
    return steps / (time.perf_counter() - t0)
print("single renderer: device-resident fps %.1f, e2e fps %.1f" % (run_base(False), run_base(True)), flush=True)
base.set_camera(*cams[3]); base.set_download(True); base.render(); ref = base.map_frame().copy()

for Q in (2, 3, 4):
    host = torch.zeros(H, W, 4, dtype=torch.float32).pin_memory()
    dev = torch.zeros(H, W, 4, dtype=torch.float32, device="cuda")
    rens = []
    for q in range(Q):
        r = vnr.Renderer(vol); r.set_size(W, H); r.set_partition(q, Q); r.set_download(False)
        rens.append(r)
    for name, target in (("device frame", dev.data_ptr()), ("pinned host frame", host.data_ptr())):
        for r in rens: r.set_frame_target(target)
        def frame(i):
            for r in rens:
                r.set_camera(*cams[i % 16]); r.render()
            torch.cuda.synchronize()
        for i in range(8): frame(i)
        t0 = time.perf_counter()
        for i in range(steps): frame(i)
        fps = steps / (time.perf_counter() - t0)
        frame(3)
        got = host.numpy() if target == host.data_ptr() else dev.cpu().numpy()
        print(f"Q={Q} {name}: fps {fps:.1f}  identical to single-renderer frame: {np.array_equal(got, ref)}", flush=True)
    for r in rens: r.set_frame_target(None)
    del rens
