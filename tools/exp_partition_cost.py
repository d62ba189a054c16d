"""Experiment: per-rank frame time of the tile-parallel renderer without any cross-rank step (one GPU renders partition 0 of N),
and the floor of an all-miss frame.  Tells how much of the N-GPU frame time is compute vs fixed per-frame cost vs barrier."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import instantvnr_b200 as vnr
from instantvnr_b200 import synthetic as syn
import bench
dims = (256, 256, 256); W = H = 1024
vol, gt, (rgb, alpha) = bench.build_scene(vnr, dims, 300, 1 << 16)
cams = [syn.default_camera(dims, v, 16) for v in range(16)]
def fps(ren, cams, steps=256):
    ren.set_download(False)
    for i in range(16): ren.set_camera(*cams[i % len(cams)]); ren.render()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(steps): ren.set_camera(*cams[i % len(cams)]); ren.render()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3
for N in (1, 2, 4, 8):
    ren = vnr.Renderer(vol); ren.set_size(W, H); ren.set_partition(0, N)
    ms = fps(ren, cams)
    line = f"partition 0/{N}: {ms:.4f} ms/frame (graph loop, one frame at a time)"
    for depth in (2, 3, 4):
        ren.set_frames_in_flight(depth)
        line += f", {fps(ren, cams):.4f} with {depth} in flight"
    ren.set_frames_in_flight(1)
    ren.set_graph(False); ms_host = fps(ren, cams)
    print(line + f", {ms_host:.4f} host-enqueued rounds", flush=True)
away = [(np.array([0, 0, -600], np.float32), np.array([0, 0, -1200], np.float32), np.array([0, 1, 0], np.float32))]
ren = vnr.Renderer(vol); ren.set_size(W, H)
print(f"all rays miss: {fps(ren, away):.4f} ms/frame", flush=True)
ren = vnr.Renderer(vol); ren.set_size(W, H); ren.set_partition(0, 8)
print(f"all rays miss, partition 0/8: {fps(ren, away):.4f} ms/frame", flush=True)
