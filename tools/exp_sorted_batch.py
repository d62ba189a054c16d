"""Does the fused training kernel get faster when the batch is ordered by coarse grid cell (requests of a warp coalesce on the
coarse levels, as they do inside a frame)?  The sample SET is the same, only its order changes.  (GPU box)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import instantvnr_b200 as vnr
import bench
torch.cuda.set_device(0)
n = 1 << 18
dims = (256,) * 3
gt = bench.synth_volume_device(dims)
vol = vnr.NeuralVolume(vnr.model_json(), dims)
vol.set_groundtruth_device(gt); vol.init_params(1337)
vol.train(50, batch=1 << 16, fast_mode=True)
st = torch.cuda.ExternalStream(vol.stream())
xyz = torch.empty(n, 3, device="cuda"); tgt = torch.empty(n, device="cuda")
vol.sample(xyz, tgt, n); torch.cuda.synchronize()


def part1by2(v):
    v = v & 0x3FF
    v = (v | (v << 16)) & 0x30000FF
    v = (v | (v << 8)) & 0x300F00F
    v = (v | (v << 4)) & 0x30C30C3
    v = (v | (v << 2)) & 0x9249249
    return v


def morton_order(xyz, res):
    c = (xyz * res).long().clamp_(0, res - 1)
    key = part1by2(c[:, 0]) | (part1by2(c[:, 1]) << 1) | (part1by2(c[:, 2]) << 2)
    return torch.argsort(key)


def linear_order(xyz, res):
    c = (xyz * res).long().clamp_(0, res - 1)
    return torch.argsort(c[:, 0] + res * (c[:, 1] + res * c[:, 2]))


def time_kernel(x, t, reps=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        vol.train_grads(x, t, n, n)
    e0.record(st)
    for _ in range(reps):
        vol.train_grads(x, t, n, n)
    e1.record(st); st.synchronize()
    vol.optimizer_step(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def per_cta_blocks(order, n_cta=148, tile=128):
    """re-deal a sorted batch so that the kernel's strided tile assignment (tile j of CTA b = b + j * n_cta) gives every CTA a
    CONTIGUOUS run of the sorted order (its own region of the volume)"""
    n_tiles = order.numel() // tile
    src = order.view(n_tiles, tile)
    dst = torch.empty_like(src)
    k = 0
    for b in range(n_cta):
        for t in range(b, n_tiles, n_cta):
            dst[t] = src[k]; k += 1
    return dst.reshape(-1)


orders = {"unsorted": None}
for res in (16, 64, 256):
    orders[f"morton {res}^3"] = morton_order(xyz, res)
    orders[f"morton {res}^3 per-CTA"] = per_cta_blocks(morton_order(xyz, res))
for res in (16, 64, 256):
    orders[f"linear {res}^3"] = linear_order(xyz, res)
    orders[f"linear {res}^3 per-CTA"] = per_cta_blocks(linear_order(xyz, res))
for flags, what in ((0, "all"), (1, "chain + gather"), (2, "chain + scatter")):
    for name, o in orders.items():
        x = xyz if o is None else xyz[o].contiguous(); t = tgt if o is None else tgt[o].contiguous()
        vol.train_debug(2, flags, False)
        print(f"{what:16s} {name:22s} {time_kernel(x, t):7.1f} us", flush=True)
vol.train_debug(2, 0, False)
