"""Seeded synthetic inputs: volume, transfer function, camera.  numpy only.

The reference ships no sample data (its data/ directory lives in the un-vendored OVR repo), so
tests and benchmarks use these generators: a float32 volume in [0,1] made of Gaussian blobs plus
a low-frequency sine (empty space for macrocell skipping, smooth features that train to > 30 dB),
a 256-entry transfer function whose alpha is zero below 0.25, and the reference's camera
convention (world-space box of size `dims` centred at the origin, fovy 60).
"""
import numpy as np


def make_volume(dims, seed=42, n_blobs=8):
    dx, dy, dz = (int(d) for d in dims)
    rng = np.random.RandomState(seed)
    centres = rng.uniform(0.2, 0.8, size=(n_blobs, 3)).astype(np.float32)
    sigmas = rng.uniform(0.05, 0.15, size=n_blobs).astype(np.float32)
    amps = rng.uniform(0.5, 1.0, size=n_blobs).astype(np.float32)
    x = ((np.arange(dx, dtype=np.float32) + 0.5) / dx)[None, None, :]
    y = ((np.arange(dy, dtype=np.float32) + 0.5) / dy)[None, :, None]
    vol = np.zeros((dz, dy, dx), dtype=np.float32)
    # slab by slab along z so a 1024^3 volume never needs more than one slab of temporaries
    for k in range(dz):
        z = np.float32((k + 0.5) / dz)
        s = np.zeros((1, dy, dx), dtype=np.float32)
        for c, sg, a in zip(centres, sigmas, amps):
            r2 = (x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2
            s += a * np.exp(-r2 / (2 * sg * sg))
        s += 0.05 * (np.sin(2 * np.pi * x) * np.sin(2 * np.pi * y) * np.sin(2 * np.pi * z) + 1.0)
        vol[k] = s[0]
    lo, hi = float(vol.min()), float(vol.max())
    vol = (vol - lo) / (hi - lo)
    return np.ascontiguousarray(vol, dtype=np.float32)   # [z][y][x], x fastest


def make_tfn(n=256):
    t = np.linspace(0.0, 1.0, n, dtype=np.float32)
    stops = np.array([0.0, 0.25, 0.5, 0.75, 1.0], dtype=np.float32)
    cols = np.array([[0.05, 0.05, 0.3], [0.1, 0.5, 0.8], [0.2, 0.8, 0.3], [0.9, 0.8, 0.2], [0.9, 0.2, 0.1]], dtype=np.float32)
    rgb = np.stack([np.interp(t, stops, cols[:, c]) for c in range(3)], axis=1).astype(np.float32)
    u = np.clip((t - 0.25) / 0.75, 0.0, 1.0)
    alpha = (0.8 * u * u * (3 - 2 * u)).astype(np.float32)   # 0 below 0.25, smooth ramp to 0.8
    return rgb, alpha


def default_camera(dims, view=0, n_views=16):
    """from=(0,0,-1.5*dim) orbiting about the y axis; at=origin; up=+y (instantvnr_types.h:74-83)."""
    d = float(max(dims))
    ang = 2 * np.pi * view / n_views
    cam_from = np.array([1.5 * d * np.sin(ang), 0.0, -1.5 * d * np.cos(ang)], dtype=np.float32)
    return cam_from, np.zeros(3, np.float32), np.array([0, 1, 0], np.float32)


def psnr(a, b, peak=1.0):
    mse = float(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2))
    if mse == 0:
        return float("inf")
    return 10.0 * np.log10(peak * peak / mse)
