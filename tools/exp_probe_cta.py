"""Random 16-byte loads from ONE 256-thread CTA per SM: rate against the CTA's dynamic shared memory, i.e. against the
shared-memory / L1 split the driver picks for it (GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import instantvnr_b200 as vnr
TAB = 2920448 * 16
ops = (1 << 20) * 64
for kb in (0, 4, 8, 12, 16, 24, 32, 48, 64, 80, 96, 100, 104, 116, 128, 132, 136, 148, 160, 164, 168, 180, 192, 196, 200, 208, 216, 224, 227):
    best, mean = vnr.probe_memory(100 + kb, TAB, ops, 5)
    print(f"{kb:3d} KB smem: {ops / (best * 1e-3) / 1e9:7.1f} Gops/s  ({ops / (best * 1e-3) / 148 / 1.965e9:.3f} per cycle per SM)", flush=True)
