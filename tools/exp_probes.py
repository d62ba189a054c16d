"""Memory-system probes (GPU box): random 16-byte vs 32-byte loads (is the L1 bound per request?), stores into pinned host memory by
pattern (what bounds the zero-copy frame download)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import instantvnr_b200 as vnr          # noqa: E402

TAB19, TAB22 = 2920448 * 16, 19173376 * 16
ops = (1 << 22) * 64
for name, tab in (("46.7 MB", TAB19), ("306.8 MB", TAB22)):
    b16, _ = vnr.probe_memory("loads", tab, ops, 5)
    b32, _ = vnr.probe_memory("loads32", tab, ops, 5)
    print(f"table {name}: 16-byte loads {ops / b16 / 1e6:.1f} G/s ({ops * 16 / b16 / 1e6:.0f} GB/s);  32-byte loads {ops / b32 / 1e6:.1f} G/s ({ops * 32 / b32 / 1e6:.0f} GB/s)")
frame = 1024 * 1024 * 16
for kind in ("host_scanline", "host_tiles8x4", "host_scanline32", "host_tiles16x2"):
    b, m = vnr.probe_memory(kind, frame, 0, 8)
    print(f"{kind}: 16.8 MB frame into pinned host memory in {b:.3f} ms best / {m:.3f} mean = {frame / b / 1e6:.1f} GB/s")
