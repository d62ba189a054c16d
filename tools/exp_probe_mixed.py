import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import instantvnr_b200 as vnr
TAB19 = 2920448 * 16
ops = (1 << 18) * 64
l, _ = vnr.probe_memory("loads", TAB19, ops, 8)
r, _ = vnr.probe_memory("reds", TAB19, ops, 8)
m, _ = vnr.probe_memory("mixed", TAB19, ops, 8)
print(f"2^18 x 64 ops over 46.7 MB tables: loads alone {l*1e3:.1f} us, reductions alone {r*1e3:.1f} us, both in one launch (loads over one table, reductions over another) {m*1e3:.1f} us")
ops = (1 << 22) * 64
l, _ = vnr.probe_memory("loads", TAB19, ops, 4)
r, _ = vnr.probe_memory("reds", TAB19, ops, 4)
m, _ = vnr.probe_memory("mixed", TAB19, ops, 4)
print(f"2^22 x 64 ops: loads {l*1e3:.1f} us, reductions {r*1e3:.1f} us, mixed {m*1e3:.1f} us")
