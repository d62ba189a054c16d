"""Small driver for ncu: build the bench scene (fewer training steps), render a few frames, run a few
training steps.  Used only under the profiler (numbers printed here are never bench values)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import instantvnr_b200 as vnr
from instantvnr_b200 import synthetic as syn
import bench

dims = (256, 256, 256)
LOG2 = int(os.environ.get("LOG2_HASHMAP", "19")); W = int(os.environ.get("FRAME_W", "1024")); H = int(os.environ.get("FRAME_H", "1024"))
vol, gt, (rgb, alpha) = bench.build_scene(vnr, dims, int(os.environ.get("TRAIN_STEPS", "200")), 1 << 16, dict(log2_hashmap=LOG2))
ren = vnr.Renderer(vol)
ren.set_size(W, H)
if os.environ.get("NO_DOWNLOAD") == "1":      # device-resident frames (what bench.py's `value` times): no PCIe stores in the kernels
    ren.set_download(False)
for v in range(int(os.environ.get("FRAMES", "3"))):
    ren.set_camera(*syn.default_camera(dims, v))
    ren.render()
    if os.environ.get("NO_DOWNLOAD") != "1":
        ren.map_frame()
    print("round_counts", v, ren.round_counts(), flush=True)
print(ren.stats())
vol.train(int(os.environ.get("EXTRA_TRAIN", "4")), batch=1 << 18, fast_mode=True)
print(vol.stats())
