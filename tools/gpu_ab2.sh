#!/bin/bash
mkdir -p gpurun_out
for cfg in "1 1" "0 0"; do
  set -- $cfg
  VNR_RM_TILED=$1 VNR_RM_TRANSPOSE=$2 timeout 300 python bench.py --cpu-seconds 1 --width 3840 --height 2160 --log2-hashmap 22 --steps 32 --warmup 4 --train-steps 300 > gpurun_out/ab4k_t$1_x$2.json 2> gpurun_out/ab4k_t$1_x$2.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab4k_t$1_x$2.json"))
print("4K T22 tiled=$1 transpose=$2: fps", round(d["fps"],1), "value", round(d["value"]/1e9,3), "G/s e2e fps", round(d["e2e"]["fps"],1), "decode G/s", round(d["roofline"]["decode_samples_per_sec"]/1e9,3), "ms/frame", round(d["ms_per_step"],4))
PY
done
timeout 300 python bench.py --workload interleaved --steps 100 --warmup 5 > gpurun_out/bench_interleaved_1gpu_r01g.json 2> gpurun_out/bench_interleaved_1gpu_r01g.err; echo "interleaved rc=$?"; cut -c1-900 gpurun_out/bench_interleaved_1gpu_r01g.json
timeout 300 python bench.py > gpurun_out/bench_render_1gpu_r01g.json 2> gpurun_out/bench_render_1gpu_r01g.err; echo "bench rc=$?"
