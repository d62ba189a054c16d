#!/bin/bash
# what the driver runs at round end, on one GPU: reference arm, then our arm (default flags), for both workloads
mkdir -p gpurun_out
TAG=${1:-r02w}
timeout 900 python bench.py --impl reference > gpurun_out/bench_reference_1gpu_$TAG.json 2> gpurun_out/bench_reference_1gpu_$TAG.err; echo "reference rc=$?"
timeout 900 python bench.py > gpurun_out/bench_render_1gpu_$TAG.json 2> gpurun_out/bench_render_1gpu_$TAG.err; echo "ours rc=$?"
timeout 600 python bench.py --impl reference --workload train --steps 100 --warmup 16 > gpurun_out/bench_reference_train_1gpu_$TAG.json 2> gpurun_out/bench_reference_train_1gpu_$TAG.err; echo "reference train rc=$?"
timeout 600 python bench.py --workload train --steps 100 --warmup 16 > gpurun_out/bench_train_1gpu_$TAG.json 2> gpurun_out/bench_train_1gpu_$TAG.err; echo "train rc=$?"
python - <<PY
import json
def L(n): return json.load(open(f"gpurun_out/{n}_$TAG.json"))
r, o = L("bench_reference_1gpu"), L("bench_render_1gpu")
print("render: ours value", round(o["value"]/1e9,3), "G fps", round(o["fps"],1), "e2e", round(o["e2e"]["value"]/1e9,3), "G fps", round(o["e2e"]["fps"],1), "| reference value", round(r["value"]/1e9,3), "fps", round(r["fps"],1), "| ratio", round(o["value"]/r["value"],2), "e2e ratio", round(o["e2e"]["value"]/r["e2e"]["value"],2))
print("same workload string:", o["config"]["workload"] == r["config"]["workload"], "| roofline", o["roofline"]["bound"], o["roofline"]["frac"], "in-frame", o["roofline"]["in_frame"]["ratio_to_peak"], "| clocks", o["clocks"], "| cpu", o["cpu_baseline"]["value"], "| steps", o["steps"], "launches", o["gpu_launches"])
r, o = L("bench_reference_train_1gpu"), L("bench_train_1gpu")
print("train: ours", round(o["value"],1), "e2e", round(o["e2e"]["value"],1), "| reference", round(r["value"],1), "e2e", round(r["e2e"]["value"],1), "| ratio", round(o["value"]/r["value"],2), "same workload string:", o["config"]["workload"] == r["config"]["workload"], o["roofline"]["train_step_kernel"])
PY
