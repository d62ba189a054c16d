#!/bin/bash
# full single-GPU check: the GPU test suite, smoke(), the default bench line and the train workload line
mkdir -p gpurun_out
TAG=${1:-r02t}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --steps 2 --cpu-seconds 1 > gpurun_out/bench_render_1gpu_$TAG.json 2> gpurun_out/bench_render_1gpu_$TAG.err; echo "bench rc=$?"
timeout 300 python bench.py --workload train --steps 200 --warmup 20 > gpurun_out/bench_train_1gpu_$TAG.json 2> gpurun_out/bench_train_1gpu_$TAG.err; echo "bench train rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_render_1gpu_$TAG.json")); r=d["roofline"]
print("render: value", round(d["value"]/1e9,3), "fps", round(d["fps"],1), "e2e fps", round(d["e2e"]["fps"],1), "inflight", d["e2e"]["fps_with_frames_in_flight"], "roofline", r["bound"], r["achieved"], r["peak"], r["frac"], "in-frame", r["in_frame"]["ratio_to_peak"], "train", round(d["train_steps_per_sec_batch_2p18"],1), "cpu", d["cpu_baseline"]["value"])
d=json.load(open("gpurun_out/bench_train_1gpu_$TAG.json"))
print("train:", {k: d[k] for k in ("value","ms_per_step","mean_loss","last_loss","volume_psnr_db")}, d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["train_step_kernel"])
PY
