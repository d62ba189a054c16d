#!/bin/bash
# Quick GPU-box pass: parity suite + both default bench arms.   gpurun --timeout 1500 -- bash tools/gpu_check.sh r01f
TAG=${1:-r01}
mkdir -p gpurun_out
t0=$(date +%s)
timeout 1200 python -m pytest tests -m gpu -q ${PYTEST_ARGS:--x} > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - t0 )) s)"; tail -5 gpurun_out/pytest_gpu_$TAG.log
t0=$(date +%s)
timeout 400 python bench.py > gpurun_out/bench_render_1gpu_$TAG.json 2> gpurun_out/bench_render_1gpu_$TAG.err; echo "bench rc=$? ($(( $(date +%s) - t0 )) s)"
timeout 300 python bench.py --workload train > gpurun_out/bench_train_1gpu_$TAG.json 2> gpurun_out/bench_train_1gpu_$TAG.err; echo "train rc=$?"
cat gpurun_out/bench_render_1gpu_$TAG.json
