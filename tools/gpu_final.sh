#!/bin/bash
# Round-end style pass on one GPU: smoke, parity suite, both arms of both workloads.   gpurun --timeout 1700 -- bash tools/gpu_final.sh r01h
TAG=${1:-r01}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log
t0=$(date +%s)
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - t0 )) s)"; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 400 python bench.py --impl reference > gpurun_out/bench_reference_1gpu_$TAG.json 2> gpurun_out/bench_reference_1gpu_$TAG.err; echo "reference rc=$?"
timeout 400 python bench.py > gpurun_out/bench_render_1gpu_$TAG.json 2> gpurun_out/bench_render_1gpu_$TAG.err; echo "bench rc=$?"
timeout 300 python bench.py --workload train --impl reference > gpurun_out/bench_reference_train_1gpu_$TAG.json 2> gpurun_out/bench_reference_train_1gpu_$TAG.err; echo "reference train rc=$?"
timeout 300 python bench.py --workload train > gpurun_out/bench_train_1gpu_$TAG.json 2> gpurun_out/bench_train_1gpu_$TAG.err; echo "train rc=$?"
for f in bench_reference_1gpu bench_render_1gpu bench_reference_train_1gpu bench_train_1gpu; do cut -c1-420 gpurun_out/${f}_$TAG.json; echo; done
