#!/bin/bash
# One GPU-box pass: parity suite, both bench arms, ncu launch list + one full capture of the decode kernel.
#   gpurun --timeout 1700 -- bash tools/gpu_round.sh r01e
TAG=${1:-r01}
mkdir -p gpurun_out
t0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - t0 )) s)"; tail -3 gpurun_out/pytest_gpu_$TAG.log
t0=$(date +%s)
timeout 400 python bench.py > gpurun_out/bench_render_1gpu_$TAG.json 2> gpurun_out/bench_render_1gpu_$TAG.err; echo "bench rc=$? ($(( $(date +%s) - t0 )) s)"
t0=$(date +%s)
timeout 400 python bench.py --impl reference > gpurun_out/bench_reference_1gpu_$TAG.json 2> gpurun_out/bench_reference_1gpu_$TAG.err; echo "reference rc=$? ($(( $(date +%s) - t0 )) s)"
t0=$(date +%s)
timeout 300 python bench.py --workload train > gpurun_out/bench_train_1gpu_$TAG.json 2> gpurun_out/bench_train_1gpu_$TAG.err; echo "train rc=$? ($(( $(date +%s) - t0 )) s)"
export TRAIN_STEPS=100 FRAMES=3 EXTRA_TRAIN=4 VNR_RM_GRAPH=0
t0=$(date +%s)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/launches_$TAG.csv python tools/profile_render.py > gpurun_out/prof_launch.log 2>&1; echo "launches rc=$? ($(( $(date +%s) - t0 )) s)"
t0=$(date +%s)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 0 -c 2 -f -o gpurun_out/decode_$TAG python tools/profile_render.py > gpurun_out/prof_decode.log 2>&1; echo "decode rc=$? ($(( $(date +%s) - t0 )) s)"
cat gpurun_out/bench_render_1gpu_$TAG.json
