#!/bin/bash
# r02b: epilogue phase timers of the training chain, training parity vs the reference by batch size, full GPU suite after
# the frame-ring refactor, default bench line.
mkdir -p gpurun_out
t0=$(date +%s)
timeout 300 python tools/exp_train_roles.py > gpurun_out/exp_train_roles_b.log 2>&1; echo "exp rc=$? ($(( $(date +%s) - t0 )) s)"; grep -v "^loads\|^reds" gpurun_out/exp_train_roles_b.log | tail -30
t0=$(date +%s)
timeout 400 python tools/exp_train_parity.py > gpurun_out/exp_train_parity.log 2>&1; echo "parity-exp rc=$? ($(( $(date +%s) - t0 )) s)"; tail -60 gpurun_out/exp_train_parity.log
t0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_fullsize.py > gpurun_out/pytest_gpu_r02b.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - t0 )) s)"; tail -25 gpurun_out/pytest_gpu_r02b.log
timeout 400 python bench.py > gpurun_out/bench_render_1gpu_r02b.json 2> gpurun_out/bench_render_1gpu_r02b.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_render_1gpu_r02b.err; cat gpurun_out/bench_render_1gpu_r02b.json
