#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/exp_train_roles.py > gpurun_out/exp_train_roles_c.log 2>&1; echo "exp rc=$?"; grep -v "loads_\|reds_\|copy_" gpurun_out/exp_train_roles_c.log | tail -30
timeout 400 python bench.py > gpurun_out/bench_render_1gpu_r02c.json 2> gpurun_out/bench_render_1gpu_r02c.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_render_1gpu_r02c.err; cat gpurun_out/bench_render_1gpu_r02c.json
