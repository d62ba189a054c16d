#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/exp_dp_debug.py > gpurun_out/exp_dp_debug_i.log 2>&1; echo "dp debug rc=$?"; grep -v "^$" gpurun_out/exp_dp_debug_i.log | head -12
timeout 600 python tools/exp_train_parity.py 262144:64:300 65536:64:300 16384:64:0 > gpurun_out/exp_train_parity_i.log 2>&1; echo "parity-exp rc=$?"; tail -30 gpurun_out/exp_train_parity_i.log
timeout 600 python -m pytest tests/test_gpu_comm.py tests/test_gpu_outofcore.py tests/test_gpu_render.py tests/test_gpu_train.py -m gpu -q > gpurun_out/pytest_gpu_r02i.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_r02i.log
