#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/exp_train_roles.py > gpurun_out/exp_train_roles_e.log 2>&1; echo "exp rc=$?"; grep -v "loads_\|reds_\|copy_" gpurun_out/exp_train_roles_e.log | tail -30
timeout 600 python -m pytest tests/test_gpu_comm.py -m gpu -q -x -s > gpurun_out/pytest_comm_r02e.log 2>&1; echo "comm pytest rc=$?"; tail -30 gpurun_out/pytest_comm_r02e.log
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py --deselect tests/test_gpu_comm.py > gpurun_out/pytest_gpu_r02e.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu_r02e.log
