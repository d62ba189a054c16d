#!/bin/bash
# r02a: where the training step's time goes + full-size parity against the reference tcnn.
mkdir -p gpurun_out
t0=$(date +%s)
timeout 600 python tools/exp_train_roles.py > gpurun_out/exp_train_roles.log 2>&1; echo "exp rc=$? ($(( $(date +%s) - t0 )) s)"; tail -60 gpurun_out/exp_train_roles.log
t0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_r02a.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - t0 )) s)"; tail -15 gpurun_out/pytest_gpu_r02a.log
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -s > gpurun_out/pytest_fullsize_r02a.log 2>&1; echo "fullsize rc=$?"; grep -v "^$" gpurun_out/pytest_fullsize_r02a.log | tail -40
