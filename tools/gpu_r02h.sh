#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/exp_dp_debug.py > gpurun_out/exp_dp_debug.log 2>&1; echo "dp debug rc=$?"; tail -40 gpurun_out/exp_dp_debug.log
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 120 python tools/exp_dp_debug.py > gpurun_out/exp_dp_debug32.log 2>&1; echo "dp debug (32 connections) rc=$?"; tail -12 gpurun_out/exp_dp_debug32.log
timeout 600 python tools/exp_train_parity.py > gpurun_out/exp_train_parity_h.log 2>&1; echo "parity-exp rc=$?"; tail -70 gpurun_out/exp_train_parity_h.log
