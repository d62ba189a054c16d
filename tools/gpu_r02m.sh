#!/bin/bash
mkdir -p gpurun_out
t0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_r02m.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - t0 )) s)"; grep -v "^$" gpurun_out/pytest_gpu_r02m.log | grep -i "loss \|first step\|volume PSNR\|MLP (half\|grid:\|passed\|failed\|FAILED\|Error" | tail -40
