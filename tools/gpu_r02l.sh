#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/exp_wgrad_half.py > gpurun_out/exp_wgrad_half.log 2>&1; echo "wgrad rc=$?"; tail -30 gpurun_out/exp_wgrad_half.log
