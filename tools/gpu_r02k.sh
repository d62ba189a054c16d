#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python tools/exp_train_quality.py > gpurun_out/exp_train_quality.log 2>&1; echo "quality rc=$?"; tail -40 gpurun_out/exp_train_quality.log
