#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/exp_train_pace.py > gpurun_out/exp_train_pace_q.log 2>&1; echo "pace rc=$?"; cat gpurun_out/exp_train_pace_q.log
