#!/bin/bash
mkdir -p gpurun_out
export VNR_L2_PERSIST=0
VNR_TRAIN_TIMING=1 timeout 200 python bench.py --workload train --steps 200 --warmup 20 2>&1 >/dev/null | grep "train step"
