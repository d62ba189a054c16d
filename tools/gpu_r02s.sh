#!/bin/bash
timeout 300 python tools/exp_partition_cost.py 2>&1 | tee gpurun_out/exp_partition_cost_r02.log | tail -8
