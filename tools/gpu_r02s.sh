#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_loop$i.log 2>&1; echo "run $i rc=$?"; tail -2 gpurun_out/pytest_gpu_loop$i.log; done
