#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/exp_grads_fullsize.py > gpurun_out/exp_grads_fullsize.log 2>&1; echo "grads rc=$?"; tail -60 gpurun_out/exp_grads_fullsize.log
