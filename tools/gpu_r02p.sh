#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/exp_probe_mixed.py > gpurun_out/exp_probe_mixed.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/exp_probe_mixed.log
