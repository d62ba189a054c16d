#!/bin/bash
# ncu evidence for profiles/: launch list of one profiling pass + full captures of the three hot kernels.
# Run on the GPU box:  gpurun --timeout 1500 -- bash tools/profile_r01.sh r01b
TAG=${1:-r01}
# VNR_RM_GRAPH=0: ncu does not list kernels that run inside a conditional graph body; the host-enqueued path launches the same kernels
export TRAIN_STEPS=100 FRAMES=3 EXTRA_TRAIN=4 VNR_RM_GRAPH=0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/launches_$TAG.csv python tools/profile_render.py > gpurun_out/prof_launch.log 2>&1; echo launches rc=$?
timeout 400 ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 0 -c 3 -f -o gpurun_out/decode_$TAG python tools/profile_render.py > gpurun_out/prof_decode.log 2>&1; echo decode rc=$?
EXTRA_TRAIN=3 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"train_step_kernel|adam_grid_kernel" -s 202 -c 4 -f -o gpurun_out/train_$TAG python tools/profile_render.py > gpurun_out/prof_train.log 2>&1; echo train rc=$?
timeout 400 ncu --set full --clock-control none --import-source on -k regex:march_round_kernel -s 0 -c 3 -f -o gpurun_out/march_$TAG python tools/profile_render.py > gpurun_out/prof_march.log 2>&1; echo march rc=$?
