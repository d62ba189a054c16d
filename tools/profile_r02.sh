#!/bin/bash
# ncu evidence for profiles/ (round 2): launch list of one profiling pass + full captures of the hot kernels, for the default
# configuration (T = 2^19, 1024^2) and a decode capture of the 4K / T = 2^22 configuration (its DRAM traffic).
# Run on the GPU box:  gpurun --timeout 2400 -- bash tools/profile_r02.sh r02a
TAG=${1:-r02b}
mkdir -p gpurun_out
# VNR_RM_GRAPH=0: ncu does not list kernels that run inside a conditional graph body; the host-enqueued path launches the same kernels
export TRAIN_STEPS=100 FRAMES=3 EXTRA_TRAIN=4 VNR_RM_GRAPH=0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/launches_$TAG.csv python tools/profile_render.py > gpurun_out/prof_launch.log 2>&1; echo launches rc=$?
timeout 500 ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 0 -c 3 -f -o gpurun_out/decode_$TAG python tools/profile_render.py > gpurun_out/prof_decode.log 2>&1; echo decode rc=$?
EXTRA_TRAIN=3 timeout 500 ncu --set full --clock-control none --import-source on -k regex:"train_step_kernel|adam_grid_kernel" -s 202 -c 4 -f -o gpurun_out/train_$TAG python tools/profile_render.py > gpurun_out/prof_train.log 2>&1; echo train rc=$?
timeout 500 ncu --set full --clock-control none --import-source on -k regex:march_round_kernel -s 0 -c 3 -f -o gpurun_out/march_$TAG python tools/profile_render.py > gpurun_out/prof_march.log 2>&1; echo march rc=$?
# configs[4] shape on one GPU: 3840 x 2160, T = 2^22 (307 MB table > L2)
LOG2_HASHMAP=22 FRAME_W=3840 FRAME_H=2160 FRAMES=1 EXTRA_TRAIN=0 timeout 600 ncu --set full --clock-control none -k regex:decode_kernel -s 0 -c 1 -f -o gpurun_out/decode4k_$TAG python tools/profile_render.py > gpurun_out/prof_decode4k.log 2>&1; echo decode4k rc=$?
ls -la gpurun_out/*.ncu-rep
