#!/bin/bash
TAG=${1:-r01g}
export TRAIN_STEPS=100 FRAMES=3 EXTRA_TRAIN=0 VNR_RM_GRAPH=0 NO_DOWNLOAD=1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/launches_$TAG.csv python tools/profile_render.py > gpurun_out/prof_launch.log 2>&1; echo "launches rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:march_round_kernel -s 0 -c 2 -f -o gpurun_out/march_$TAG python tools/profile_render.py > gpurun_out/prof_march.log 2>&1; echo "march rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 0 -c 2 -f -o gpurun_out/decode_$TAG python tools/profile_render.py > gpurun_out/prof_decode.log 2>&1; echo "decode rc=$?"
