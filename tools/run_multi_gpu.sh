#!/bin/bash
# Multi-GPU bench lines for profiles/ (run on an 8-GPU box):  gpurun --gpus 8 --timeout 1500 -- bash tools/run_multi_gpu.sh
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { # name nproc port args...
  local name=$1 np=$2 port=$3; shift 3
  timeout 400 $TR --nproc-per-node $np --master-port $port bench.py --gpus $np "$@" 2> gpurun_out/$name.err | grep '^{' > gpurun_out/$name.json
  echo "$name rc=$? $(python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/$name.json')); print(d['metric'], round(d['value'],1), d['unit'], 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1))
except Exception as e: print('no json', e)")"
}
run render_8gpu 8 29601 --steps 64 --warmup 8 --train-steps 300
run train_8gpu 8 29602 --workload train --steps 50 --warmup 5
run render4k_t22_8gpu 8 29603 --width 3840 --height 2160 --log2-hashmap 22 --steps 32 --warmup 4 --train-steps 300
run render_4gpu 4 29604 --steps 64 --warmup 8 --train-steps 300
run train_4gpu 4 29605 --workload train --steps 50 --warmup 5
run train_8gpu_allreduce 8 29606 --workload train --steps 50 --warmup 5 --dp-mode allreduce
tail -3 gpurun_out/*.err | tail -30
