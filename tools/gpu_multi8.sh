#!/bin/bash
# 8-GPU lines for profiles/: render (strong), train (weak), 4K / T=2^22 render.   gpurun --gpus 8 --timeout 900 -- bash tools/gpu_multi8.sh r01i
TAG=${1:-r01}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
mkdir -p gpurun_out
run() { local name=$1 np=$2 port=$3; shift 3
  timeout 300 $TR --nproc-per-node $np --master-port $port bench.py --gpus $np "$@" 2> gpurun_out/$name.err | grep '^{' > gpurun_out/$name.json
  echo "$name rc=$? $(python -c "
import json
try:
    d=json.load(open('gpurun_out/$name.json')); print(d['metric'], round(d['value'],1), d['unit'], 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1))
except Exception as e: print('no json', e)")"
}
run bench_render_8gpu_$TAG 8 29601 --steps 128 --warmup 8 --train-steps 300 --cpu-seconds 1
run bench_train_8gpu_$TAG 8 29602 --workload train --steps 100 --warmup 5
run bench_render4k_t22_8gpu_$TAG 8 29603 --width 3840 --height 2160 --log2-hashmap 22 --steps 32 --warmup 4 --train-steps 300 --cpu-seconds 1
run bench_render_4gpu_$TAG 4 29604 --steps 128 --warmup 8 --train-steps 300 --cpu-seconds 1
tail -2 gpurun_out/*_$TAG.err | tail -12
