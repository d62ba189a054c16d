#!/bin/bash
# 8-GPU box: the driver's scaling run (default flags) at N = 8 and 4, plus the train workload at 8
TAG=${1:-r02x}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { local name=$1 np=$2 port=$3; shift 3
  timeout 400 $TR --nproc-per-node $np --master-port $port bench.py --gpus $np "$@" 2> gpurun_out/$name.err | grep '^{' > gpurun_out/$name.json
  echo "$name rc=$? $(python -c "
import json
try:
    d=json.load(open('gpurun_out/$name.json')); e=d['e2e']
    print(d['metric'], round(d['value'],1), d['unit'], 'ms/step', round(d['ms_per_step'],4), 'e2e', round(e['value'],1), 'e2e fps', e.get('fps'), 'inflight', e.get('fps_with_frames_in_flight'), 'dp', d.get('dp_steps_per_sec'), 'parity', d.get('parity_checked'), d['config'].get('frames_in_flight','')[:20])
except Exception as ex: print('no json', ex)")"
}
run bench_render_8gpu_$TAG 8 29601 --steps 256 --warmup 16 --train-steps 300 --cpu-seconds 1
run bench_render_4gpu_$TAG 4 29605 --steps 256 --warmup 16 --train-steps 300 --cpu-seconds 1
run bench_train_8gpu_$TAG 8 29602 --workload train --steps 100 --warmup 10
tail -q -n 2 gpurun_out/*_$TAG.err | tail -6
