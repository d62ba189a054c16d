#!/bin/bash
# N GPUs (default 2): the multi-device tests on real devices, then the bench lines of every multi-GPU workload
NP=${1:-2}; TAG=${2:-r02v}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
if [ "$NP" = "2" ]; then
VNR_COMM_SHARE_DEVICES=0 timeout 600 python -m pytest tests/test_gpu_comm.py tests/test_gpu_distributed.py tests/test_gpu_outofcore.py -m gpu -q > gpurun_out/pytest_${NP}gpu_$TAG.log 2>&1; echo "pytest ${NP}gpu rc=$?"; tail -3 gpurun_out/pytest_${NP}gpu_$TAG.log
fi
run() { local name=$1 np=$2 port=$3; shift 3
  timeout 400 $TR --nproc-per-node $np --master-port $port bench.py --gpus $np "$@" 2> gpurun_out/$name.err | grep '^{' > gpurun_out/$name.json
  echo "$name rc=$? $(python -c "
import json
try:
    d=json.load(open('gpurun_out/$name.json')); e=d['e2e']
    print(d['metric'], round(d['value'],1), d['unit'], 'ms/step', round(d['ms_per_step'],4), 'e2e', round(e['value'],1), 'e2e fps', e.get('fps'), 'inflight', e.get('fps_with_frames_in_flight'), 'dp', d.get('dp_steps_per_sec'), 'parity', d.get('parity_checked'), 'ooc', (d.get('out_of_core') or {}).get('bytes_uploaded_per_step'))
except Exception as ex: print('no json', ex)")"
}
run bench_render_${NP}gpu_$TAG $NP 29601 --steps 128 --warmup 8 --train-steps 300 --cpu-seconds 1
run bench_train_${NP}gpu_$TAG $NP 29602 --workload train --steps 100 --warmup 10
run bench_train_ooc1024_${NP}gpu_$TAG $NP 29603 --workload train --out-of-core --volume 1024 --steps 60 --warmup 10
if [ "$NP" = "8" ]; then
run bench_render4k_t22_8gpu_$TAG 8 29604 --width 3840 --height 2160 --log2-hashmap 22 --steps 32 --warmup 4 --train-steps 300 --cpu-seconds 1
run bench_render_4gpu_$TAG 4 29605 --steps 128 --warmup 8 --train-steps 300 --cpu-seconds 1
run bench_train_4gpu_$TAG 4 29606 --workload train --steps 100 --warmup 10

fi
tail -q -n 2 gpurun_out/*_$TAG.err | tail -8
