#!/bin/bash
mkdir -p gpurun_out
VNR_COMM_SHARE_DEVICES=0 timeout 600 python -m pytest tests/test_gpu_comm.py tests/test_gpu_distributed.py tests/test_gpu_outofcore.py -m gpu -q > gpurun_out/pytest_2gpu_r02y.log 2>&1; echo "pytest 2gpu rc=$?"; tail -2 gpurun_out/pytest_2gpu_r02y.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for m in serial overlap; do
if [ $m = serial ]; then export VNR_TRAIN_SERIAL=1; else unset VNR_TRAIN_SERIAL; fi
timeout 400 $TR --nproc-per-node 2 --master-port 29611 bench.py --gpus 2 --workload train --steps 200 --warmup 20 2> gpurun_out/bench_train_2gpu_$m.err | grep '^{' > gpurun_out/bench_train_2gpu_$m.json
python -c "
import json; d=json.load(open('gpurun_out/bench_train_2gpu_$m.json')); print('$m', round(d['value'],1), 'steps/s, ms', round(d['ms_per_step'],4), 'loss', d['mean_loss'], d['last_loss'])"
done
