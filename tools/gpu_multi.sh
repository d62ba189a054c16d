#!/bin/bash
# N-GPU pass: distributed parity tests + both workloads at N ranks.   gpurun --gpus 2 --timeout 900 -- bash tools/gpu_multi.sh 2 r01f
N=${1:-2}; TAG=${2:-r01}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q > gpurun_out/pytest_dist_${N}gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_dist_${N}gpu_$TAG.log
P=$((29500 + RANDOM % 500))
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N > gpurun_out/bench_render_${N}gpu_$TAG.json 2> gpurun_out/bench_render_${N}gpu_$TAG.err; echo "render rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus $N --workload train > gpurun_out/bench_train_${N}gpu_$TAG.json 2> gpurun_out/bench_train_${N}gpu_$TAG.err; echo "train rc=$?"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+2)) bench.py --gpus $N --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference_${N}gpu_$TAG.json 2> gpurun_out/bench_reference_${N}gpu_$TAG.err; echo "reference rc=$?"
cat gpurun_out/bench_render_${N}gpu_$TAG.json gpurun_out/bench_train_${N}gpu_$TAG.json | cut -c1-700
