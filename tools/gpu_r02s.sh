#!/bin/bash
mkdir -p gpurun_out
timeout 400 python bench.py --workload interleaved --steps 200 --warmup 16 > gpurun_out/bench_interleaved_1gpu_r02z.json 2> gpurun_out/bench_interleaved_1gpu_r02z.err; echo "rc=$?"; tail -2 gpurun_out/bench_interleaved_1gpu_r02z.err
python -c "
import json; d=json.load(open('gpurun_out/bench_interleaved_1gpu_r02z.json')); print({k: d[k] for k in d if k in ('metric','value','unit','ms_per_step')}); print(d.get('e2e'))"
