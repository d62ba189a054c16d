#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/exp_pair.py 2>&1 | tail -8
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 128 --cpu-seconds 1 > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "$name rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$name.json")); r=d["roofline"]
print("$name: value", round(d["value"]/1e9,3), "fps", round(d["fps"],1), "e2e fps", round(d["e2e"]["fps"],1), "decode in-frame G/s", round(r["decode_samples_per_sec"]/1e9,3), "uniform G/s", round(r["decode_uniform_samples_per_sec"]/1e9,3))
PY
}
run pair0 VNR_DECODE_PAIR=0
run pair1 VNR_DECODE_PAIR=1
VNR_DECODE_PAIR=1 timeout 300 python -m pytest tests/test_gpu_decode.py tests/test_gpu_render.py -m gpu -x -q 2>&1 | tail -3
