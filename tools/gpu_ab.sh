#!/bin/bash
# A/B of the marcher's ray-order / slot-layout switches on the default bench + parity suite.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_ab.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_ab.log
for cfg in "1 1" "0 0" "1 0" "0 1"; do
  set -- $cfg
  VNR_RM_TILED=$1 VNR_RM_TRANSPOSE=$2 timeout 300 python bench.py --cpu-seconds 1 > gpurun_out/ab_t$1_x$2.json 2> gpurun_out/ab_t$1_x$2.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_t$1_x$2.json"))
print("tiled=$1 transpose=$2: fps", round(d["fps"],1), "value", round(d["value"]/1e9,3), "G/s e2e fps", round(d["e2e"]["fps"],1), "decode G/s", round(d["roofline"]["decode_samples_per_sec"]/1e9,3), "decode ms", d["roofline"]["decode_ms_per_frame"], "ms/frame", round(d["ms_per_step"],4))
PY
done
