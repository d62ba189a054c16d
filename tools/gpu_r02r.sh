#!/bin/bash
mkdir -p gpurun_out
for it in 16 32; do
VNR_RM_N_ITERS=$it timeout 400 python bench.py --steps 128 --cpu-seconds 1 > gpurun_out/bench_iters$it.json 2> gpurun_out/bench_iters$it.err; echo "iters $it rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_iters$it.json"))
print("n_iters $it: value", round(d["value"]/1e9,3), "fps", round(d["fps"],1), "e2e fps", round(d["e2e"]["fps"],1), "copy", round(d["e2e"]["fps_copy_after_frame"],1), "inflight", d["e2e"]["fps_with_frames_in_flight_by_download"], "launches/frame", d["gpu_launches"]/d["steps"], "samples/frame", d["samples_per_frame"])
PY
done
timeout 600 python -m pytest tests/test_gpu_render.py tests/test_gpu_modes.py tests/test_gpu_reference_marcher.py -m gpu -q > gpurun_out/pytest_gpu_r02r.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r02r.log
