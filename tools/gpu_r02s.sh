#!/bin/bash
mkdir -p gpurun_out
show() { python - <<PY
import json
d=json.load(open("$1"))
print("$1".split("/")[-1], {k: d[k] for k in ("value","ms_per_step","mean_loss","last_loss","volume_psnr_db")}, "e2e", round(d["e2e"]["value"],1), d["roofline"]["frac"], d.get("out_of_core"))
PY
}
VNR_TRAIN_TIMING=1 timeout 300 python bench.py --workload train --steps 200 --warmup 20 > gpurun_out/bench_train_1gpu_r02u.json 2> gpurun_out/bench_train_1gpu_r02u.err; echo "train rc=$?"; grep "train step" gpurun_out/bench_train_1gpu_r02u.err | tail -1; show gpurun_out/bench_train_1gpu_r02u.json
timeout 600 python bench.py --workload train --out-of-core --volume 1024 --steps 200 --warmup 20 > gpurun_out/bench_train_ooc1024_1gpu_r02u.json 2> gpurun_out/bench_train_ooc1024_1gpu_r02u.err; echo "ooc rc=$?"; tail -3 gpurun_out/bench_train_ooc1024_1gpu_r02u.err; show gpurun_out/bench_train_ooc1024_1gpu_r02u.json
timeout 300 python -m pytest tests/test_gpu_train.py tests/test_gpu_fullsize.py tests/test_gpu_comm.py -m gpu -x -q 2>&1 | tail -3
