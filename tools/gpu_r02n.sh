#!/bin/bash
# N = 2: the comm GPU tests on two real devices, then the default bench line at N = 2 (render through the communicator + DP training)
mkdir -p gpurun_out
VNR_COMM_SHARE_DEVICES=0 timeout 600 python -m pytest tests/test_gpu_comm.py tests/test_gpu_distributed.py tests/test_gpu_outofcore.py -m gpu -q -s > gpurun_out/pytest_2gpu_r02n.log 2>&1; echo "pytest 2gpu rc=$?"; tail -8 gpurun_out/pytest_2gpu_r02n.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 64 --warmup 8 > gpurun_out/bench_render_2gpu_r02n.json 2> gpurun_out/bench_render_2gpu_r02n.err; echo "bench N=2 rc=$?"; tail -5 gpurun_out/bench_render_2gpu_r02n.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_render_2gpu_r02n.json"))
    print("value", d["value"]/1e9, "fps", d["fps"], "e2e fps", d["e2e"]["fps"], "copy", d["e2e"]["fps_copy_after_frame"], "inflight", d["e2e"]["fps_with_frames_in_flight_by_download"])
    print("dp steps/s", d["dp_steps_per_sec"], "parity", d["parity"], d["parity_checked"])
except Exception as e:
    print("no line:", e)
PY
