#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/exp_train_roles.py > gpurun_out/exp_train_roles_d.log 2>&1; echo "exp rc=$?"; grep -v "loads_\|reds_\|copy_" gpurun_out/exp_train_roles_d.log | tail -30
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py > gpurun_out/pytest_gpu_r02d.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu_r02d.log
timeout 400 python bench.py > gpurun_out/bench_render_1gpu_r02d.json 2> gpurun_out/bench_render_1gpu_r02d.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_render_1gpu_r02d.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_render_1gpu_r02d.json"))
print("value", d["value"]/1e9, "fps", d["fps"], "e2e fps", d["e2e"]["fps"], "inflight", d["e2e"]["fps_with_frames_in_flight_by_download"], "train", d["train_steps_per_sec_batch_2p18"])
r=d["roofline"]; print({k:r[k] for k in ("achieved","peak","frac","decode_samples_per_sec","decode_uniform_samples_per_sec","frac_uniform","decode_ms_per_frame")})
PY
