#!/bin/bash
# r02g: probes (32-byte loads, host stores), full GPU suite after the bias-table / per-rank slab fixes, bench line through the
# reworked multi-GPU plumbing (N = 1 here).
mkdir -p gpurun_out
timeout 200 python tools/exp_probes.py > gpurun_out/exp_probes_g.log 2>&1; echo "probes rc=$?"; cat gpurun_out/exp_probes_g.log | tail -12
t0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r02g.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - t0 )) s)"; tail -25 gpurun_out/pytest_gpu_r02g.log
t0=$(date +%s)
timeout 400 python bench.py > gpurun_out/bench_render_1gpu_r02g.json 2> gpurun_out/bench_render_1gpu_r02g.err; echo "bench rc=$? ($(( $(date +%s) - t0 )) s)"; tail -3 gpurun_out/bench_render_1gpu_r02g.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_render_1gpu_r02g.json"))
print("value", d["value"]/1e9, "fps", d["fps"], "e2e fps", d["e2e"]["fps"], "copy", d["e2e"]["fps_copy_after_frame"], "inflight", d["e2e"]["fps_with_frames_in_flight_by_download"], "train", d["train_steps_per_sec_batch_2p18"])
PY
