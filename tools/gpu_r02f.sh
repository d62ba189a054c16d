#!/bin/bash
# r02f: state of the tree after the re-created container: role timers of the training kernel, the full GPU suite,
# the default bench line, the train line and both reference arms.
mkdir -p gpurun_out
t0=$(date +%s)
timeout 300 python tools/exp_train_roles.py > gpurun_out/exp_train_roles_f.log 2>&1; echo "exp rc=$? ($(( $(date +%s) - t0 )) s)"; grep -v "loads_\|reds_\|copy_" gpurun_out/exp_train_roles_f.log | tail -30
cp gpurun_out/exp_train_roles.json gpurun_out/exp_train_roles_f.json 2>/dev/null
t0=$(date +%s)
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_r02f.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - t0 )) s)"; tail -12 gpurun_out/pytest_gpu_r02f.log
t0=$(date +%s)
timeout 400 python bench.py > gpurun_out/bench_render_1gpu_r02f.json 2> gpurun_out/bench_render_1gpu_r02f.err; echo "bench rc=$? ($(( $(date +%s) - t0 )) s)"; tail -3 gpurun_out/bench_render_1gpu_r02f.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_render_1gpu_r02f.json"))
print("value", d["value"]/1e9, "fps", d["fps"], "e2e fps", d["e2e"]["fps"], "copy", d["e2e"]["fps_copy_after_frame"], "inflight", d["e2e"]["fps_with_frames_in_flight_by_download"], "train", d["train_steps_per_sec_batch_2p18"])
r=d["roofline"]; print({k:r[k] for k in ("bound","achieved","peak","frac","l2_gather_gbs","hbm_gather_gbs","decode_samples_per_sec","decode_uniform_samples_per_sec","frac_uniform","decode_ms_per_frame","hbm_copy_frac")})
print(d["cpu_baseline"]); print(d["clocks"])
PY
timeout 300 python bench.py --workload train --steps 100 > gpurun_out/bench_train_1gpu_r02f.json 2> gpurun_out/bench_train_1gpu_r02f.err; echo "train rc=$?"; cat gpurun_out/bench_train_1gpu_r02f.json | cut -c1-600
timeout 400 python bench.py --impl reference --steps 40 > gpurun_out/bench_reference_1gpu_r02f.json 2> gpurun_out/bench_reference_1gpu_r02f.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/bench_reference_1gpu_r02f.json
timeout 300 python bench.py --impl reference --workload train --steps 100 > gpurun_out/bench_reference_train_1gpu_r02f.json 2> gpurun_out/bench_reference_train_1gpu_r02f.err; echo "ref train rc=$?"; cut -c1-300 gpurun_out/bench_reference_train_1gpu_r02f.json
