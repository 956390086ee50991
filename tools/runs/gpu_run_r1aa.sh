#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E|passed|failed" | head -12
timeout 300 python bench.py --steps 40 --warmup 10 > gpurun_out/bench_r1aa.json 2> gpurun_out/bench_r1aa.err; grep -v "bench rank" gpurun_out/bench_r1aa.err | tail -3; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r1aa.json"))
print({k:d[k] for k in ("value","ms_per_step","e2e","gpu_launches")})
print(d["kernel_time_share"])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 500 --csv --log-file gpurun_out/launches_r1aa.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_under_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/launches_r1aa.csv > gpurun_out/launch_summary_r1aa.txt 2>&1; head -30 gpurun_out/launch_summary_r1aa.txt
bash tools/gpu_profile_r1.sh 2>&1 | tail -3
