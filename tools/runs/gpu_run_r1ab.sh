#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E|passed|failed|^FAILED" | head -20
timeout 300 python bench.py --steps 40 --warmup 10 --no-cpu-baseline > gpurun_out/bench_r1ab.json 2> gpurun_out/bench_r1ab.err; grep -v "bench rank" gpurun_out/bench_r1ab.err | tail -3; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r1ab.json"))
print({k:d[k] for k in ("value","ms_per_step","e2e","gpu_launches")})
print(d["kernel_time_share"]); print(d["roofline"])
PY
