#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python bench.py --steps 40 --warmup 10 > gpurun_out/bench_r1z.json 2> gpurun_out/bench_r1z.err; tail -3 gpurun_out/bench_r1z.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r1z.json"))
print({k:d[k] for k in ("value","ms_per_step","e2e","gpu_launches","cpu_baseline")})
PY
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
