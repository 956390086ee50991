#!/bin/bash
# final verification of the round: the driver's GPU test command, smoke(), the default bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/ -x -q -m gpu 2>&1 | grep -E "^E|passed|failed|^FAILED|^ERROR|Error" | head -12
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 400 python bench.py --steps 30 --warmup 8 > gpurun_out/bench_r1bz.json 2> gpurun_out/bench_r1bz.err; grep -v "bench rank" gpurun_out/bench_r1bz.err | tail -3; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r1bz.json"))
print({k:d[k] for k in ("value","ms_per_step","e2e","gpu_launches")})
print(d["roofline"]["frac"], d["cpu_baseline"]["value"], d.get("other_workloads"))
PY
