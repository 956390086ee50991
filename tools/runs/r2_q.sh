#!/bin/bash
# round 2, call Q (1 GPU): the GPU suite as the driver runs it, smoke(), the default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -6 | cut -c1-600
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (defaults)"
timeout 900 python bench.py > gpurun_out/r2q_bench_n1.json 2> gpurun_out/r2q_bench_n1.err
echo "rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2q_bench_n1.json'))
print({k: d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'])
print(d.get('precision_modes')); print(d.get('eager_gpu_baseline', {}).get('ratio'), d.get('cpu_baseline'))
print(d.get('roofline')); print(d.get('roofline_augment')); print(d.get('other_workloads'))
PY
grep -E "bench rank|Error|error" gpurun_out/r2q_bench_n1.err | tail -8
