#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "augment or adam" 2>&1 | tail -5
timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "graph" 2>&1 | tail -25
timeout 300 python bench.py --steps 40 --warmup 10 --no-cpu-baseline > gpurun_out/bench_r1x.json 2> gpurun_out/bench_r1x.err; tail -5 gpurun_out/bench_r1x.err; python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_r1x.json"))
    print({k:d[k] for k in ("value","ms_per_step","e2e","gpu_launches")})
except Exception as e: print("no json", e)
PY
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-graph 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('eager:', d['value'], d['e2e']['value'])"
