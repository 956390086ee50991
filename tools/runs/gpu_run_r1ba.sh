#!/bin/bash
# round-1 session 2, call A: new StyleGAN2 GPU tests first, then the existing GPU suite, the config-4 side bench, the headline bench
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_sg2.py -x -q 2>&1 | tail -30
timeout 420 python -m pytest tests -m gpu -q --ignore=tests/test_gpu_sg2.py 2>&1 | grep -E "^E|passed|failed|^FAILED" | head -12
timeout 240 python tools/bench_sg2.py > gpurun_out/bench_sg2.json 2> gpurun_out/bench_sg2.err; tail -3 gpurun_out/bench_sg2.err; cat gpurun_out/bench_sg2.json
timeout 300 python bench.py --steps 30 --warmup 8 > gpurun_out/bench_r1ba.json 2> gpurun_out/bench_r1ba.err; grep -v "bench rank" gpurun_out/bench_r1ba.err | tail -3; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r1ba.json"))
print({k:d[k] for k in ("value","ms_per_step","e2e","gpu_launches")})
print(d["roofline"]); print(d["cpu_baseline"])
PY
