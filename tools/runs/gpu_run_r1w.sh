#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "augment or colsum or losses" 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -q 2>&1 | tail -3
echo "--- colsum"; timeout 100 python tools/bench_colsum.py | tail -4
timeout 300 python bench.py --steps 40 --warmup 10 --no-cpu-baseline > gpurun_out/bench_r1w.json 2> gpurun_out/bench_r1w.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r1w.json"))
print({k:d[k] for k in ("value","ms_per_step","e2e","roofline_augment")})
print(d["kernel_time_share"])
PY
CB200_AUGMENT_COLS=0 timeout 300 python bench.py --steps 10 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('quad mapping:', d['roofline_augment']['achieved'])"
