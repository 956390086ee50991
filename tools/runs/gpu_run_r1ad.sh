#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E|passed|failed|^FAILED" | head -20
timeout 300 python bench.py --steps 30 --warmup 8 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_r1ad.json; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r1ad.json"))
print({k:d[k] for k in ("value","ms_per_step","e2e","gpu_launches")})
print(d["kernel_time_share"])
PY
CB200_TAPGEMM_PERSIST=1 timeout 100 python tools/profile_target.py conv_first_wgrad
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv_first|tap_gemm_persist_kernel<32" -s 5 -c 12 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph 2>&1 | grep -E "conv_first|tap_gemm|gpu__time" | paste - - | cut -c1-200 | head -12
