#!/bin/bash
# round 2, call I (1 GPU): augment A/B + ncu, strict-mode tests (SNDCGAN two steps, StyleGAN2), dropin worker, bench
mkdir -p gpurun_out
echo "== augment parity"
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "augment" 2>&1 | tail -3
echo "== augment OCC=6 (default)"; timeout 120 python tools/bench_augment.py
echo "== augment OCC=5"; CB200_AUGMENT_OCC=5 timeout 120 python tools/bench_augment.py
echo "== ncu augment"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:augment_simclr_fwd_cols -c 1 -s 3 -o gpurun_out/prof_r2_augment -f python tools/profile_target.py augment > gpurun_out/prof_r2_augment.log 2>&1; tail -2 gpurun_out/prof_r2_augment.log
echo "== config1 strict full + config2"
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_zz_next_rows.py -x -q -k "config1 or config2" -s 2>&1 | grep -E "^step|config 2|passed|failed|assert|Error" | head -20
echo "== sg2 strict"
timeout 900 python -m pytest tests/test_gpu_sg2.py -x -q -k "strict or matches_reference" 2>&1 | tail -6
echo "== dropin worker"
timeout 600 python -m pytest tests/test_gpu_dropin.py -x -q -s 2>&1 | tail -6
echo "== bench N=1"
timeout 900 python bench.py --steps 20 --warmup 5 --no-side-workloads > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err
echo "rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2i_bench_n1.json'))
print({k: d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d.get('eager_gpu_baseline'), d.get('cpu_baseline'), d.get('roofline_augment'))
print(d.get('roofline')); print(d.get('tensor_kernels')); print(d.get('kernel_time_share'))
PY
grep -E "bench rank|Error|error" gpurun_out/r2i_bench_n1.err | tail -8
