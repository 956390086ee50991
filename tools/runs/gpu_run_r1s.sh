#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_model.py -m gpu -q -s -k "config1" 2>&1 | grep -E "step [12]:|^E  |passed|failed" | cut -c1-700
timeout 300 python bench.py --steps 30 --warmup 8 --no-cpu-baseline > gpurun_out/bench_r1s.json 2> gpurun_out/bench_r1s.err; cut -c1-250 gpurun_out/bench_r1s.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 500 --csv --log-file gpurun_out/launches_r1s.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/launches_r1s.csv 2>/dev/null | head -16
timeout 200 python tools/host_profile.py 2>&1 | head -60
