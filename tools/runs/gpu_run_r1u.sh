#!/bin/bash
# round-1 checkpoint: GPU tests, bench, launch list, full ncu captures of the dominant kernels
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 300 python bench.py --steps 40 --warmup 10 > gpurun_out/bench_r1u.json 2> gpurun_out/bench_r1u.err; cat gpurun_out/bench_r1u.json; tail -3 gpurun_out/bench_r1u.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 500 --csv --log-file gpurun_out/launches_r1u.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/launches_r1u.csv > gpurun_out/launch_summary_r1u.txt 2>&1; head -45 gpurun_out/launch_summary_r1u.txt
bash tools/gpu_profile_r1.sh
