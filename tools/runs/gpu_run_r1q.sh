#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 120 python tools/gpu_probe_r1a.py 2>&1 | grep -E "conv fwd|conv dgrad|heads"
timeout 300 python bench.py --steps 30 --warmup 8 > gpurun_out/bench_r1q.json 2> gpurun_out/bench_r1q.err; cat gpurun_out/bench_r1q.json; tail -3 gpurun_out/bench_r1q.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 500 --csv --log-file gpurun_out/launches_r1q.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/launches_r1q.csv | head -40
