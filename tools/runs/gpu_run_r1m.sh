#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_r1m.log; tail -4 gpurun_out/pytest_r1m.log
timeout 300 python bench.py --steps 30 --warmup 8 --no-cpu-baseline > gpurun_out/bench_r1m.json 2> gpurun_out/bench_r1m.err; cut -c1-330 gpurun_out/bench_r1m.json; tail -3 gpurun_out/bench_r1m.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 500 --csv --log-file gpurun_out/launches_r1m.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/launches_r1m.csv | head -48
