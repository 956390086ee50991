#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -s 2>&1 > gpurun_out/pytest_r1h.log; grep -E "passed|failed|^E  |G grad-norm|step [12]:|^FAILED" gpurun_out/pytest_r1h.log | cut -c1-600
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r1h.json 2> gpurun_out/bench_r1h.err; cut -c1-400 gpurun_out/bench_r1h.json; tail -3 gpurun_out/bench_r1h.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 700 --csv --log-file gpurun_out/launches_r1h.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/launches_r1h.csv | head -40
