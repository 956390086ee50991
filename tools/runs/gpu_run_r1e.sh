#!/bin/bash
# GPU visit: model tests (verbose prints), bench, ncu launch list + full capture of the dominant kernels.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_model.py -m gpu -q -s 2>&1 | tail -60 > gpurun_out/pytest_r1e.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1e.json 2> gpurun_out/bench_r1e.err
cat gpurun_out/bench_r1e.json; tail -5 gpurun_out/bench_r1e.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -3 gpurun_out/bench_under_ncu.log
