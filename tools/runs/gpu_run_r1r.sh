#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "wgrad" 2>&1 | tail -5
timeout 120 python tools/gpu_probe_r1a.py 2>&1 | grep -E "wgrad"
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python bench.py --steps 30 --warmup 8 --no-cpu-baseline > gpurun_out/bench_r1r.json 2> gpurun_out/bench_r1r.err; cut -c1-300 gpurun_out/bench_r1r.json; tail -3 gpurun_out/bench_r1r.err
