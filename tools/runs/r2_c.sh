#!/bin/bash
# round 2, call C (1 GPU): split-K + glue diet: kernel tests, model tests, traces, short bench
mkdir -p gpurun_out
echo "== kernel tests"
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm or conv or split" 2>&1 | tail -8
echo "== model tests"
timeout 900 python -m pytest tests/test_gpu_model.py -x -q 2>&1 | tail -8
echo "== trace 64"
timeout 300 python tools/trace_step.py --out gpurun_out/r2c_trace_n1_b64 --global-batch 64 2> gpurun_out/r2c_trace64.err | cut -c1-400
tail -3 gpurun_out/r2c_trace64.err
echo "== trace 512"
timeout 300 python tools/trace_step.py --out gpurun_out/r2c_trace_n1_b512 2> gpurun_out/r2c_trace512.err | cut -c1-400
echo "== bench"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-side-workloads --no-eager-baseline > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err
echo "rc=$?"; cut -c1-900 gpurun_out/r2c_bench_n1.json; grep -E "bench rank|Error|error" gpurun_out/r2c_bench_n1.err | tail -8
