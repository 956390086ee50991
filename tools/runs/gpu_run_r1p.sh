#!/bin/bash
mkdir -p gpurun_out
echo "== persist kernels: correctness"
CB200_TAPGEMM_PERSIST=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "conv_fwd_and_dgrad or gemm_nt" 2>&1 | tail -8
echo "== perf"
for mode in "0 0" "1 0" "0 1"; do
  set -- $mode
  echo "== pair=$1 persist=$2"
  CB200_TAPGEMM_PAIR=$1 CB200_TAPGEMM_PERSIST=$2 timeout 120 python tools/gpu_probe_r1a.py 2>&1 | grep -E "conv fwd|conv dgrad|heads"
done
CB200_TAPGEMM_PERSIST=1 timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -q 2>&1 | tail -5
CB200_TAPGEMM_PERSIST=1 timeout 300 python bench.py --steps 30 --warmup 8 --no-cpu-baseline > gpurun_out/bench_r1p.json 2> gpurun_out/bench_r1p.err; cut -c1-330 gpurun_out/bench_r1p.json; tail -3 gpurun_out/bench_r1p.err
