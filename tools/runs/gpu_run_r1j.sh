#!/bin/bash
mkdir -p gpurun_out
CB200_TAPGEMM_PAIR=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "conv_fwd_and_dgrad or gemm_nt" 2>&1 | tail -15 > gpurun_out/pytest_r1j_pair.log; tail -6 gpurun_out/pytest_r1j_pair.log
timeout 500 python -m pytest tests -m gpu -q -s 2>&1 > gpurun_out/pytest_r1j.log; grep -E "passed|failed|^FAILED|G grad-norm|G grad rel|per-parameter" gpurun_out/pytest_r1j.log | cut -c1-1500
for mode in 0 1; do
  echo "== pair mode $mode"; CB200_TAPGEMM_PAIR=$mode timeout 200 python tools/gpu_probe_r1a.py 2>&1 | grep -E "conv|heads" | grep -v cudnn
done
CB200_TAPGEMM_PAIR=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r1j.json 2> gpurun_out/bench_r1j.err; cut -c1-330 gpurun_out/bench_r1j.json; tail -3 gpurun_out/bench_r1j.err
