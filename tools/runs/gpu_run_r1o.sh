#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -6
timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -q -s 2>&1 > gpurun_out/pytest_r1o.log; grep -E "passed|failed|^FAILED|^E  " gpurun_out/pytest_r1o.log | cut -c1-600 | head -20
for pair in 0 1; do
  echo "== pair=$pair"
  CB200_TAPGEMM_PAIR=$pair timeout 120 python tools/gpu_probe_r1a.py 2>&1 | grep -E "conv fwd|conv dgrad|heads"
done
timeout 300 python bench.py --steps 30 --warmup 8 --no-cpu-baseline > gpurun_out/bench_r1o.json 2> gpurun_out/bench_r1o.err; cut -c1-330 gpurun_out/bench_r1o.json; tail -3 gpurun_out/bench_r1o.err
