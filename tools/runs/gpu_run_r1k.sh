#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_r1k_kernels.log; tail -6 gpurun_out/pytest_r1k_kernels.log
timeout 500 python -m pytest tests/test_gpu_model.py -m gpu -q -s 2>&1 > gpurun_out/pytest_r1k.log; grep -E "passed|failed|^FAILED|G grad-norm|^E  " gpurun_out/pytest_r1k.log | cut -c1-400 | head -30
for mode in 1; do
  echo "== pair mode $mode"; CB200_TAPGEMM_PAIR=$mode timeout 200 python tools/gpu_probe_r1a.py 2>&1 | grep -E "conv|heads" | grep -v cudnn
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r1k.json 2> gpurun_out/bench_r1k.err; cut -c1-330 gpurun_out/bench_r1k.json; tail -3 gpurun_out/bench_r1k.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8
