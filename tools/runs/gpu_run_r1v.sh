#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -q -x -s -k "config1" 2>&1 | grep -E "step [0-9]|assert|passed|failed|Error" | head -20
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "augment or colsum or losses" 2>&1 | tail -3
echo "--- colsum vec"; timeout 100 python tools/bench_colsum.py
echo "--- colsum scalar"; CB200_COLSUM_VEC=0 timeout 100 python tools/bench_colsum.py
