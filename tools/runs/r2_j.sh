#!/bin/bash
# round 2, call J (1 GPU): strict StyleGAN2 test (verbose), dropin worker, tensor-core contrastive path, config-4 reference leg
mkdir -p gpurun_out
echo "== sg2 strict"
timeout 600 python -m pytest tests/test_gpu_sg2.py -x -q -k "strict" 2>&1 | grep -E "strict:|Error|error|passed|failed" | cut -c1-1500 | head -12
echo "== dropin worker"
timeout 300 python -m pytest tests/test_gpu_dropin.py -x -q -s 2>&1 | tail -5 | cut -c1-600
echo "== tensor-core contrastive"
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "tensor_core_contrastive" -s 2>&1 | tail -5 | cut -c1-900
echo "== config 4 reference eager (unmodified reference, prebuilt ops)"
timeout 400 python tools/bench_sg2.py --impl reference --steps 10 --warmup 3 2>&1 | tail -3 | cut -c1-700
echo "== config 4 native graph"
timeout 300 python tools/bench_sg2.py --graph --steps 12 --warmup 3 2>&1 | tail -1 | cut -c1-500
