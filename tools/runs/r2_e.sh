#!/bin/bash
mkdir -p gpurun_out
echo "== kernel tests"
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm or conv or split" 2>&1 | tail -3
echo "== small GEMMs, pair tiles + split-K"
CB200_TAPGEMM_VERBOSE=1 timeout 300 python tools/bench_small_gemm.py gpurun_out/r2f_small_gemm.json 2> gpurun_out/r2f_verbose.err
sort gpurun_out/r2f_verbose.err | uniq -c | sort -rn | head -24
echo "== small GEMMs, deep ring only"
CB200_TAPGEMM_SPLITK=0 timeout 300 python tools/bench_small_gemm.py gpurun_out/r2f_small_gemm_nosplit.json
echo "== trace 64"
timeout 300 python tools/trace_step.py --out gpurun_out/r2f_trace_n1_b64 --global-batch 64 2> gpurun_out/r2f_trace64.err | cut -c1-400
