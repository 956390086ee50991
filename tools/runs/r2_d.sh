#!/bin/bash
mkdir -p gpurun_out
echo "== small GEMMs, split-K on"
CB200_TAPGEMM_VERBOSE=1 timeout 300 python tools/bench_small_gemm.py gpurun_out/r2d_small_gemm_split1.json 2> gpurun_out/r2d_verbose.err
sort gpurun_out/r2d_verbose.err | uniq -c | sort -rn | head -40
echo "== small GEMMs, split-K off"
CB200_TAPGEMM_SPLITK=0 timeout 300 python tools/bench_small_gemm.py gpurun_out/r2d_small_gemm_split0.json
echo "== trace 64"
timeout 300 python tools/trace_step.py --out gpurun_out/r2d_trace_n1_b64 --global-batch 64 2> gpurun_out/r2d_trace64.err | cut -c1-400
