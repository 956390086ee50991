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
echo "== model tests (default + strict)"
timeout 900 python -m pytest tests/test_gpu_model.py -x -q 2>&1 | tail -4
echo "== config 2 full batch (default + strict)"
timeout 900 python -m pytest tests/test_gpu_zz_next_rows.py -x -q -k "config2" -s 2>&1 | grep -E "config 2|passed|failed|Error|assert" | head
echo "== bench strict vs default (short)"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-side-workloads --no-eager-baseline --no-u8-input 2>/dev/null | cut -c1-330
CB200_PRECISION=strict timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-side-workloads --no-eager-baseline --no-u8-input 2>/dev/null | cut -c1-330
