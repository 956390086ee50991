#!/bin/bash
# round 2, call A (1 GPU): new bench legs (reference eager on the GPU, reference CPU at b512), baseline traces
mkdir -p gpurun_out
export CB200_BENCH_WATCHDOG=600
echo "== bench N=1"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err
echo "rc=$?"; cut -c1-3000 gpurun_out/r2a_bench_n1.json; grep -E "bench rank|Error|error|File \"/root" gpurun_out/r2a_bench_n1.err | tail -20
echo "== trace N=1, 64 images (per-rank shape of the 8-GPU run, no collectives)"
timeout 300 python tools/trace_step.py --out gpurun_out/r2a_trace_n1_b64 --global-batch 64 2> gpurun_out/r2a_trace64.err | cut -c1-600
echo "== trace N=1, 512 images"
timeout 300 python tools/trace_step.py --out gpurun_out/r2a_trace_n1_b512 2> gpurun_out/r2a_trace512.err | cut -c1-600
tail -5 gpurun_out/r2a_trace64.err
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-600
