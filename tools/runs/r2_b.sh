#!/bin/bash
# round 2, call B (1 GPU): kernel-name traces at 64 and 512 images per GPU
mkdir -p gpurun_out
timeout 300 python tools/trace_step.py --out gpurun_out/r2a_trace_n1_b64 --global-batch 64 2> gpurun_out/r2a_trace64.err | cut -c1-400
timeout 300 python tools/trace_step.py --out gpurun_out/r2a_trace_n1_b512 2> gpurun_out/r2a_trace512.err | cut -c1-400
