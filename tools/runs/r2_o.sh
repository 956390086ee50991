#!/bin/bash
# round 2, call O (2 GPUs): config-5 path through nn.DataParallel at 2 replicas x 8 images of 512x512 (validates the fix before the 8-GPU run)
mkdir -p gpurun_out
timeout 400 python tools/bench_sg2.py --data-parallel 2 --batch 16 --steps 16 --warmup 2 > gpurun_out/r2o_config5_dp2.json 2> gpurun_out/r2o_config5_dp2.err; tail -c 1200 gpurun_out/r2o_config5_dp2.json; tail -4 gpurun_out/r2o_config5_dp2.err | cut -c1-300
