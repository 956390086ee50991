#!/bin/bash
# round 2, call P (8 GPUs): config 5 native - StyleGAN2_512 + ContraD, b64, 512x512, nn.DataParallel over 8 GPUs
mkdir -p gpurun_out
timeout 300 python tools/bench_sg2.py --data-parallel 8 --batch 64 --steps 16 --warmup 2 > gpurun_out/r2p_config5_native.json 2> gpurun_out/r2p_config5_native.err; tail -c 1200 gpurun_out/r2p_config5_native.json; tail -3 gpurun_out/r2p_config5_native.err | cut -c1-300
