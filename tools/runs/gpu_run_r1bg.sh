#!/bin/bash
# launch list of the StyleGAN2 config-4 step (2 steps after 1 warm-up; the capture skips the warm-up's launches approximately)
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 2048 --csv --log-file gpurun_out/launches_sg2_r1.csv python tools/bench_sg2.py --steps 3 --warmup 1 > gpurun_out/bench_sg2_under_ncu.log 2>&1
tail -2 gpurun_out/bench_sg2_under_ncu.log | cut -c1-300
cd tools && python launch_summary_all.py ../gpurun_out/launches_sg2_r1.csv 2 > ../gpurun_out/launches_sg2_r1.md; head -40 ../gpurun_out/launches_sg2_r1.md
