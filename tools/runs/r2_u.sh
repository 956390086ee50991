#!/bin/bash
# round 2, call U (1 GPU, what is left of the budget): ncu --set full of the second build of the fused augment forward kernel
mkdir -p gpurun_out
timeout 40 ncu --set full --clock-control none --import-source on -k regex:augment_simclr_fwd_cols2 -c 1 -s 3 -o gpurun_out/prof_r2_augment_v2 -f python tools/profile_target.py augment > gpurun_out/prof_r2_augment_v2.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/prof_r2_augment_v2.log; ls -la gpurun_out/*.ncu-rep 2>/dev/null
