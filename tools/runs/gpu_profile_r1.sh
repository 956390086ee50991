#!/bin/bash
# ncu --set full captures of the dominant kernels (one launch each), summarised into profiles/ by tools/ncu_summary.py
mkdir -p gpurun_out
for tgt in conv3x3 conv4x4 dgrad wgrad augment; do
  case $tgt in
    conv3x3|conv4x4|dgrad) pat="regex:tap_gemm";;
    wgrad) pat="regex:wgrad_kernel";;
    augment) pat="regex:augment_simclr_fwd";;
  esac
  timeout 200 ncu --set full --clock-control none --import-source on -k $pat -s 2 -c 1 -f -o gpurun_out/prof_r1_$tgt python tools/profile_target.py $tgt > gpurun_out/prof_r1_$tgt.log 2>&1
  tail -2 gpurun_out/prof_r1_$tgt.log
done
ls -la gpurun_out/*.ncu-rep
