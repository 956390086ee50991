#!/bin/bash
# ncu --set full captures of the new HBM-bound kernels (one launch each), summarised into profiles/ by tools/ncu_summary.py
mkdir -p gpurun_out
for tgt in upfirdn bias_act augment_large; do
  case $tgt in
    upfirdn) pat="regex:upfirdn2d_nhwc4";;
    bias_act) pat="regex:bias_act_vec";;
    augment_large) pat="regex:augment_large_apply";;
  esac
  timeout 200 ncu --set full --clock-control none --import-source on -k $pat -s 2 -c 1 -f -o gpurun_out/prof_r1_$tgt python tools/profile_target.py $tgt > gpurun_out/prof_r1_$tgt.log 2>&1
  tail -1 gpurun_out/prof_r1_$tgt.log
done
ls -la gpurun_out/*.ncu-rep | tail -4
