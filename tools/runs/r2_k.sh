#!/bin/bash
# round 2, call K (1 GPU): call J's list + first-layer kernels (A/B, ncu)
# (in the same call, first: strict StyleGAN2 test, dropin worker test, tensor-core contrastive path, config-4 reference leg)
echo "== conv_first (default)"; timeout 120 python tools/bench_conv_first.py
echo "== conv_first wgrad variant 2"; CB200_CONV_FIRST_WGRAD=2 timeout 120 python tools/bench_conv_first.py
echo "== conv_first fwd 4 CTAs/SM"; CB200_CONV_FIRST_FWD_CTAS=4 timeout 120 python tools/bench_conv_first.py
echo "== conv_first parity"; timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "conv_first or first" 2>&1 | tail -2
for k in conv_first_fwd conv_first_wgrad; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -s 2 -o gpurun_out/prof_r2_$k -f python tools/profile_target.py $k > gpurun_out/prof_r2_$k.log 2>&1; tail -1 gpurun_out/prof_r2_$k.log
done
