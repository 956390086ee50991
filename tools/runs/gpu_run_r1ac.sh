#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E|passed|failed|^FAILED" | head -20
for mode in 1 0; do
CB200_FUSED_COLSUM=$mode timeout 300 python bench.py --steps 30 --warmup 8 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('fused_colsum=$mode', d['value'], d['ms_per_step'], 'dgrad share', d['kernel_time_share'].get('conv2d_nhwc_dgrad'), 'colsum', d['kernel_time_share'].get('colsum'), 'dgrad TF', d['roofline']['achieved'])"
done
for tgt in conv_first_wgrad conv_first_fwd; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$tgt -s 2 -c 1 -f -o gpurun_out/prof_r1_$tgt python tools/profile_target.py $tgt > gpurun_out/prof_r1_$tgt.log 2>&1
  tail -1 gpurun_out/prof_r1_$tgt.log
done
