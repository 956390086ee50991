#!/bin/bash
# tap-GEMM bottleneck experiments: where does the time go (TMA supply / MMA / epilogue)?
for pair in 0 1; do
 for dbg in 0 1 2 3 6 10; do
  if [ $pair = 1 ] && [ $dbg -ge 2 ]; then continue; fi
  echo "== pair=$pair debug=$dbg (1 no-store, 2 no-mma, 4 no-A, 8 no-B)"
  CB200_TAPGEMM_PAIR=$pair CB200_TAPGEMM_DEBUG=$dbg timeout 120 python tools/gpu_probe_r1a.py 2>&1 | grep -E "conv fwd|heads"
 done
done
