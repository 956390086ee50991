#!/bin/bash
# round 2, call T (1 GPU, the last ~90 s of the budget): A/B of the two builds of the fused augment forward kernel
mkdir -p gpurun_out
timeout 50 python tools/augment_ab.py > gpurun_out/augment_ab.log 2>&1; echo "rc=$?"; tail -c 1800 gpurun_out/augment_ab.log
CB200_AUGMENT_V=2 timeout 45 python -m pytest tests/test_gpu_kernels.py -q -x -k "augment_builds or augment_matches_reference_golden or per_image_order_row" 2>&1 | tail -3
