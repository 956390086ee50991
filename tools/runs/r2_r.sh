#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_sg2.py -x -q -k "config4_dstep or config5_per_replica" 2>&1 | grep -E "config4_b64|config5_replica|passed|failed|Error|assert" | cut -c1-700 | head
