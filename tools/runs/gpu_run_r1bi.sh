#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_kernels.py -q -x -k "augment" 2>&1 | grep -E "^E|passed|failed|^FAILED|^ERROR|Error" | head -20
timeout 600 python -m pytest tests/test_gpu_sg2.py -q -x -k "512 or upfirdn or layout" 2>&1 | grep -E "^E|passed|failed|^FAILED|^ERROR|Error" | head -20
python - <<'PY'
import json
d=json.load(open("gpurun_out/sg2_parity.json")) if False else None
PY
