#!/bin/bash
mkdir -p gpurun_out
nproc; cat /proc/loadavg; python - <<'PY'
import time, numpy as np
t=time.perf_counter(); x=0
for i in range(2000000): x+=i
print("python loop 2M: %.3f s" % (time.perf_counter()-t))
PY
timeout 200 python tools/host_profile.py 2>&1 | head -45
timeout 300 python tools/debug_graph.py 2>&1 | grep -v Warning | tail -32
