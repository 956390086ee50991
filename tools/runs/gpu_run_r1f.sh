#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "augment" 2>&1 | tail -15 > gpurun_out/pytest_r1f_aug.log; cat gpurun_out/pytest_r1f_aug.log | tail -5
timeout 400 python -m pytest tests/test_gpu_model.py -m gpu -q -s 2>&1 > gpurun_out/pytest_r1f.log; grep -E "passed|failed|^E  |G grad-norm|step [12]:|AssertionError" gpurun_out/pytest_r1f.log | cut -c1-700
python - <<'PY'
import sys, torch, numpy as np
sys.path.insert(0, '.')
from contrad_b200 import kernels as K
from oracle import contrad_oracle as O
def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for s, e in evs:
        s.record(); fn(); e.record()
    torch.cuda.synchronize()
    ts = sorted(s.elapsed_time(e) for s, e in evs)
    return ts[len(ts) // 2]
for B, size in ((65536, 32), (8192, 64)):
    np.random.seed(0); torch.manual_seed(0)
    params, order = O.sample_simclr_params(B, size, size)
    p = O.pack_params(params).cuda()
    x = torch.rand(B, 3, size, size, device="cuda"); dy = torch.randn_like(x)
    for od in (0, 1):
        ms = timeit(lambda: K.augment_simclr_fwd(x, p, od)); print("aug fwd B=%d s=%d o=%d %.3f ms %.0f GB/s" % (B, size, od, ms, 8 * x.numel() / ms / 1e6))
        ms = timeit(lambda: K.augment_simclr_bwd(x, dy, p, od)); print("aug bwd B=%d s=%d o=%d %.3f ms %.0f GB/s(12B)" % (B, size, od, ms, 12 * x.numel() / ms / 1e6))
PY
