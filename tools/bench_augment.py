#!/usr/bin/env python
"""Fused SimCLR augmentation at the HBM-saturating size (B = 65536 images of 32x32: 805 MB in, 805 MB out): median
CUDA-event time of 10 launches, achieved GB/s at 8 algorithmic bytes per element, fraction of the measured copy bandwidth."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tools"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from contrad_b200 import kernels as K  # noqa: E402
from _params import random_simclr_params  # noqa: E402


def main():
    peak = 6554.9
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        peak = json.load(open(path))["hbm_gbs"]
    out = {}
    for B, S in ((65536, 32), (16384, 64)):
        x = torch.rand(B, 3, S, S, device="cuda")
        p, order = random_simclr_params(B)
        for _ in range(3):
            K.augment_simclr_fwd(x, p, order)
        evs = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); K.augment_simclr_fwd(x, p, order); e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        ms = float(np.median([a.elapsed_time(b) for a, b in evs]))
        gbs = 8.0 * x.numel() / (ms * 1e-3) / 1e9
        out["fwd_%d" % S] = {"ms": ms, "gbs": gbs, "frac": gbs / peak}
        if S == 32:
            dy = torch.rand_like(x)
            for _ in range(3):
                K.augment_simclr_bwd(x, dy, p, order)
            evs = []
            for _ in range(10):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); K.augment_simclr_bwd(x, dy, p, order); e1.record()
                evs.append((e0, e1))
            torch.cuda.synchronize()
            ms = float(np.median([a.elapsed_time(b) for a, b in evs]))
            gbs = 12.0 * x.numel() / (ms * 1e-3) / 1e9
            out["bwd_32"] = {"ms": ms, "gbs": gbs, "frac": gbs / peak}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
