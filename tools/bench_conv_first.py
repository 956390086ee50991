#!/usr/bin/env python
"""First discriminator layer (3 -> 64, SIMT) at the D-step batch: forward and weight-gradient kernel times, achieved GB/s
at the algorithmic bytes (forward 12 B/px in + 256 B/px out; wgrad 12 + 256 B/px in), fraction of the measured copy bandwidth."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from contrad_b200 import kernels as K  # noqa: E402


def med(fn, reps=10):
    for _ in range(3):
        fn()
    evs = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in evs]))


def main():
    peak = 6554.9
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        peak = json.load(open(path))["hbm_gbs"]
    out = {}
    for B in (1536, 512, 192):
        xs = [torch.rand(B, 3, 32, 32, device="cuda") for _ in range(3)]
        w = torch.randn(64, 3, 3, 3, device="cuda") * 0.1
        bias = torch.zeros(64, device="cuda")
        dys = [K.round_tf32(torch.randn(B, 32, 32, 64, device="cuda")) for _ in range(3)]
        i = [0]

        def fwd():
            i[0] += 1
            K.conv_first_fwd(xs[i[0] % 3], w, None, bias)

        def wgrad():
            i[0] += 1
            K.conv_first_wgrad(xs[i[0] % 3], dys[i[0] % 3])

        nbytes = B * 1024 * (12 + 256)
        tf, tw = med(fwd), med(wgrad)
        out["B%d" % B] = {"fwd_ms": tf, "fwd_gbs": nbytes / tf / 1e6, "fwd_frac": nbytes / tf / 1e6 / peak,
                          "wgrad_ms": tw, "wgrad_gbs": nbytes / tw / 1e6, "wgrad_frac": nbytes / tw / 1e6 / peak}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
