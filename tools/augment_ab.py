#!/usr/bin/env python
"""A/B of the two builds of the fused SimCLR forward kernel (CB200_AUGMENT_V=1 | 2, csrc/augment.cu) on the GPU:
bit-equality of the outputs (reference fixtures, B = 1536 / 65536 at 32x32 with launch-wide and per-image jitter order,
64x64), then interleaved CUDA-event timings at the HBM-saturating size (8 algorithmic bytes per element, against the
measured copy bandwidth) and at the train step's size.  Writes gpurun_out/augment_ab.json."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tools"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from contrad_b200 import kernels as K  # noqa: E402
from _params import random_simclr_params  # noqa: E402


def run(variant, x, p, order):
    os.environ["CB200_AUGMENT_V"] = str(variant)
    try:
        return K.augment_simclr_fwd(x, p, order)
    finally:
        os.environ.pop("CB200_AUGMENT_V", None)


def main():
    out = {"equal": {}, "timing": {}}
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    peak = json.load(open(path))["hbm_gbs"] if os.path.exists(path) else 6554.9

    def dump():
        os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
        with open(os.path.join(REPO, "gpurun_out", "augment_ab.json"), "w") as f:
            json.dump(out, f, indent=1, sort_keys=True)

    # ---- parity: fixtures of the unmodified reference chain, then bit-equality between the builds
    fx = torch.load(os.path.join(REPO, "tests", "golden", "augment_simclr.pt"), weights_only=False)
    worst = 0.0
    for case in fx["cases"]:
        x, p = case["x"].cuda(), case["params"].cuda()
        y1, y2 = run(1, x, p, case["order"]), run(2, x, p, case["order"])
        out["equal"]["fixture_%dx%d" % tuple(x.shape[-2:])] = bool(torch.equal(y1, y2))
        worst = max(worst, float((y2.cpu() - case["y"]).abs().max()))
    out["fixture_max_abs_err_v2"] = worst
    for B, S in ((1, 32), (889, 32), (1536, 32), (65536, 32), (4096, 64)):
        x = torch.rand(B, 3, S, S, device="cuda")
        p, order = random_simclr_params(B, seed=B)
        ok = torch.equal(run(1, x, p, order), run(2, x, p, order))
        row = torch.cat([p, (torch.rand(1, B, device="cuda") < 0.5).float()])
        ok_row = torch.equal(run(1, x, row, -1), run(2, x, row, -1))
        out["equal"]["B%d_S%d" % (B, S)] = bool(ok and ok_row)
    torch.cuda.synchronize()
    dump()

    # ---- timing, builds interleaved
    for B, S in ((65536, 32), (1536, 32), (16384, 64)):
        x = torch.rand(B, 3, S, S, device="cuda")
        p, order = random_simclr_params(B)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if B < 8192 else None
        evs = {1: [], 2: []}
        for rep in range(13):
            for v in (1, 2):
                if flush is not None:
                    flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                os.environ["CB200_AUGMENT_V"] = str(v)
                e0.record(); K.augment_simclr_fwd(x, p, order); e1.record()
                if rep >= 3:
                    evs[v].append((e0, e1))
        os.environ.pop("CB200_AUGMENT_V", None)
        torch.cuda.synchronize()
        for v in (1, 2):
            ms = float(np.median([a.elapsed_time(b) for a, b in evs[v]]))
            gbs = 8.0 * x.numel() / (ms * 1e-3) / 1e9
            out["timing"]["B%d_S%d_v%d" % (B, S, v)] = {"ms": ms, "gbs": gbs, "frac": gbs / peak}
        dump()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
