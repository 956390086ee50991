"""Roofline points of the kernels of rows f3 / f4 (uint8-source augmentation, hfrt gather, Gaussian noise): algorithmic
bytes / CUDA-event time, operands rotated so inputs are cold, against the measured copy bandwidth.
    python tools/bench_next_rows.py > gpurun_out/next_rows_kernels.json"""
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tools"))
from contrad_b200 import kernels as K          # noqa: E402
from _params import random_shift_flip_params, random_simclr_params      # noqa: E402


def timeit(fn, sets, reps=6):
    for a in sets:
        fn(*a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(reps):
        fn(*sets[r % len(sets)])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    peak = 6554.9
    try:
        peak = float(json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:                                       # noqa: BLE001
        pass
    out = {"peak_gbps": peak, "kernels": []}

    def add(name, shape, ms, nbytes):
        gbps = nbytes / ms / 1e6
        out["kernels"].append({"kernel": name, "shape": shape, "ms": ms, "algorithmic_bytes": nbytes, "gbps": gbps,
                               "frac": gbps / peak})

    np.random.seed(0); torch.manual_seed(0)
    # ---- mixed-source augmentation: n uint8 images x 2 views + n fp32 images (the D-step batch of contrad), 32x32
    for n, size in ((16384, 32), (512, 32), (4096, 64)):
        total = 3 * n
        packed, order = random_simclr_params(total)
        sets = [(torch.randint(0, 256, (n, 3, size, size), dtype=torch.uint8, device="cuda"),
                 torch.rand(n, 3, size, size, device="cuda")) for _ in range(3)]
        ms = timeit(lambda u, f: K.augment_simclr_mixed_fwd(u, 2 * n, f, packed, order), sets)
        elems = 3 * size * size
        # reads: n uint8 images once (second view hits L2 at best; counted once) + n fp32; writes 3n fp32
        add("augment_simclr_mixed_fwd", [n, 3, size, size], ms, n * elems * 1 + n * elems * 4 + total * elems * 4)
        cat_sets = [(torch.rand(total, 3, size, size, device="cuda"),) for _ in range(3)]
        ms2 = timeit(lambda c: K.augment_simclr_fwd(c, packed, order), cat_sets)
        add("augment_simclr_fwd (fp32 cat, same views)", [total, 3, size, size], ms2, total * elems * 8)
    # ---- hfrt gather
    for shape in ((16384, 3, 32, 32), (1536, 3, 32, 32), (48, 3, 512, 512)):
        b, _, h, w = shape
        prm = random_shift_flip_params(b, 4, w)
        sets = [(torch.rand(*shape, device="cuda"),) for _ in range(3)]
        nel = int(np.prod(shape))
        add("shift_flip_fwd", list(shape), timeit(lambda x: K.shift_flip(x, prm, "reflection"), sets), 8 * nel)
        add("shift_flip_bwd", list(shape), timeit(lambda x: K.shift_flip(x, prm, "reflection", adjoint=True), sets), 8 * nel)
    # ---- Gaussian noise
    shape = (16384, 3, 32, 32)
    sets = [(torch.rand(*shape, device="cuda"), torch.randn(*shape, device="cuda"), torch.randn(*shape, device="cuda"))
            for _ in range(3)]
    nel = int(np.prod(shape))
    add("noise_clamp_fwd", list(shape), timeit(lambda x, z, g: K.noise_clamp_fwd(x, z, 0.12), sets), 12 * nel)
    add("noise_clamp_bwd", list(shape), timeit(lambda x, z, g: K.noise_clamp_bwd(x, z, g, 0.12), sets), 16 * nel)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
