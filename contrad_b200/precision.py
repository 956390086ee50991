"""Precision policy of the tensor-core GEMMs.

default   single-pass TF32 (fp32 storage, operands rounded to TF32 by their producer, fp32 accumulate) - the arithmetic
          class of the reference's own GPU execution (cuDNN TF32 convolutions, PyTorch default flags).
strict    "3xTF32" in the GENERATOR step: every GEMM / convolution of the G forward, of the frozen-D forward and of their
          data-gradient chain runs on error-compensated operands (x = hi + lo, both TF32; a_hi b_hi + a_lo b_hi + a_hi b_lo
          evaluated by the unchanged tcgen05 kernels over a concatenated reduction axis, cb200_split_tf32).  This is what
          it takes to hold the generator's gradient norm to 1e-3 of the fp32 CPU reference at initialisation: that norm is
          a small residual (BatchNorm removes the common mode of an almost constant dL/dD), and tools/tf32_sensitivity.py
          shows that rounding ANY single layer's operands to TF32 moves it by 1e-3 ... 6e-3, while the losses and the
          discriminator's gradient norm stay within 1e-3 at single-pass TF32.  Cost: +38 % tensor FLOPs per step
          (3x the G-step forward GEMMs, 2x its backward GEMMs); measured throughput in DESIGN.md.

full      the same for the discriminator step as well (G forward, D forward on 3N images, data- and weight-gradient
          GEMMs).  Needed for 1e-3 on the generator's gradient norm BEYOND the first step: the generator step runs through
          the discriminator weights the preceding Adam update produced, and Adam's first updates are ~lr * sign(grad), so
          single-pass-TF32 noise in near-zero gradient elements flips update signs (CPU emulation, two steps at n = 64:
          generator step exact + discriminator step TF32 leaves 3e-4 ... 7e-4; measured on the B200 2e-3 at step 2).
          Cost: 2.4x the default tensor FLOPs.

Select with `contrad_b200.precision.set_strict(True | "full")`, the context manager `strict(...)`, or
CB200_PRECISION=strict | full."""
import contextlib
import os

_LEVELS = {"": 0, "0": 0, "default": 0, "tf32": 0, "strict": 1, "gstep": 1, "3xtf32": 1, "tf32x3": 1, "full": 2, "strict_full": 2}
_LEVEL = _LEVELS.get(os.environ.get("CB200_PRECISION", "").strip().lower(), 0)


def _level(flag):
    if isinstance(flag, str):
        return _LEVELS[flag.strip().lower()]
    if flag is True:
        return 1
    return int(flag or 0)


def level():
    """0 = single-pass TF32, 1 = error-compensated generator step, 2 = error-compensated everywhere."""
    return _LEVEL


def strict_enabled():
    return _LEVEL >= 1


def strict_full():
    return _LEVEL >= 2


def set_strict(flag=True):
    global _LEVEL
    _LEVEL = _level(flag)


@contextlib.contextmanager
def strict(flag=True):
    global _LEVEL
    old = _LEVEL
    _LEVEL = _level(flag)
    try:
        yield
    finally:
        _LEVEL = old
