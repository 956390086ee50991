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

Select with `contrad_b200.precision.set_strict(True)`, the context manager `strict()`, or CB200_PRECISION=strict."""
import contextlib
import os

_STRICT = os.environ.get("CB200_PRECISION", "").strip().lower() in ("strict", "3xtf32", "tf32x3")


def strict_enabled():
    return _STRICT


def set_strict(flag=True):
    global _STRICT
    _STRICT = bool(flag)


@contextlib.contextmanager
def strict(flag=True):
    global _STRICT
    old = _STRICT
    _STRICT = bool(flag)
    try:
        yield
    finally:
        _STRICT = old
