"""TEST INFRASTRUCTURE: run the asm-free kernels of contrad_b200/csrc on the host (see cuda_runtime.h, build_emu.py).

`emulated()` routes the product's raw bindings (contrad_b200.kernels / sg2_kernels) to the host-compiled library for the
duration of a test, so the bindings, the C launchers (grid / shared-memory sizing, argument checks) and the kernel source
itself are exercised on CPU tensors.  Entry points that live in PTX files (tcgen05 GEMMs, the bulk-copy augmentation
kernels) are not in the emulated library; touching them raises AttributeError."""
import contextlib
import ctypes

from . import build_emu


class _Missing(AttributeError):
    pass


@contextlib.contextmanager
def emulated():
    from contrad_b200 import _capi, kernels, sg2_kernels
    lib = build_emu.load()

    def ptr(t):
        if t is None:
            return ctypes.c_void_p(0)
        if t.is_cuda:
            raise RuntimeError("emulated kernels take CPU tensors")
        return ctypes.c_void_p(t.data_ptr())

    patched = {"lib": lambda: lib, "ptr": ptr, "stream_ptr": lambda: ctypes.c_void_p(0)}
    saved = []
    for mod in (kernels, sg2_kernels):
        for name, fn in patched.items():
            saved.append((mod, name, getattr(mod, name)))
            setattr(mod, name, fn)
    saved.append((_capi, "_lib", _capi._lib))
    _capi._lib = lib                      # last_error() / launch_count() of the emulated library
    try:
        yield lib
    finally:
        for mod, name, fn in saved:
            setattr(mod, name, fn)
