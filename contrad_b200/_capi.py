"""ctypes loader for ``libcontrad_b200.so`` -- the C ABI declared in ``include/contrad_b200.h``.

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a only).  There is no
CPU fallback: if the shared library is missing, or a kernel is asked to run on something that is
not a CUDA tensor, the call raises.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcontrad_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "contrad_b200.h")

_lib = None


class CB200Error(RuntimeError):
    pass


def declared_symbols():
    """Every function name declared in include/contrad_b200.h."""
    with open(HEADER_PATH) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cb200_[a-z0-9_]+)\s*\(", text)))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CB200Error(
                "contrad_b200: %s not found - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)" % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.cb200_last_error.restype = ctypes.c_char_p
        _lib.cb200_launch_count.restype = ctypes.c_ulonglong
        _lib.cb200_reset_launch_count.restype = None
        _lib.cb200_add_launch_count.restype = None
        _lib.cb200_add_launch_count.argtypes = [ctypes.c_longlong]
    return _lib


def last_error():
    msg = lib().cb200_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc, what):
    if rc != 0:
        raise CB200Error("%s failed (code %d): %s" % (what, rc, last_error()))


def launch_count():
    return int(lib().cb200_launch_count())


def reset_launch_count():
    lib().cb200_reset_launch_count()


def add_launch_count(n):
    lib().cb200_add_launch_count(int(n))


def ptr(t):
    """Device pointer of a CUDA tensor as c_void_p (None -> NULL)."""
    if t is None:
        return ctypes.c_void_p(0)
    if not t.is_cuda:
        raise CB200Error("contrad_b200 kernels need CUDA tensors (got %s); there is no CPU path" % t.device)
    return ctypes.c_void_p(t.data_ptr())


_raw_stream = None


def stream_ptr():
    """cudaStream_t of torch's current stream on the current device (raw C accessor: ~1 us instead of ~18 us)."""
    global _raw_stream
    import torch
    if _raw_stream is None:
        if not torch.cuda.is_available():
            raise CB200Error("contrad_b200 kernels need a CUDA device; there is no CPU path")
        _raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None) or (
            lambda dev: torch.cuda.current_stream(dev).cuda_stream)
    return ctypes.c_void_p(_raw_stream(torch.cuda.current_device()))


def i32(v):
    return ctypes.c_int(int(v))


def i64(v):
    return ctypes.c_longlong(int(v))


def f32(v):
    return ctypes.c_float(float(v))
