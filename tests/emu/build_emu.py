"""TEST INFRASTRUCTURE: build `libcontrad_b200_emu.so` - the asm-free kernel files of contrad_b200/csrc compiled for the
HOST against the CUDA emulation header tests/emu/cuda_runtime.h, exporting the same C ABI as the product library for the
entry points those files define.  Launches `k<<<cfg>>>(args)` and `extern __shared__` declarations are rewritten textually;
nothing else in the kernel source is touched."""
import ctypes
import hashlib
import os
import re
import shutil
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(REPO, "contrad_b200", "csrc")

# everything except the tcgen05 / TMA-descriptor files (tc_gemm.cu, tc_wgrad.cu); the bulk-copy + mbarrier helpers of
# common.cuh have host equivalents, so the persistent augmentation and first-layer kernels are included
EMULATED = ("core.cu", "adam.cu", "augment.cu", "augment_aux.cu", "augment_hq.cu", "conv_first.cu", "gen_ops.cu", "losses.cu",
            "sg2_ops.cu", "sn_weights.cu")

_LAUNCH = re.compile(r"([A-Za-z_]\w*(?:<[^<>;(){}]*>)?)\s*<<<(.*?)>>>\s*\(", re.S)
_EXTERN_SMEM = re.compile(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?(\w+)\s+(\w+)\[\];")


def _matching_paren(text, start):
    depth = 0
    for i in range(start, len(text)):
        if text[i] == "(":
            depth += 1
        elif text[i] == ")":
            depth -= 1
            if depth == 0:
                return i
    raise ValueError("unbalanced launch at %d" % start)


def transform(text):
    """k<<<g, b, s, st>>>(args);  ->  emu::launch(emu::cfg(g, b, s, st), [&]() { k(args); });"""
    out, pos = [], 0
    while True:
        m = _LAUNCH.search(text, pos)
        if not m:
            out.append(text[pos:])
            break
        close = _matching_paren(text, m.end() - 1)
        out.append(text[pos:m.start()])
        out.append("emu::launch(emu::cfg(%s), [&]() { %s(%s); })" % (m.group(2), m.group(1), text[m.end():close]))
        pos = close + 1
    text = "".join(out)
    return _EXTERN_SMEM.sub(r"\1* \2 = reinterpret_cast<\1*>(emu::dyn_smem());", text)


def build(cache_dir=None):
    srcs = [open(os.path.join(CSRC, f)).read() for f in EMULATED]
    deps = srcs + [open(os.path.join(CSRC, "common.cuh")).read(), open(os.path.join(HERE, "cuda_runtime.h")).read(),
                   open(os.path.abspath(__file__)).read()]
    tag = hashlib.sha1("\0".join(deps).encode()).hexdigest()[:16]
    cache_dir = cache_dir or os.path.join(tempfile.gettempdir(), "contrad_b200_emu")
    os.makedirs(cache_dir, exist_ok=True)
    so = os.path.join(cache_dir, "libcontrad_b200_emu_%s.so" % tag)
    if os.path.exists(so):
        return so
    objs, procs = [], []
    work = tempfile.mkdtemp(prefix="build_%d_" % os.getpid(), dir=cache_dir)     # private intermediates: concurrent builds
    for name, text in zip(EMULATED, srcs):
        cpp = os.path.join(work, "%s.cpp" % name[:-3])
        with open(cpp, "w") as f:
            f.write(transform(text))
        obj = cpp[:-4] + ".o"
        objs.append(obj)
        cmd = ["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-fPIC", "-w", "-pthread", "-I", HERE, "-I", CSRC,
               "-I", os.path.join(REPO, "include"), "-c", cpp, "-o", obj]
        procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for name, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("g++ failed for the emulated %s:\n%s" % (name, out.decode()[-4000:]))
    tmp_so = os.path.join(work, "lib.so")
    r = subprocess.run(["g++", "-shared", "-pthread", "-o", tmp_so] + objs, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s" % r.stdout.decode())
    os.replace(tmp_so, so)                                  # atomic publish into the shared cache
    shutil.rmtree(work, ignore_errors=True)
    return so


def load():
    lib = ctypes.CDLL(build())
    lib.cb200_last_error.restype = ctypes.c_char_p
    return lib


if __name__ == "__main__":
    print(build())
