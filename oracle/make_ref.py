"""Recipe that materialises the UNMODIFIED reference under ``oracle/_ref/`` (TEST / MEASUREMENT INFRASTRUCTURE).

    python oracle/make_ref.py            # copies /root/reference -> oracle/_ref (git-ignored, travels with gpurun)

The reference is pure Python (+ two JIT-built CUDA ops that only the StyleGAN2 configs touch), so "building" it is a
byte-for-byte copy of its sources and gin configs from where they lie under ``/root/reference``; nothing is edited
and nothing is committed (``oracle/_ref/`` is listed in ``.gitignore``, not in ``.gpurunignore``: it has to reach the
GPU box, where ``/root/reference`` does not exist).  ``bench.py --impl reference`` and the ``eager_gpu_baseline`` leg
import these files through ``oracle/ref_import.py`` + the ``contrad_b200/compat`` shims for the four packages the
image lacks (gin, tensorboardX, imageio, kornia).  ``__graft_entry__.build()`` runs this recipe whenever
``/root/reference`` is present.  A manifest with the sha256 of every copied file is written next to the copy so that
a reader can check that nothing was modified.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("CONTRAD_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")

# resources/ = README images (5 MB); third_party/tf = TensorFlow FID (needs tensorflow, out of scope)
SKIP_DIRS = {".git", "resources", "__pycache__", "logs", "data"}
SKIP_REL = {os.path.join("third_party", "tf")}
KEEP_EXT = {".py", ".gin", ".cpp", ".cu", ".h", ".txt", ".yml", ".md"}


def _up_to_date(src, dst):
    """True when oracle/_ref already holds an unmodified copy of every file the recipe would copy."""
    try:
        with open(os.path.join(dst, "MANIFEST.json")) as f:
            files = json.load(f)["files"]
        for rel, digest in files.items():
            with open(os.path.join(src, rel), "rb") as f:
                if hashlib.sha256(f.read()).hexdigest() != digest:
                    return False
        return verify(dst)
    except (OSError, KeyError, ValueError):
        return False


def make(src=SRC, dst=DST, quiet=False, force=False):
    if not os.path.isdir(os.path.join(src, "augment")):
        raise RuntimeError("reference sources not found at %s" % src)
    if not force and _up_to_date(src, dst):
        if not quiet:
            print("oracle/_ref: up to date")
        return dst
    tmp = dst + ".tmp"
    shutil.rmtree(tmp, ignore_errors=True)
    manifest = {}
    for root, dirs, files in os.walk(src):
        rel_root = os.path.relpath(root, src)
        dirs[:] = sorted(d for d in dirs if d not in SKIP_DIRS
                         and os.path.normpath(os.path.join(rel_root, d)) not in SKIP_REL)
        for name in sorted(files):
            if os.path.splitext(name)[1] not in KEEP_EXT and name != "LICENSE":
                continue
            rel = os.path.normpath(os.path.join(rel_root, name))
            out = os.path.join(tmp, rel)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            shutil.copyfile(os.path.join(root, name), out)
            with open(out, "rb") as f:
                manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(tmp, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "files": manifest}, f, indent=1, sort_keys=True)
    shutil.rmtree(dst, ignore_errors=True)
    os.rename(tmp, dst)
    if not quiet:
        print("oracle/_ref: %d files copied unmodified from %s" % (len(manifest), src))
    return dst


def verify(dst=DST):
    """True when every file under oracle/_ref still has the sha256 recorded at copy time."""
    with open(os.path.join(dst, "MANIFEST.json")) as f:
        files = json.load(f)["files"]
    for rel, digest in files.items():
        with open(os.path.join(dst, rel), "rb") as f:
            if hashlib.sha256(f.read()).hexdigest() != digest:
                return False
    return True


EXT_DIR = os.path.join(DST, "_torch_ext")


def build_extensions(dst=DST, quiet=False):
    """JIT-build the reference's two CUDA ops (models/gan/stylegan2/op: upfirdn2d, fused bias-act) for sm_100 with the
    reference's OWN `torch.utils.cpp_extension.load` calls, into oracle/_ref/_torch_ext, so that the StyleGAN2 reference
    legs find them pre-built on the GPU box (no compiler run inside a timed GPU lease).  ~90 s on first use, cached."""
    import subprocess
    env = dict(os.environ, TORCH_CUDA_ARCH_LIST="10.0", TORCH_EXTENSIONS_DIR=EXT_DIR, MAX_JOBS="4")
    code = ("import sys; sys.path.insert(0, %r); from oracle import ref_import; ref_import.REFERENCE_ROOT = %r; "
            "ref_import.activate(); import models.gan.stylegan2.op as op; print('reference ops:', op.upfirdn2d.__name__)"
            % (os.path.dirname(HERE), dst))
    r = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if not quiet:
        print(r.stdout.decode()[-300:])
    return r.returncode == 0


if __name__ == "__main__":
    make(force="--force" in sys.argv)
    ok = verify()
    if "--ext" in sys.argv:
        ok = build_extensions() and ok
    sys.exit(0 if ok else 1)
