"""Make the *unmodified* reference importable in the build container (TEST INFRASTRUCTURE).

The reference is pure Python but needs `gin`, `tensorboardX`, `kornia`, `imageio`, none of
which are installed; the product ships faithful shims for them under
``contrad_b200/compat`` (they are part of the drop-in boundary).  This helper puts
those shims and ``/root/reference`` on ``sys.path`` so that ``import augment``,
``import training.gan.contrad`` ... resolve to the reference's own files.  It is used only by
``tests/golden/make_golden.py`` (fixture generation) and by CPU tests that are skipped when
the reference is absent (it does not exist on the GPU box).
"""
import importlib
import os
import sys

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_VENDORED = os.path.join(_REPO, "oracle", "_ref")        # unmodified copy made by oracle/make_ref.py (travels to the GPU box)


def _find_root():
    env = os.environ.get("CONTRAD_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", _VENDORED):
        if os.path.isdir(os.path.join(cand, "augment")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()
_COMPAT = os.path.join(_REPO, "contrad_b200", "compat")

_REF_TOP_LEVEL = ("augment", "training", "third_party", "models", "penalty", "utils", "datasets", "evaluate")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "augment"))


def activate(gin_files=("configs/defaults/gan.gin", "configs/defaults/augment.gin")):
    """Put the reference first on sys.path, (re)load gin defaults, return the gin module."""
    if not reference_available():
        raise RuntimeError("reference not found at %s" % REFERENCE_ROOT)
    for mod in list(sys.modules):
        if mod.split(".")[0] in _REF_TOP_LEVEL:
            origin = getattr(sys.modules[mod], "__file__", "") or ""
            if not origin.startswith(REFERENCE_ROOT):
                del sys.modules[mod]
    for path in (_COMPAT, REFERENCE_ROOT):
        if path in sys.path:
            sys.path.remove(path)
    sys.path.insert(0, _COMPAT)
    sys.path.insert(0, REFERENCE_ROOT)
    gin = importlib.import_module("gin")
    gin.clear_config()
    for f in gin_files:
        gin.parse_config_file(os.path.join(REFERENCE_ROOT, f))
    return gin


def deactivate():
    for mod in list(sys.modules):
        if mod.split(".")[0] in _REF_TOP_LEVEL:
            origin = getattr(sys.modules[mod], "__file__", "") or ""
            if origin.startswith(REFERENCE_ROOT):
                del sys.modules[mod]
    for path in (_COMPAT, REFERENCE_ROOT):
        while path in sys.path:
            sys.path.remove(path)
