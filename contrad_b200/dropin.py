"""Run the UNMODIFIED reference training scripts on the B200 path.

    python -m contrad_b200.dropin /path/to/ContraD/train_gan.py configs/gan/cifar10/c10_b512.gin sndcgan \\
           --mode=contrad --aug=simclr --use_warmup

``install()`` registers this package's mirrors under the reference's own module names (``augment``,
``training``, ``training.criterion``, ``training.gan``, ``training.gan.contrad``,
``third_party.gather_layer``, ``models``, ``models.gan`` ..., ``penalty``) so that the script's
``from augment import get_augment`` / ``from training.gan import setup`` / ``from models.gan import
get_architecture`` resolve to the sm_100a implementations, and makes ``gin`` / ``tensorboardX`` /
``imageio`` / ``kornia`` importable through the shims in ``contrad_b200/compat`` when the real packages
are absent.  Everything else the script imports (``utils``, ``datasets``, ``evaluate``) is the
reference's own host glue and is left alone (harness, runs unchanged).
"""
import importlib
import importlib.util
import os
import runpy
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_COMPAT = os.path.join(_HERE, "compat")

_ALIASES = {
    "augment": "contrad_b200.augment",
    "augment.layers": "contrad_b200.augment.layers",
    # the reference's augment submodules: their classes live in one module here (`from augment.spatial import CutOut`,
    # `from augment.color_jitter import ColorJitterLayer` keep working; `augment.utils.rgb2hsv / hsv2rgb` are not mirrored -
    # the colour-space round trip only exists inside the fused kernel)
    "augment.spatial": "contrad_b200.augment.layers",
    "augment.color_jitter": "contrad_b200.augment.layers",
    "training": "contrad_b200.training",
    "training.criterion": "contrad_b200.training.criterion",
    "training.gan": "contrad_b200.training.gan",
    "training.gan.contrad": "contrad_b200.training.gan.contrad",
    "training.gan.std": "contrad_b200.training.gan.std",
    "training.gan.aug": "contrad_b200.training.gan.aug",
    "training.gan.aug_both": "contrad_b200.training.gan.aug_both",
    "training.gan.simclr_only": "contrad_b200.training.gan.simclr_only",
    "third_party.gather_layer": "contrad_b200.third_party.gather_layer",
    "models": "contrad_b200.models",
    "models.gan": "contrad_b200.models.gan",
    "models.gan.base": "contrad_b200.models.gan.base",
    "models.gan.sndcgan": "contrad_b200.models.gan.sndcgan",
    "models.gan.snresnet": "contrad_b200.models.gan.snresnet",
    "models.gan.stylegan2": "contrad_b200.models.gan.stylegan2",
    "models.gan.stylegan2.layers": "contrad_b200.models.gan.stylegan2.layers",
    "models.gan.stylegan2.generator": "contrad_b200.models.gan.stylegan2.generator",
    "models.gan.stylegan2.discriminator": "contrad_b200.models.gan.stylegan2.discriminator",
    "models.gan.stylegan2.op": "contrad_b200.models.gan.stylegan2.op",
    "models.gan.stylegan2.op.fused_act": "contrad_b200.models.gan.stylegan2.op.fused_act",
    "models.gan.stylegan2.op.upfirdn2d": "contrad_b200.models.gan.stylegan2.op.upfirdn2d",
    "penalty": "contrad_b200.penalty",
}


def install(shims=True):
    """Idempotent.  Returns the list of module names that now resolve to contrad_b200."""
    if shims:
        for name in ("gin", "tensorboardX", "imageio", "kornia"):
            if importlib.util.find_spec(name) is None and _COMPAT not in sys.path:
                sys.path.append(_COMPAT)
    installed = []
    for alias, target in _ALIASES.items():
        sys.modules[alias] = importlib.import_module(target)
        installed.append(alias)
    # `third_party` stays the reference's own package (fid, inception, ...); only gather_layer is replaced
    tp = sys.modules.get("third_party")
    if tp is not None:
        tp.gather_layer = sys.modules["third_party.gather_layer"]
    return installed


def _wrap_reference_accumulate():
    """The reference's `utils.accumulate` (utils.py:130-143) writes the EMA generator through `.data` in-place ops that
    no version counter records; wrap it so that derived-weight memos (sg2_functional.weight_memo) are dropped after
    every call.  (No-grad forwards already bypass the memo; this also covers a g_ema evaluated with grad enabled.)"""
    try:
        utils = importlib.import_module("utils")
    except Exception:                                            # noqa: BLE001 - no reference `utils` on the path
        return False
    fn = getattr(utils, "accumulate", None)
    if fn is None or getattr(fn, "_cb200_wrapped", False):
        return False
    from . import sg2_functional

    def accumulate(*args, **kwargs):
        out = fn(*args, **kwargs)
        sg2_functional.bump_weight_epoch()
        return out

    accumulate._cb200_wrapped = True
    accumulate.__doc__ = fn.__doc__
    utils.accumulate = accumulate
    return True


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit(__doc__)
    script = os.path.abspath(argv[0])
    install()
    sys.argv = [script] + argv[1:]
    sys.path.insert(0, os.path.dirname(script))
    os.chdir(os.path.dirname(script))
    _wrap_reference_accumulate()
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
