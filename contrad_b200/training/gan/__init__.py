"""Mirror of training/gan/__init__.py:4-29 - the reference's training-mode plug-in boundary.  mode='contrad' is the hot
path; the baselines (std / aug / aug_both / simclr_only) are thin compositions of the same kernels (row f4)."""
from importlib import import_module

_FILENAMES = {
    "std": lambda P: f"{P.mode}_{P.penalty}" + (f"_{P.aug}" if "cr" in P.penalty else ""),
    "aug": lambda P: f"{P.mode}_{P.aug}_{P.penalty}",
    "aug_both": lambda P: f"{P.mode}_{P.aug}_{P.penalty}",
    "simclr_only": lambda P: f"{P.mode}_{P.aug}_T{P.temp}",
    "contrad": lambda P: f"{P.mode}_{P.aug}_L{P.lbd_a}_T{P.temp}",
}


def setup(P):
    if P.mode not in _FILENAMES:
        raise NotImplementedError()
    mod = import_module(f".{P.mode}", __name__)
    P.filename = _FILENAMES[P.mode](P)
    P.train_fn = {"G": mod.loss_G_fn, "D": mod.loss_D_fn}
    return P
