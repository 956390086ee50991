"""Mirror of training/criterion.py: ``nt_xent`` on the fused sm_100a kernels."""
import torch

from ..functional import contrastive_loss, RowNormalizeFn
from ..third_party.gather_layer import GatherLayer


def nt_xent(out1, out2, temperature=0.1, distributed=False, normalize=False):
    """NT-Xent loss (training/criterion.py:24-45).  out1, out2: [N, 128]."""
    assert out1.size(0) == out2.size(0)
    if normalize:
        out1 = RowNormalizeFn.apply(out1)
        out2 = RowNormalizeFn.apply(out2)
    if distributed:
        out1 = torch.cat(GatherLayer.apply(out1), dim=0)
        out2 = torch.cat(GatherLayer.apply(out2), dim=0)
    n = out1.size(0)
    return contrastive_loss(torch.cat([out1, out2], dim=0), n, 0, float(temperature))


def target_nll_loss(inputs, targets, reduction="none"):
    raise NotImplementedError("target_nll_loss is unused by every reference script (SURVEY 2.1) and not built")
