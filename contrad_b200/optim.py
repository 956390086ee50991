"""Fused Adam on the sm_100a kernels (SURVEY "next" row f2).  API- and state_dict-compatible with
``torch.optim.Adam`` as the reference configures it (train_gan.py:273-274: lr, betas, eps=1e-8, no weight decay,
no amsgrad), so ``optim.pt`` checkpoints interchange.  One kernel launch per ``step()``."""
import torch

from . import kernels as K


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1:
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, maximize=False,
                                      foreach=None, capturable=False, differentiable=False, fused=None))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            entries, step_no = [], None
            for p in group["params"]:
                if p.grad is None:
                    continue
                state = self.state[p]
                if len(state) == 0:
                    state["step"] = torch.tensor(0.0)
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                state["step"] += 1
                s = int(state["step"])
                if step_no is None:
                    step_no = s
                if s != step_no:                     # tensors with a different history: separate launch
                    K.adam_step([(p, p.grad.contiguous(), state["exp_avg"], state["exp_avg_sq"])], group["lr"],
                                group["betas"][0], group["betas"][1], group["eps"], s)
                    continue
                entries.append((p, p.grad if p.grad.is_contiguous() else p.grad.contiguous(), state["exp_avg"],
                                state["exp_avg_sq"]))
            if entries:
                K.adam_step(entries, group["lr"], group["betas"][0], group["betas"][1], group["eps"], step_no)
        return loss
