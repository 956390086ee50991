"""Fused Adam on the sm_100a kernels (SURVEY "next" row f2).  API- and state_dict-compatible with
``torch.optim.Adam`` as the reference configures it (train_gan.py:273-274: lr, betas, eps=1e-8, no weight decay,
no amsgrad), so ``optim.pt`` checkpoints interchange.  One kernel launch per ``step()``.

Under ``staging.Recorder`` (CUDA-graph capture of the train step) the step-dependent scalars - learning rate and
the two bias corrections - are staged as a 3-float device tensor, so the captured launch stays valid for every
later step; the host-side ``state['step']`` counters still advance once per replay (inside the staged producer)."""
import math

import torch

from . import kernels as K
from . import staging


def _grad_operand(p):
    """The gradient as the kernel wants it: contiguous and 16-byte aligned (float4 loads).  nn.DataParallel hands the
    master parameters gradients that are views into its coalesced reduce buffers at arbitrary offsets - those are
    copied once; everything the library itself produces is already aligned."""
    g = p.grad
    if not g.is_contiguous() or g.data_ptr() % 16:
        g = g.contiguous() if not g.is_contiguous() else g.clone()
    return g


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1:
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, maximize=False,
                                      foreach=None, capturable=False, differentiable=False, fused=None))

    def _init_state(self, p):
        state = self.state[p]
        if len(state) == 0:
            state["step"] = torch.tensor(0.0)
            state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return state

    def _step_recording(self, group):
        """Capture-safe variant: one launch for the group, scalars from a staged device tensor."""
        params = [p for p in group["params"] if p.grad is not None]
        if not params:
            return
        states = [self._init_state(p) for p in params]
        if len({int(s["step"]) for s in states}) != 1:
            raise RuntimeError("FusedAdam: graph capture needs all tensors of a group at the same step count")

        def hyper():
            for s in states:
                s["step"] += 1
            t = int(states[0]["step"])
            b1, b2 = group["betas"]
            return torch.tensor([group["lr"], 1.0 - b1 ** t, math.sqrt(1.0 - b2 ** t)], dtype=torch.float32)

        dev_hyper = staging.stage(hyper, params[0].device, shape=(3,), late=True)
        entries = [(p, _grad_operand(p), s["exp_avg"], s["exp_avg_sq"]) for p, s in zip(params, states)]
        K.adam_step(entries, dev_hyper, group["betas"][0], group["betas"][1], group["eps"], 0)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        from . import sg2_functional
        sg2_functional.bump_weight_epoch()         # the kernel writes the parameters behind torch's version counters
        for group in self.param_groups:
            if staging.recording():
                self._step_recording(group)
                continue
            entries, step_no = [], None
            for p in group["params"]:
                if p.grad is None:
                    continue
                state = self._init_state(p)
                state["step"] += 1
                s = int(state["step"])
                if step_no is None:
                    step_no = s
                if s != step_no:                     # tensors with a different history: separate launch
                    K.adam_step([(p, _grad_operand(p), state["exp_avg"], state["exp_avg_sq"])], group["lr"],
                                group["betas"][0], group["betas"][1], group["eps"], s)
                    continue
                entries.append((p, _grad_operand(p), state["exp_avg"], state["exp_avg_sq"]))
            if entries:
                K.adam_step(entries, group["lr"], group["betas"][0], group["betas"][1], group["eps"], step_no)
        return loss
