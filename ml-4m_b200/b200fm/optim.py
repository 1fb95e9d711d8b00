"""Fused AdamW on the b200fm kernel (one launch per parameter tensor: read p,g,m,v / write p,m,v in a single pass).

Same update rule and defaults as torch.optim.AdamW as configured by the reference (fourm/utils/optim_factory.py:239-240:
betas (0.9, 0.95), weight decay 0.05 with no decay on norm / bias / 1-D tensors via param groups)."""
import torch

from . import functional as BF
from . import ops


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                views = BF.shadow_views(p)
                ops.adamw_step(p.data, g, st["exp_avg"], st["exp_avg_sq"], group["lr"], b1, b2, group["eps"], group["weight_decay"],
                               st["step"], grad_scale, shadow=views[0] if views else None)
                for extra in views[1:]:          # a weight mirrored in more than one operand buffer
                    ops.cast_bf16(p.data, extra)
                # the kernel writes through raw pointers: p._version is unchanged and the mirrors were refreshed in the same
                # pass, so the bf16 weight cache stays valid without a re-cast.
        return loss


def param_groups_like_reference(model, weight_decay=0.05):
    """fourm/utils/optim_factory.py:111-168: no weight decay for 1-D tensors, biases, and names the model lists."""
    skip = model.no_weight_decay() if hasattr(model, "no_weight_decay") else set()
    decay, no_decay = [], []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        if p.ndim == 1 or name.endswith(".bias") or name in skip or "norm." in name or ".norm" in name:
            no_decay.append(p)
        else:
            decay.append(p)
    return [dict(params=decay, weight_decay=weight_decay), dict(params=no_decay, weight_decay=0.0)]
