"""clip_grad_norm_ for the fused optimizer (train_temporal.py:228: ``nn.utils.clip_grad_norm_(model.parameters(),
args.grad_clip)`` followed by ``optimizer.step()``).

``torch.nn.utils.clip_grad_norm_`` keeps working unchanged on our model (it is what the unchanged reference script
calls).  This variant is the fast path a maintainer opts into with one import change: the global norm is ONE
multi-tensor kernel (``dpot_grad_sqnorm``) and, when the fused ``Adam``/``AdamW`` of this package is passed as
``optimizer``, the scaling itself is deferred into the optimizer kernel -- gradients are read once per step and never
rewritten, and nothing synchronises with the host (the returned norm is a device tensor, like torch's)."""
from __future__ import annotations

from typing import Iterable, Optional

import torch

from .. import ops


@torch.no_grad()
def clip_grad_norm_(parameters: Iterable[torch.Tensor], max_norm: float, norm_type: float = 2.0, optimizer=None,
                    grad_scale: Optional[float] = None) -> torch.Tensor:
    """Returns the total gradient 2-norm (device tensor, fp32) BEFORE clipping, like torch.  With ``optimizer`` (a
    dpot_b200 Adam/AdamW) the clip coefficient is applied inside the next ``optimizer.step()``; without it the gradients
    are scaled in place here (torch semantics).  ``grad_scale``: the gradients will be multiplied by this factor in the
    step (1/world_size when the DDP average is folded into the optimizer); the norm is that of the scaled gradients."""
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    if float(norm_type) != 2.0:
        raise NotImplementedError("dpot_b200 clip_grad_norm_: only the 2-norm the reference uses is built")
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads:
        return torch.zeros((), device="cuda" if torch.cuda.is_available() else "cpu")
    gs = float(grad_scale if grad_scale is not None else (optimizer.grad_scale if optimizer is not None else 1.0))
    sq = ops.grad_sqnorm(grads)
    total = (sq.sqrt() * gs).to(torch.float32).reshape(())
    if optimizer is not None and hasattr(optimizer, "_pending_clip"):
        optimizer._pending_clip = (sq, float(max_norm))
    else:
        coef = torch.clamp(float(max_norm) / (total + 1e-6), max=1.0)
        torch._foreach_mul_(grads, coef)
    return total
