"""The autoregressive training step of the reference (train_temporal.py:201-230 / train_temporal_parallel.py:205-246)
with every piece either side of ``model(xx)`` on kernels of libdpot_b200 and no host synchronisation:

    noise injection (:205)          -> NoiseInjectFn   (Philox normals, backward regenerates them)
    SimpleLpLoss (:207)             -> LpLossFn        (masked relative L2, forward + backward kernels)
    gradient exchange (DDP)         -> OverlappedGradArena (bucketed all-reduce overlapped with backward)
    clip_grad_norm_ + Adam (:228-9) -> one norm kernel + the clip coefficient applied inside the fused Adam kernel

The unchanged reference script keeps working on the drop-in ``DPOTNet`` / ``Adam``; this module is the loop a
maintainer switches to for the fused path."""
from __future__ import annotations

from typing import Optional

import torch
from torch.autograd import Function

from . import ops
from .utils.clip import clip_grad_norm_


class NoiseInjectFn(Function):
    """xx + scale * ||xx||_{(X,Y,T)} * randn_like(xx)  (train_temporal.py:205), differentiated through the norm like
    the reference's autograd does."""

    @staticmethod
    def forward(ctx, x, scale, seed, offset):
        x = x.contiguous()
        out, sumsq = ops.noise_inject(x, scale, seed, offset)
        ctx.save_for_backward(x, sumsq)
        ctx.args = (float(scale), int(seed), int(offset))
        return out

    @staticmethod
    def backward(ctx, dy):
        x, sumsq = ctx.saved_tensors
        scale, seed, offset = ctx.args
        return ops.noise_inject_bwd(x, dy, scale, seed, offset, sumsq), None, None, None


class LpLossFn(Function):
    """SimpleLpLoss(size_average=False)(x, y, mask=msk)  (utils/criterion.py:38-59) -> 0-dim device tensor."""

    @staticmethod
    def forward(ctx, x, y, mask):
        x, y = x.contiguous(), y.contiguous()
        mask = mask.contiguous() if mask is not None else None
        loss, coef = ops.lp_loss(x, y, mask)
        ctx.save_for_backward(x, y, mask, coef)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        x, y, mask, coef = ctx.saved_tensors
        return ops.lp_loss_bwd(x, y, mask, coef, g.reshape(1).contiguous().float()), None, None


def ar_train_step(model, optimizer, xx: torch.Tensor, yy: torch.Tensor, msk: Optional[torch.Tensor], *, T_bundle: int = 1,
                  noise_scale: float = 0.0, grad_clip: float = 10000.0, arena=None, seed: int = 0, step: int = 0,
                  scheduler=None) -> torch.Tensor:
    """One optimizer step of the loop of train_temporal.py:189-230.  xx[B,X,Y,T_in,C], yy[B,X,Y,T_ar,C],
    msk[B,X,Y,1,C] or None.  `arena`: an OverlappedGradArena / GradArena for data-parallel runs (its hooks fire during
    backward; finish()/allreduce() here).  Returns the summed step loss as a device tensor (no .item())."""
    loss = None
    n_ar = 0
    for t in range(0, yy.shape[-2], T_bundle):
        y = yy[..., t:t + T_bundle, :]
        if noise_scale > 0.0:
            xx = NoiseInjectFn.apply(xx, noise_scale, seed, step * 4096 + n_ar)
        im, _cls = model(xx)
        term = LpLossFn.apply(im, y, msk)
        loss = term if loss is None else loss + term
        n_ar += 1
        if t + T_bundle < yy.shape[-2]:      # the window shift (:219) feeds the NEXT forward only: not after the last one
            xx = torch.cat((xx[..., T_bundle:, :], im), dim=-2)
    optimizer.zero_grad(set_to_none=True)
    loss.backward()
    if arena is not None:
        arena.finish() if hasattr(arena, "finish") else arena.allreduce()
    clip_grad_norm_(model.parameters(), grad_clip, optimizer=optimizer)
    optimizer.step()
    if scheduler is not None:
        scheduler.step()
    return loss.detach()
