"""The evaluation loop of the reference (evaluate.py:182-222, train_temporal.py:258-283) on the library's rollout:

    for xx, yy, msk, _ in test_loader:   im, _ = model(xx); loss += myloss(im, y, mask=msk); pred = cat(pred, im); xx = cat(...)

becomes one RolloutEngine run per batch (ring window, no torch.cat, optionally one CUDA graph per trajectory length) plus
SimpleLpLoss kernels that accumulate on the device: nothing synchronises with the host until a loader is finished
(the reference calls loss.item() once per batch)."""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import torch

from . import ops
from .rollout import RolloutEngine


@torch.no_grad()
def evaluate_loaders(model, test_loaders: Sequence[Iterable], ntests: Sequence[int], T_bundle: int = 1,
                     use_graph: bool = False) -> Tuple[List[float], List[float]]:
    """Returns (test_l2_fulls, test_l2_steps) exactly as evaluate.py:213-217 defines them: per loader, the rel-L2 of the whole
    predicted trajectory / ntests and the mean over AR steps of the per-step rel-L2 / ntests."""
    model.eval()
    dev = next(model.parameters()).device
    engines = {}
    fulls, steps = [], []
    for lid, loader in enumerate(test_loaders):
        l2_step = torch.zeros(1, device=dev)
        l2_full = torch.zeros(1, device=dev)
        n_ar = 1
        for xx, yy, msk, _ in loader:
            xx = xx.to(dev, non_blocking=True).float().contiguous()
            yy = yy.to(dev, non_blocking=True).float().contiguous()
            msk = msk.to(dev, non_blocking=True).float().contiguous()
            n_ar = yy.shape[-2] // T_bundle
            key = (xx.shape[0], n_ar)
            eng = engines.get(key)
            if eng is None:
                eng = engines[key] = RolloutEngine(model, xx.shape[0], n_ar, device=dev, use_graph=use_graph, want_cls=True)
            pred = eng.run(xx)                                              # [B, X, Y, n_ar * T_bundle, C]
            for t in range(0, n_ar * T_bundle, T_bundle):                   # evaluate.py:201 loss += myloss(im, y, mask=msk)
                ops.lp_loss(pred[..., t:t + T_bundle, :].contiguous(), yy[..., t:t + T_bundle, :].contiguous(), msk,
                            loss=l2_step, accumulate=True)
            ops.lp_loss(pred, yy[..., :n_ar * T_bundle, :].contiguous(), msk, loss=l2_full, accumulate=True)   # :211
        fulls.append(float(l2_full) / ntests[lid])
        steps.append(float(l2_step) / ntests[lid] / n_ar)
    return fulls, steps
