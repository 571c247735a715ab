"""GPU-side batch assembly of the reference's data path (utils/griddataset.py:88-174).

The reference builds every sample on the CPU inside DataLoader workers: ``pad_data`` (bilinear resize of all T*C planes to
``res`` + channel padding with 1.0), a random training window, the mask.  At B200 speeds (25 k field-steps/s) that path is
the bottleneck, so here the raw samples are copied once, as they lie in the HDF5 file, from pinned host memory and
everything else is ONE kernel (``dpot_assemble_batch``).  Reading HDF5 itself stays with h5py on the host (not in this
image); this module takes the raw arrays."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr


class BatchAssembler:
    """xx, yy, msk = assembler(raw, starts): raw = [B, H0, W0, T0, C0] float32 (pinned host tensor / numpy array: copied on
    a private stream, double-buffered, or a CUDA tensor), starts = per-sample first frame (None: random like
    utils/griddataset.py:153, ``np.random.randint(max(T0 - (t_in + t_ar) + 1, 1))``)."""

    def __init__(self, res: int, t_in: int, t_ar: int, n_channels: int, device="cuda", train: bool = True,
                 pred_channels: Optional[int] = None):
        self.res, self.t_in, self.t_ar, self.C = res, t_in, t_ar, n_channels
        self.dev = torch.device(device)
        self.train, self.pred_channels = train, pred_channels
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.stage = [None, None]
        self.turn = 0

    def _to_device(self, raw) -> torch.Tensor:
        if isinstance(raw, np.ndarray):
            raw = torch.from_numpy(np.ascontiguousarray(raw, dtype=np.float32))
        if raw.is_cuda:
            return raw.contiguous().float()
        i, self.turn = self.turn, self.turn ^ 1
        if self.stage[i] is None or self.stage[i].shape != raw.shape:
            self.stage[i] = torch.empty(raw.shape, device=self.dev, dtype=torch.float32)
        cur = torch.cuda.current_stream(self.dev)
        self.copy_stream.wait_stream(cur)                 # the kernel that last read this stage buffer has been queued
        with torch.cuda.stream(self.copy_stream):
            self.stage[i].copy_(raw, non_blocking=True)
        cur.wait_stream(self.copy_stream)
        return self.stage[i]

    def __call__(self, raw, starts=None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        with torch.cuda.device(self.dev):
            d = self._to_device(raw)
            if d.dim() == 4:                              # [B, H, W, T]: augment the channel axis (:146-147)
                d = d.unsqueeze(-1)
            B, H0, W0, T0, C0 = d.shape
            if starts is None:
                starts = np.random.randint(max(T0 - (self.t_in + self.t_ar) + 1, 1), size=B) if self.train else np.zeros(B)
            st = torch.as_tensor(np.asarray(starts, dtype=np.int32)).to(self.dev, non_blocking=True)
            xx = torch.empty((B, self.res, self.res, self.t_in, self.C), device=self.dev)
            yy = torch.empty((B, self.res, self.res, self.t_ar, self.C), device=self.dev)
            msk = torch.empty((B, self.res, self.res, 1, self.C), device=self.dev)
            check(_lib.load().dpot_assemble_batch(ptr(d), ptr(st), B, H0, W0, T0, C0, self.res, self.t_in, self.t_ar, self.C,
                                                  0 if self.train else 1, self.pred_channels or 0, ptr(xx), ptr(yy), ptr(msk),
                                                  torch.cuda.current_stream().cuda_stream), "dpot_assemble_batch")
        return xx, yy, msk
