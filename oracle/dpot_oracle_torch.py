"""CPU timing port of the DPOT forward on PyTorch's CPU operators.

TEST / BENCH INFRASTRUCTURE ONLY (same rules as dpot_oracle.py: imported by ``tests/`` and the CPU legs of
``bench.py``, never by the product package).  The reference is a PyTorch program whose CPU path runs on ATen's
oneDNN / MKL / pocketfft kernels; this file restates DPOTNet.forward (models/dpot.py:364-403) with the SAME operator
classes (conv2d, group_norm, rfft2 / irfft2, einsum, conv_transpose2d), so that the ``cpu_baseline`` of bench.py is
timed at the speed the reference itself would reach on the host cores instead of at numpy speed.  Parameters are the
reference's state-dict tensors.  Checked against the numpy oracle in
tests/test_oracle_golden.py::test_torch_port_matches_numpy_oracle.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn.functional as F

_ACT = {"gelu": F.gelu, "tanh": torch.tanh, "sigmoid": torch.sigmoid, "relu": F.relu,
        "leaky_relu": lambda x: F.leaky_relu(x, 0.1), "softplus": F.softplus, "ELU": F.elu, "silu": F.silu}


def to_torch(params: Dict[str, np.ndarray]) -> Dict[str, torch.Tensor]:
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in params.items()}


def _afno2d(x: torch.Tensor, p, pre: str, nb: int, modes: int, act) -> torch.Tensor:
    """AFNO2D.forward, models/dpot.py:51-110, x[B,E,H,W] (channel_first)."""
    B, E, H, W = x.shape
    bs = E // nb
    xl = x.permute(0, 2, 3, 1)
    f = torch.fft.rfft2(xl, dim=(1, 2), norm="ortho").reshape(B, H, W // 2 + 1, nb, bs)      # :59-62
    w1, b1, w2, b2 = p[pre + "w1"], p[pre + "b1"], p[pre + "w2"], p[pre + "b2"]
    km = modes
    fr, fi = f.real[:, :km, :km], f.imag[:, :km, :km]
    mm = lambda a, w: torch.einsum("...bi,bio->...bo", a, w)
    o1r = act(mm(fr, w1[0]) - mm(fi, w1[1]) + b1[0])                                         # :72-82
    o1i = act(mm(fi, w1[0]) + mm(fr, w1[1]) + b1[1])
    o2r = mm(o1r, w2[0]) - mm(o1i, w2[1]) + b2[0]                                            # :84-94
    o2i = mm(o1i, w2[0]) + mm(o1r, w2[1]) + b2[1]
    o = torch.zeros(f.shape, dtype=torch.complex64)
    o[:, :km, :km] = torch.complex(o2r, o2i)
    y = torch.fft.irfft2(o.reshape(B, H, W // 2 + 1, E), s=(H, W), dim=(1, 2), norm="ortho")  # :101-102
    return (y + xl).permute(0, 3, 1, 2)                                                       # :106-108


@torch.no_grad()
def dpot_forward(x: torch.Tensor, p: Dict[str, torch.Tensor], cfg: dict) -> Tuple[torch.Tensor, torch.Tensor]:
    """x[B,X,Y,T,C] fp32 -> (y[B,X,Y,To,Co], cls[B,n_cls]); normalize=False, time_agg in {'exp_mlp','mlp'}."""
    assert not cfg["normalize"], "the timing port covers the BASELINE configs (normalize=False)"
    act = _ACT[cfg["act"]]
    B, X, Y, T, C = x.shape
    P = cfg["patch_size"]
    lin = lambda n: torch.tensor(np.linspace(0, 1, n), dtype=torch.float)                     # :350-357
    g = torch.cat([lin(X).reshape(1, X, 1, 1, 1).expand(B, X, Y, T, 1), lin(Y).reshape(1, 1, Y, 1, 1).expand(B, X, Y, T, 1),
                   lin(T).reshape(1, 1, 1, T, 1).expand(B, X, Y, T, 1)], dim=-1)
    x7 = torch.cat([x, g], dim=-1).permute(0, 3, 4, 1, 2).reshape(B * T, C + 3, X, Y)          # :373-375
    z = F.conv2d(x7, p["patch_embed.proj.0.weight"], p["patch_embed.proj.0.bias"], stride=P)  # :199
    z = F.conv2d(act(z), p["patch_embed.proj.2.weight"], p["patch_embed.proj.2.bias"])        # :200-201
    z = z + p["pos_embed"]                                                                    # :378
    E, h, w = z.shape[1:]
    z = z.reshape(B, T, E, h, w).permute(0, 3, 4, 1, 2)                                        # :380
    if cfg["time_agg"] == "exp_mlp":                                                          # :229-232
        t = torch.linspace(0, 1, T).unsqueeze(-1)
        z = z * torch.cos(t @ p["time_agg_layer.gamma"])
    a = torch.einsum("tij,...ti->...j", p["time_agg_layer.w"], z).permute(0, 3, 1, 2)          # :228/232, :384
    for i in range(cfg["depth"]):                                                             # Block.forward :165-180
        pre = f"blocks.{i}."
        r = a
        a = F.group_norm(a, 8, p[pre + "norm1.weight"], p[pre + "norm1.bias"], 1e-5)
        a = _afno2d(a, p, pre + "filter.", cfg["n_blocks"], cfg["modes"], act)
        a = F.group_norm(a, 8, p[pre + "norm2.weight"], p[pre + "norm2.bias"], 1e-5)
        a = F.conv2d(act(F.conv2d(a, p[pre + "mlp.0.weight"], p[pre + "mlp.0.bias"])), p[pre + "mlp.2.weight"], p[pre + "mlp.2.bias"])
        a = a + r
    tok = a.mean(dim=(2, 3))                                                                  # :394-395
    cls = F.linear(act(F.linear(act(F.linear(tok, p["cls_head.0.weight"], p["cls_head.0.bias"])), p["cls_head.2.weight"],
                                p["cls_head.2.bias"])), p["cls_head.4.weight"], p["cls_head.4.bias"])
    y = act(F.conv_transpose2d(a, p["out_layer.0.weight"], p["out_layer.0.bias"], stride=P))  # :316-321
    y = act(F.conv2d(y, p["out_layer.2.weight"], p["out_layer.2.bias"]))
    y = F.conv2d(y, p["out_layer.4.weight"], p["out_layer.4.bias"]).permute(0, 2, 3, 1)
    return y.reshape(*y.shape[:3], cfg["out_timesteps"], cfg["out_channels"]).contiguous(), cls  # :398
