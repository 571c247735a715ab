"""Model-zoo configurations of the reference and seeded synthetic weights for benchmarks / smoke runs.

Sizes: configs/pretrain_tiny.yaml:62-85, pretrain_s.yaml:61-83, pretrain_medium.yaml:67-89,
pretrain_large.yaml:63-87, README.md:19-25 of the reference.  There is no network for checkpoints, so benchmarks run
random weights of the right architecture; `synthetic_weights_` draws them so that EVERY tensor matters to the output
(at the reference's default init the spectral branch is ~5e-5 of the signal and would hide a broken AFNO mixer).
"""
from __future__ import annotations

import math

import numpy as np
import torch

MODEL_ZOO = {
    "Ti": dict(embed_dim=512, depth=4, n_blocks=4, mlp_ratio=1, out_layer_dim=32),
    "S": dict(embed_dim=1024, depth=6, n_blocks=8, mlp_ratio=1, out_layer_dim=32),
    "M": dict(embed_dim=1024, depth=12, n_blocks=8, mlp_ratio=4, out_layer_dim=32),
    "L": dict(embed_dim=1536, depth=24, n_blocks=16, mlp_ratio=4, out_layer_dim=128),
    "H": dict(embed_dim=2048, depth=27, n_blocks=8, mlp_ratio=4, out_layer_dim=128),
}


def zoo_cfg(name: str, img_size: int = 128, patch_size: int = 8, **kw) -> dict:
    """Constructor keywords of DPOTNet (models/dpot.py:246-247) for one of the published sizes."""
    cfg = dict(img_size=img_size, patch_size=patch_size, mixing_type="afno", in_channels=4, out_channels=4,
               in_timesteps=10, out_timesteps=1, modes=32, n_cls=12, normalize=False, act="gelu", time_agg="exp_mlp")
    cfg.update(MODEL_ZOO[name])
    cfg.update(kw)
    return cfg


@torch.no_grad()
def synthetic_weights_(model: torch.nn.Module, seed: int = 0) -> torch.nn.Module:
    """Fill `model` (ours or the reference's DPOTNet: same state-dict schema) in place with seeded weights:
    conv / linear weights U(+-1/sqrt(fan_in)), spectral weights randn/sqrt(bs) with 0.1*randn biases, norm affines
    1 + 0.1*randn / 0.1*randn, pos_embed 0.02*randn, time-aggregation weights randn/(T*sqrt(E)), gamma as the reference
    initialises it.  One numpy Generator consumed in state-dict order, so the same seed gives the same weights for
    every implementation of the schema."""
    rng = np.random.default_rng(seed)
    sd = model.state_dict()
    E = sd["pos_embed"].shape[1]
    for name, t in sd.items():
        shp = tuple(t.shape)
        if name == "pos_embed":
            a = 0.02 * rng.standard_normal(shp)
        elif name == "time_agg_layer.gamma":
            a = 2.0 ** np.linspace(-10, 10, E).reshape(1, E)
        elif name == "time_agg_layer.w":
            a = rng.standard_normal(shp) / (shp[0] * math.sqrt(E))
        elif ".filter.w" in name:
            a = rng.standard_normal(shp) / math.sqrt(shp[2])
        elif ".filter.b" in name:
            a = 0.1 * rng.standard_normal(shp)
        elif ".norm" in name and name.endswith("weight"):
            a = 1.0 + 0.1 * rng.standard_normal(shp)
        elif ".norm" in name:
            a = 0.1 * rng.standard_normal(shp)
        elif name.endswith("bias"):
            a = 0.1 * rng.uniform(-1, 1, shp)
        else:
            fan_in = shp[0] if name == "out_layer.0.weight" else int(np.prod(shp[1:]))   # ConvTranspose2d: (in, out, kh, kw)
            a = rng.uniform(-1, 1, shp) / math.sqrt(fan_in)
        t.copy_(torch.from_numpy(np.ascontiguousarray(a.astype(np.float32))).reshape(shp))
    return model


def forward_flops(cfg: dict) -> dict:
    """FLOPs per field-step of one forward (2 per MAC).  'algorithmic' follows the reference's formulation
    (SURVEY.md 8d); 'executed' is what this library runs after folding PatchEmbed-1x1 + pos_embed + TimeAggregator into
    one K = T*(C_out*P+3) contraction."""
    E, P, C, Co = cfg["embed_dim"], cfg["patch_size"], cfg["in_channels"], cfg["out_channels"]
    T, To, nb, D = cfg["in_timesteps"], cfg["out_timesteps"], cfg["n_blocks"], cfg["depth"]
    h = cfg["img_size"] // P
    n, bs, hid, old, mid = h * h, E // nb, int(E * cfg["mlp_ratio"]), cfg["out_layer_dim"], Co * P + 3
    km1, km2 = min(cfg["modes"], h), min(cfg["modes"], h // 2 + 1)
    R = cfg["img_size"]
    conv0 = 2 * T * n * (C + 3) * P * P * mid
    conv1 = 2 * T * n * mid * E
    tagg = 2 * n * T * E * E
    spectral = D * (2 * 2 * km1 * km2 * (2 * bs) * (2 * bs) * nb)   # two layers, real form [2bs x 2bs] per mode and block
    mlp = D * 2 * 2 * n * E * hid
    out = 2 * n * E * old * P * P + 2 * R * R * (old * old + old * Co * To)
    fft = D * 2 * E * 2.5 * n * math.log2(n)                         # real transforms: half of 5 N log2 N
    algorithmic = conv0 + conv1 + tagg + spectral + mlp + out + fft
    executed = 2 * T * n * C * P * P * mid + 2 * n * (T * mid) * E + spectral + mlp + out + fft
    return dict(algorithmic=float(algorithmic), executed=float(executed), spectral=float(spectral), mlp=float(mlp))
