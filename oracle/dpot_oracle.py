"""CPU oracle for the DPOT autoregressive Fourier-operator hot path.

TEST INFRASTRUCTURE ONLY.  This file is a plain-numpy restatement of the
reference algorithm (HaoZhongkai/DPOT @ dcd2f9a).  It may be imported only by
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` -- never by the product package
``dpot_b200``.  Nothing here touches CUDA.

Parity pin: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
itself, generated in the build container by ``tests/golden/make_golden.py``
(which imports ``/root/reference`` read-only) and committed as
``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks every fixture.

Every function cites the reference file:line it restates.  All arithmetic is
done in the dtype of the inputs (float32 to mimic the reference, float64 for
an error-budget ground truth).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np

try:  # exact erf; scipy is in the image, math.erf is the fallback
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf, otypes=[np.float64])

Params = Dict[str, np.ndarray]


# --------------------------------------------------------------------------- #
# configuration helpers
# --------------------------------------------------------------------------- #
DEFAULT_CFG = dict(
    img_size=224, patch_size=16, mixing_type="afno", in_channels=1, out_channels=4,
    in_timesteps=1, out_timesteps=1, n_blocks=4, embed_dim=768, out_layer_dim=32,
    depth=12, modes=32, mlp_ratio=1.0, n_cls=12, normalize=False, act="gelu",
    time_agg="exp_mlp",
)  # defaults of DPOTNet.__init__, models/dpot.py:246-247

# model sizes: configs/pretrain_tiny.yaml:62-85, pretrain_s.yaml:61-83,
# pretrain_medium.yaml:67-89, pretrain_large.yaml:63-87, README.md:19-25
MODEL_ZOO = {
    "Ti": dict(embed_dim=512, depth=4, n_blocks=4, mlp_ratio=1, out_layer_dim=32),
    "S": dict(embed_dim=1024, depth=6, n_blocks=8, mlp_ratio=1, out_layer_dim=32),
    "M": dict(embed_dim=1024, depth=12, n_blocks=8, mlp_ratio=4, out_layer_dim=32),
    "L": dict(embed_dim=1536, depth=24, n_blocks=16, mlp_ratio=4, out_layer_dim=128),
    "H": dict(embed_dim=2048, depth=27, n_blocks=8, mlp_ratio=4, out_layer_dim=128),
}


def make_cfg(**kw) -> dict:
    cfg = dict(DEFAULT_CFG)
    cfg.update(kw)
    return cfg


def zoo_cfg(name: str, img_size=128, patch_size=8, **kw) -> dict:
    base = dict(img_size=img_size, patch_size=patch_size, in_channels=4, out_channels=4,
                in_timesteps=10, out_timesteps=1, modes=32, n_cls=12)
    base.update(MODEL_ZOO[name])
    base.update(kw)
    return make_cfg(**base)


# --------------------------------------------------------------------------- #
# elementwise
# --------------------------------------------------------------------------- #
def activation(x: np.ndarray, act: str = "gelu") -> np.ndarray:
    """ACTIVATION table, models/dpot.py:19 (torch module defaults)."""
    dt = x.dtype
    if act == "gelu":  # nn.GELU() exact erf form
        return (0.5 * x * (1.0 + _erf(x / math.sqrt(2.0)))).astype(dt)
    if act == "tanh":
        return np.tanh(x)
    if act == "sigmoid":
        return (1.0 / (1.0 + np.exp(-x))).astype(dt)
    if act == "relu":
        return np.maximum(x, 0).astype(dt)
    if act == "leaky_relu":  # nn.LeakyReLU(0.1)
        return np.where(x > 0, x, 0.1 * x).astype(dt)
    if act == "softplus":  # beta=1, threshold=20
        return np.where(x > 20, x, np.log1p(np.exp(np.minimum(x, 20)))).astype(dt)
    if act == "ELU":
        return np.where(x > 0, x, np.expm1(np.minimum(x, 0))).astype(dt)
    if act == "silu":
        return (x / (1.0 + np.exp(-x))).astype(dt)
    raise KeyError(act)


def group_norm(x: np.ndarray, gamma: np.ndarray, beta: np.ndarray, groups: int = 8,
               eps: float = 1e-5) -> np.ndarray:
    """torch.nn.GroupNorm(8, E) on x[B,E,h,w]; models/dpot.py:142,152,167,175.
    Biased variance over (E/groups * h * w) values per (sample, group)."""
    B, E = x.shape[:2]
    xg = x.reshape(B, groups, -1)
    mu = xg.mean(axis=2, keepdims=True)
    var = ((xg - mu) ** 2).mean(axis=2, keepdims=True)
    xn = ((xg - mu) / np.sqrt(var + x.dtype.type(eps))).reshape(x.shape)
    return xn * gamma.reshape(1, E, 1, 1) + beta.reshape(1, E, 1, 1)


# --------------------------------------------------------------------------- #
# model stages
# --------------------------------------------------------------------------- #
def grid_3d(B: int, X: int, Y: int, T: int, dtype) -> np.ndarray:
    """DPOTNet.get_grid_3d, models/dpot.py:350-360: linspace(0,1,n) along x, y and t
    (np float64 -> float32 in the reference)."""
    gx = np.linspace(0, 1, X).astype(np.float32).astype(dtype).reshape(1, X, 1, 1, 1)
    gy = np.linspace(0, 1, Y).astype(np.float32).astype(dtype).reshape(1, 1, Y, 1, 1)
    gt = np.linspace(0, 1, T).astype(np.float32).astype(dtype).reshape(1, 1, 1, T, 1)
    shape = (B, X, Y, T, 1)
    return np.concatenate([np.broadcast_to(gx, shape), np.broadcast_to(gy, shape),
                           np.broadcast_to(gt, shape)], axis=-1)


def patch_embed(x7: np.ndarray, p: Params, P: int, act: str) -> np.ndarray:
    """PatchEmbed.forward, models/dpot.py:198-209, on frames x7[(b t), C+3, X, Y]:
    Conv2d(C+3 -> Co*P+3, k=s=P) -> act -> Conv2d(-> E, 1x1).  Returns [(b t), E, h, w]."""
    W0, b0 = p["patch_embed.proj.0.weight"], p["patch_embed.proj.0.bias"]
    W2, b2 = p["patch_embed.proj.2.weight"], p["patch_embed.proj.2.bias"]
    N, C7, X, Y = x7.shape
    h, w = X // P, Y // P
    xp = x7.reshape(N, C7, h, P, w, P)
    z0 = np.einsum("nkpuqv,mkuv->nmpq", xp, W0, optimize=True) + b0.reshape(1, -1, 1, 1)
    z1 = activation(z0.astype(x7.dtype), act)
    z2 = np.einsum("nmpq,em->nepq", z1, W2[:, :, 0, 0], optimize=True) + b2.reshape(1, -1, 1, 1)
    return z2.astype(x7.dtype)


def time_aggregate(x: np.ndarray, p: Params, kind: str) -> np.ndarray:
    """TimeAggregator.forward, models/dpot.py:226-234 on x[B,h,w,T,E] -> [B,h,w,E].
    exp_mlp: t_embed = cos(linspace(0,1,T)[:,None] @ gamma[1,E]) (fp32 linspace)."""
    w = p["time_agg_layer.w"]
    if kind == "mlp":
        return np.einsum("tij,bxyti->bxyj", w, x, optimize=True).astype(x.dtype)
    T = x.shape[-2]
    t = np.linspace(0, 1, T).astype(np.float32).astype(x.dtype).reshape(T, 1)
    t_embed = np.cos(t @ p["time_agg_layer.gamma"].astype(x.dtype))  # [T,E]
    return np.einsum("tij,bxyti->bxyj", w, x * t_embed, optimize=True).astype(x.dtype)


def afno2d(x: np.ndarray, p: Params, prefix: str, n_blocks: int, modes: int, act: str) -> np.ndarray:
    """AFNO2D.forward with channel_first=True, models/dpot.py:51-110, x[B,E,H,W].
    rfft2(ortho) -> block-diagonal complex 2-layer MLP on the kept modes
    [:modes, :modes] (python slicing clamps) -> irfft2(s=(H,W), ortho) -> + x.
    Softshrink is commented out in the reference (:97-98) and therefore absent."""
    B, E, H, W = x.shape
    bs = E // n_blocks
    xl = np.transpose(x, (0, 2, 3, 1))  # :54  -> B,H,W,E
    F = np.fft.rfft2(xl, axes=(1, 2), norm="ortho")  # :59
    cdt = np.complex64 if x.dtype == np.float32 else np.complex128
    F = F.astype(cdt).reshape(B, H, W // 2 + 1, n_blocks, bs)  # :62
    w1, b1 = p[prefix + "w1"], p[prefix + "b1"]
    w2, b2 = p[prefix + "w2"], p[prefix + "b2"]
    km = modes
    Fk = F[:, :km, :km]
    fr, fi = Fk.real.astype(x.dtype), Fk.imag.astype(x.dtype)

    def mm(a, wgt):  # einsum('...bi,bio->...bo') as a BLAS batched matmul over the nb blocks
        lead = a.shape[:-2]
        r = np.matmul(np.ascontiguousarray(a.reshape(-1, n_blocks, bs).transpose(1, 0, 2)), wgt)
        return np.ascontiguousarray(r.transpose(1, 0, 2)).reshape(*lead, n_blocks, bs).astype(x.dtype)

    o1r = activation(mm(fr, w1[0]) - mm(fi, w1[1]) + b1[0], act)  # :72-76
    o1i = activation(mm(fi, w1[0]) + mm(fr, w1[1]) + b1[1], act)  # :78-82
    o2r = mm(o1r, w2[0]) - mm(o1i, w2[1]) + b2[0]  # :84-88
    o2i = mm(o1i, w2[0]) + mm(o1r, w2[1]) + b2[1]  # :90-94
    O = np.zeros(F.shape, dtype=cdt)  # :66-67 zeros outside the kept region
    O[:, :km, :km] = (o2r + 1j * o2i).astype(cdt)
    O = O.reshape(B, H, W // 2 + 1, E)  # :101
    y = irfft2_ortho(O, H, W).astype(x.dtype)  # :102
    y = y + xl  # :106
    return np.transpose(y, (0, 3, 1, 2))  # :108


def irfft2_ortho(O: np.ndarray, H: int, W: int) -> np.ndarray:
    """torch.fft.irfft2(O, s=(H,W), dim=(1,2), norm='ortho') for NON-Hermitian input
    (models/dpot.py:102): complex ifft over axis 1, then c2r over axis 2 which
    ignores Im of the k2=0 and k2=W/2 columns (SURVEY.md section 7 hard part 3)."""
    Z = np.fft.ifft(O, axis=1, norm="ortho")
    return np.fft.irfft(Z, n=W, axis=2, norm="ortho")


def block(x: np.ndarray, p: Params, i: int, cfg: dict) -> np.ndarray:
    """Block.forward with double_skip=False, models/dpot.py:165-180."""
    pre = f"blocks.{i}."
    r = x
    x = group_norm(x, p[pre + "norm1.weight"], p[pre + "norm1.bias"])  # :167
    x = afno2d(x, p, pre + "filter.", cfg["n_blocks"], cfg["modes"], cfg["act"])  # :168
    x = group_norm(x, p[pre + "norm2.weight"], p[pre + "norm2.bias"])  # :175
    W1, c1 = p[pre + "mlp.0.weight"][:, :, 0, 0], p[pre + "mlp.0.bias"]
    W2, c2 = p[pre + "mlp.2.weight"][:, :, 0, 0], p[pre + "mlp.2.bias"]
    hdn = np.einsum("behw,oe->bohw", x, W1, optimize=True).astype(x.dtype) + c1.reshape(1, -1, 1, 1)
    hdn = activation(hdn, cfg["act"])
    x = np.einsum("bohw,eo->behw", hdn, W2, optimize=True).astype(x.dtype) + c2.reshape(1, -1, 1, 1)  # :176
    return x + r  # :178


def cls_head(x: np.ndarray, p: Params, act: str) -> np.ndarray:
    """models/dpot.py:303-309, 394-395: spatial mean -> Linear-act-Linear-act-Linear."""
    tok = x.mean(axis=(2, 3))
    y = activation(tok @ p["cls_head.0.weight"].T + p["cls_head.0.bias"], act)
    y = activation(y @ p["cls_head.2.weight"].T + p["cls_head.2.bias"], act)
    return (y @ p["cls_head.4.weight"].T + p["cls_head.4.bias"]).astype(x.dtype)


def out_layer(x: np.ndarray, p: Params, P: int, act: str) -> np.ndarray:
    """models/dpot.py:315-321: ConvTranspose2d(E->old,k=s=P) -> act -> 1x1 -> act -> 1x1.
    x[B,E,h,w] -> [B, To*Co, h*P, w*P]."""
    Wt, bt = p["out_layer.0.weight"], p["out_layer.0.bias"]  # (E, old, P, P)
    B, E, h, w = x.shape
    old = Wt.shape[1]
    y0 = np.einsum("bepq,eouv->bopuqv", x, Wt, optimize=True).astype(x.dtype)
    y0 = y0.reshape(B, old, h * P, w * P) + bt.reshape(1, -1, 1, 1)
    y1 = activation(y0, act)
    y2 = np.einsum("boxy,no->bnxy", y1, p["out_layer.2.weight"][:, :, 0, 0], optimize=True).astype(x.dtype)
    y2 = activation(y2 + p["out_layer.2.bias"].reshape(1, -1, 1, 1), act)
    y3 = np.einsum("boxy,no->bnxy", y2, p["out_layer.4.weight"][:, :, 0, 0], optimize=True).astype(x.dtype)
    return y3 + p["out_layer.4.bias"].reshape(1, -1, 1, 1)


def latent_forward(x: np.ndarray, p: Params, cfg: dict):
    """First half of DPOTNet.forward, models/dpot.py:364-387: returns the latent
    a[B,E,h,w] after the temporal aggregation (and AdaIN when normalize), plus (mu, sigma)."""
    B, X, Y, T, C = x.shape
    dt = x.dtype
    mu = sigma = None
    if cfg["normalize"]:  # :366-370 (torch.std is unbiased)
        mu = x.mean(axis=(1, 2, 3), keepdims=True)
        sigma = x.std(axis=(1, 2, 3), keepdims=True, ddof=1) + dt.type(1e-6)
        x = (x - mu) / sigma
        ms = np.concatenate([mu, sigma], axis=-1)  # B,1,1,1,2C
        s_mu = ms @ p["scale_feats_mu.weight"].T + p["scale_feats_mu.bias"]
        s_sg = ms @ p["scale_feats_sigma.weight"].T + p["scale_feats_sigma.bias"]
        s_mu = np.transpose(s_mu[:, :, :, 0, :], (0, 3, 1, 2))  # B,E,1,1
        s_sg = np.transpose(s_sg[:, :, :, 0, :], (0, 3, 1, 2))
    g = grid_3d(B, X, Y, T, dt)  # :373
    x7 = np.concatenate([x, g], axis=-1)  # :374
    x7 = np.transpose(x7, (0, 3, 4, 1, 2)).reshape(B * T, C + 3, X, Y)  # :375
    z = patch_embed(x7, p, cfg["patch_size"], cfg["act"])  # :376
    z = z + p["pos_embed"]  # :378
    E, h, w = z.shape[1:]
    z = np.transpose(z.reshape(B, T, E, h, w), (0, 3, 4, 1, 2))  # :380 -> b x y t c
    a = time_aggregate(z, p, cfg["time_agg"])  # :382
    a = np.transpose(a, (0, 3, 1, 2))  # :384
    if cfg["normalize"]:
        a = s_sg * a + s_mu  # :387
    return a.astype(dt), mu, sigma


def dpot_forward(x: np.ndarray, p: Params, cfg: dict) -> Tuple[np.ndarray, np.ndarray]:
    """DPOTNet.forward, models/dpot.py:364-403.  x[B,X,Y,T,C] -> (y[B,X,Y,To,Co], cls[B,n_cls])."""
    a, mu, sigma = latent_forward(x, p, cfg)
    for i in range(cfg["depth"]):  # :389
        a = block(a, p, i, cfg)
    cls = cls_head(a, p, cfg["act"])  # :394-395
    y = out_layer(a, p, cfg["patch_size"], cfg["act"])  # :397
    y = np.transpose(y, (0, 2, 3, 1))
    y = y.reshape(*y.shape[:3], cfg["out_timesteps"], cfg["out_channels"])  # :398
    if cfg["normalize"]:
        y = y * sigma + mu  # :401
    return np.ascontiguousarray(y).astype(x.dtype), cls


def rollout(xx: np.ndarray, p: Params, cfg: dict, n_steps: int) -> np.ndarray:
    """Autoregressive rollout without noise, evaluate.py:192-208 / train_temporal.py:262-272:
    im = model(xx); pred = cat(pred, im); xx = cat(xx[..., T_bundle:, :], im).  Returns
    pred[B,X,Y,n_steps*T_bundle,C]."""
    Tb = cfg["out_timesteps"]
    preds = []
    for _ in range(n_steps):
        im, _ = dpot_forward(xx, p, cfg)
        preds.append(im)
        xx = np.concatenate([xx[..., Tb:, :], im], axis=-2)
    return np.concatenate(preds, axis=-2)


# --------------------------------------------------------------------------- #
# loss and optimiser
# --------------------------------------------------------------------------- #
def simple_lp_loss(x: np.ndarray, y: np.ndarray, mask: Optional[np.ndarray] = None) -> np.ndarray:
    """SimpleLpLoss(size_average=False)(x, y, mask), utils/criterion.py:38-59 (p=2):
    sum over batch of (sum_c ||x-y||_2 / (||y||_2 + 1e-8)) / #active channels."""
    n = x.shape[0]
    if mask is not None:
        x = x * mask
        y = y * mask
        msk_ch = np.count_nonzero(mask.sum(axis=tuple(range(1, mask.ndim - 1))), axis=-1)
    else:
        msk_ch = x.shape[-1]
    C = x.shape[-1]
    d = np.sqrt(((x.reshape(n, -1, C) - y.reshape(n, -1, C)) ** 2).sum(axis=1))
    yn = np.sqrt((y.reshape(n, -1, C) ** 2).sum(axis=1)) + x.dtype.type(1e-8)
    return np.sum(np.sum(d / yn, axis=-1) / msk_ch)


def adam_step(p: np.ndarray, g: np.ndarray, m: np.ndarray, v: np.ndarray, step: int, *, lr: float,
              beta1: float, beta2: float, eps: float, weight_decay: float, decoupled: bool = False,
              vmax: Optional[np.ndarray] = None):
    """adam(), utils/optimizer.py:9-52 (decoupled=False, L2-coupled decay) and
    adamw(), utils/optimizer.py:170-212 (decoupled=True).  `step` is the 1-based
    step count AFTER the increment (:148-150).  Updates p, m, v (and vmax when
    amsgrad) in place, same operation order as the reference."""
    dt = p.dtype.type
    if decoupled:
        p *= dt(1 - lr * weight_decay)  # :193
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    if weight_decay != 0 and not decoupled:
        g = g + dt(weight_decay) * p  # :36-37
    m *= dt(beta1)
    m += dt(1 - beta1) * g  # :40
    v *= dt(beta2)
    v += dt(1 - beta2) * (g * g)  # :41 (g*conj(g) for real g)
    if vmax is not None:
        np.maximum(vmax, v, out=vmax)  # :44
        denom = np.sqrt(vmax) / dt(math.sqrt(bc2)) + dt(eps)
    else:
        denom = np.sqrt(v) / dt(math.sqrt(bc2)) + dt(eps)  # :48
    p -= dt(lr / bc1) * (m / denom)  # :50-52
    return p, m, v


def lamb_step(p: np.ndarray, g: np.ndarray, m: np.ndarray, v: np.ndarray, step: int, *, lr: float, beta1: float,
              beta2: float, eps: float, weight_decay: float, clamp_value: float = 10.0, adam: bool = False,
              debias: bool = False):
    """Lamb.step for one parameter, utils/optimizer.py:459-491.  `step` is the 1-based count after the increment (:460).
    Updates p, m, v in place; returns (weight_norm, adam_norm, trust_ratio) as the reference stores them (:485-487)."""
    dt = p.dtype.type
    m *= dt(beta1)
    m += dt(1 - beta1) * g  # :463
    v *= dt(beta2)
    v += dt(1 - beta2) * (g * g)  # :465
    corr = math.sqrt(1 - beta2 ** step) / (1 - beta1 ** step) if debias else 1  # :468-472
    step_size = lr * corr  # :475
    w_norm = dt(min(max(float(np.sqrt(np.sum(p.astype(np.float64) ** 2))), 0.0), clamp_value))  # :474
    upd = m / (np.sqrt(v) + dt(eps))  # :476
    if weight_decay != 0:
        upd = upd + dt(weight_decay) * p  # :477-478
    u_norm = dt(np.sqrt(np.sum(upd.astype(np.float64) ** 2)))  # :479
    trust = dt(1) if (w_norm == 0 or u_norm == 0) else dt(w_norm / u_norm)  # :480-483
    p += dt(-step_size * float(dt(1) if adam else trust)) * upd  # :488-491
    return w_norm, u_norm, trust


# --------------------------------------------------------------------------- #
# synthetic parameters / inputs (platform-independent: numpy Generator streams)
# --------------------------------------------------------------------------- #
def param_shapes(cfg: dict) -> List[Tuple[str, Tuple[int, ...]]]:
    """state_dict schema in registration order, SURVEY.md appendix A / models/dpot.py:278-321."""
    E, P, C, Co = cfg["embed_dim"], cfg["patch_size"], cfg["in_channels"], cfg["out_channels"]
    T, To, nb = cfg["in_timesteps"], cfg["out_timesteps"], cfg["n_blocks"]
    h = cfg["img_size"] // P
    bs = E // nb
    hid = int(E * cfg["mlp_ratio"])
    old = cfg["out_layer_dim"]
    mid = Co * P + 3
    s: List[Tuple[str, Tuple[int, ...]]] = [
        ("pos_embed", (1, E, h, h)),
        ("patch_embed.proj.0.weight", (mid, C + 3, P, P)), ("patch_embed.proj.0.bias", (mid,)),
        ("patch_embed.proj.2.weight", (E, mid, 1, 1)), ("patch_embed.proj.2.bias", (E,)),
    ]
    for i in range(cfg["depth"]):
        b = f"blocks.{i}."
        s += [(b + "norm1.weight", (E,)), (b + "norm1.bias", (E,)),
              (b + "filter.w1", (2, nb, bs, bs)), (b + "filter.b1", (2, nb, bs)),
              (b + "filter.w2", (2, nb, bs, bs)), (b + "filter.b2", (2, nb, bs)),
              (b + "norm2.weight", (E,)), (b + "norm2.bias", (E,)),
              (b + "mlp.0.weight", (hid, E, 1, 1)), (b + "mlp.0.bias", (hid,)),
              (b + "mlp.2.weight", (E, hid, 1, 1)), (b + "mlp.2.bias", (E,))]
    if cfg["normalize"]:
        s += [("scale_feats_mu.weight", (E, 2 * C)), ("scale_feats_mu.bias", (E,)),
              ("scale_feats_sigma.weight", (E, 2 * C)), ("scale_feats_sigma.bias", (E,))]
    s += [("cls_head.0.weight", (E, E)), ("cls_head.0.bias", (E,)),
          ("cls_head.2.weight", (E, E)), ("cls_head.2.bias", (E,)),
          ("cls_head.4.weight", (cfg["n_cls"], E)), ("cls_head.4.bias", (cfg["n_cls"],)),
          ("time_agg_layer.w", (T, E, E))]
    if cfg["time_agg"] == "exp_mlp":
        s += [("time_agg_layer.gamma", (1, E))]
    s += [("out_layer.0.weight", (E, old, P, P)), ("out_layer.0.bias", (old,)),
          ("out_layer.2.weight", (old, old, 1, 1)), ("out_layer.2.bias", (old,)),
          ("out_layer.4.weight", (Co * To, old, 1, 1)), ("out_layer.4.bias", (Co * To,))]
    return s


def make_params(cfg: dict, seed: int = 0, dtype=np.float32) -> Params:
    """Seeded synthetic weights in which EVERY tensor matters to the output: conv/linear
    weights ~ U(+-1/sqrt(fan_in)) like torch defaults, spectral weights re-scaled to
    randn/sqrt(bs) and 0.1*randn biases (SURVEY.md finding 3: at the reference's default
    init the spectral branch is 5e-5 of the signal and would hide errors), norm affine
    1+0.1*randn / 0.1*randn, gamma as the reference (2**linspace(-10,10,E))."""
    rng = np.random.default_rng(seed)
    E, nb = cfg["embed_dim"], cfg["n_blocks"]
    bs = E // nb
    out: Params = {}
    for name, shp in param_shapes(cfg):
        if name == "pos_embed":
            a = 0.02 * rng.standard_normal(shp)
        elif name == "time_agg_layer.gamma":
            a = 2.0 ** np.linspace(-10, 10, E).reshape(1, E)
        elif name == "time_agg_layer.w":
            a = rng.standard_normal(shp) / (shp[0] * math.sqrt(E))
        elif ".filter.w" in name:
            a = rng.standard_normal(shp) / math.sqrt(bs)
        elif ".filter.b" in name:
            a = 0.1 * rng.standard_normal(shp)
        elif ".norm" in name and name.endswith("weight"):
            a = 1.0 + 0.1 * rng.standard_normal(shp)
        elif ".norm" in name:
            a = 0.1 * rng.standard_normal(shp)
        elif name.endswith("bias"):
            a = 0.1 * rng.uniform(-1, 1, shp)
        else:
            if name == "out_layer.0.weight":  # ConvTranspose2d (in,out,kh,kw)
                fan_in = shp[0]
            else:
                fan_in = int(np.prod(shp[1:]))
            a = rng.uniform(-1, 1, shp) / math.sqrt(fan_in)
        out[name] = np.ascontiguousarray(a.astype(dtype))
    return out


def make_input(cfg: dict, B: int, seed: int = 0, kind: str = "randn", dtype=np.float32) -> np.ndarray:
    """Synthetic xx[B,R,R,T,C].  'randn' = iid normal; 'ns2d' = channel 0 a
    low-pass (|k|<=12) random field, remaining channels == 1.0 (the loader pads missing
    channels with ones, utils/griddataset.py:98-99)."""
    rng = np.random.default_rng(1000 + seed)
    R, T, C = cfg["img_size"], cfg["in_timesteps"], cfg["in_channels"]
    if kind == "randn":
        return rng.standard_normal((B, R, R, T, C)).astype(dtype)
    x = np.ones((B, R, R, T, C), dtype=np.float64)
    f = rng.standard_normal((B, R, R, T))
    Fk = np.fft.fft2(f, axes=(1, 2))
    kx = np.fft.fftfreq(R, 1.0 / R).reshape(R, 1)
    ky = np.fft.fftfreq(R, 1.0 / R).reshape(1, R)
    keep = (np.sqrt(kx ** 2 + ky ** 2) <= 12).reshape(1, R, R, 1)
    f = np.fft.ifft2(Fk * keep, axes=(1, 2)).real
    x[..., 0] = f / f.std()
    return x.astype(dtype)


def rel_l2(a: np.ndarray, b: np.ndarray) -> float:
    """||a-b||_2 / ||b||_2 over the whole tensor (SURVEY.md 8d accuracy metric)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def cast_params(p: Params, dtype) -> Params:
    return {k: v.astype(dtype) for k, v in p.items()}
