"""Component-wise backward check of dpot_b200.autograd against float64 torch references (GPU)."""
import os, sys
import numpy as np
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpot_b200 import autograd as AG, ops
from dpot_b200._lib import ACT_IDS, ACT_NONE

torch.manual_seed(0)
dev = "cuda"


def rel(a, b):
    a = a.double(); b = b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def check(name, ours_fn, ref_fn, inputs):
    """inputs: list of float32 cuda tensors (requires_grad).  ref_fn gets float64 clones."""
    ins = [t.clone().requires_grad_(True) for t in inputs]
    ins64 = [t.double().clone().requires_grad_(True) for t in inputs]
    y = ours_fn(*ins); y64 = ref_fn(*ins64)
    ys = y if isinstance(y, tuple) else (y,); ys64 = y64 if isinstance(y64, tuple) else (y64,)
    msg = [f"fwd={rel(a, b):.1e}" for a, b in zip(ys, ys64)]
    ws = [torch.randn_like(a) for a in ys]
    (sum((a * w).sum() for a, w in zip(ys, ws))).backward()
    (sum((a * w.double()).sum() for a, w in zip(ys64, ws))).backward()
    for i, (a, b) in enumerate(zip(ins, ins64)):
        msg.append(f"d{i}={rel(a.grad, b.grad):.1e}" if a.grad is not None else f"d{i}=None")
    print(f"{name:28s}", " ".join(msg))


gelu = ACT_IDS["gelu"]
# ---- LinearFn
M, K, N = 96, 40, 24
x = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) / 6; b = torch.randn(N, device=dev)
rb = torch.randn(12, N, device=dev); res = torch.randn(M, N, device=dev)
check("LinearFn(gelu,all)", lambda x, W, b, rb, res: AG.LinearFn.apply(x, W, b, rb, res, gelu),
      lambda x, W, b, rb, res: F.gelu(x @ W.t() + b + rb.repeat(M // 12, 1)) + res, [x, W, b, rb, res])
check("LinearFn(none)", lambda x, W: AG.LinearFn.apply(x, W, None, None, None, ACT_NONE), lambda x, W: x @ W.t(), [x, W])
# ---- BlockDiag
nb, n_, k_ = 4, 8, 8
xs = torch.randn(M, nb * k_, device=dev); Wc = torch.randn(nb, n_, k_, device=dev) / 3; bc = torch.randn(nb, n_, device=dev)
def bd_ref(x, Wc, bc):
    return F.gelu(torch.cat([x[:, i * k_:(i + 1) * k_] @ Wc[i].t() + bc[i] for i in range(nb)], 1))
check("BlockDiagLinearFn", lambda x, Wc, bc: AG.BlockDiagLinearFn.apply(x, Wc, bc, gelu), bd_ref, [xs, Wc, bc])
# ---- PackAfno
bs = 4
w = torch.randn(2, nb, bs, bs, device=dev); bb = torch.randn(2, nb, bs, device=dev)
def pack_ref(w, b):
    wr, wi = w[0], w[1]   # [nb, in, out]
    top = torch.cat([wr.transpose(1, 2), -wi.transpose(1, 2)], 2)   # out real rows: [n, k]: k<bs wr, k>=bs -wi
    bot = torch.cat([wi.transpose(1, 2), wr.transpose(1, 2)], 2)
    return torch.cat([top, bot], 1), torch.cat([b[0], b[1]], 1)
check("PackAfnoFn", lambda w, b: AG.PackAfnoFn.apply(w, b), pack_ref, [w, bb])
# ---- GroupNorm
B, h, E = 3, 4, 32
nn_ = h * h
xg = torch.randn(B * nn_, E, device=dev) * 2 + 0.5; gam = torch.randn(E, device=dev); bet = torch.randn(E, device=dev)
def gn_ref(x, g, b):
    xc = x.view(B, nn_, E).permute(0, 2, 1)
    return F.group_norm(xc, 8, g, b, 1e-5).permute(0, 2, 1).reshape(B * nn_, E)
check("GroupNormFn", lambda x, g, b: AG.GroupNormFn.apply(x, g, b, B, nn_, 1e-5), gn_ref, [xg, gam, bet])
# ---- Spectral fwd / inv
for (h, km) in [(4, 2), (8, 3), (8, 8), (16, 16)]:
    nn_ = h * h; nbk = 4; E = 32; bsz = E // nbk
    km1, km2 = min(km, h), min(km, h // 2 + 1)
    xs_ = torch.randn(B * nn_, E, device=dev)
    def spec_ref(x):
        Fx = torch.fft.rfft2(x.view(B, h, h, E), dim=(1, 2), norm="ortho")[:, :km1, :km2]      # B,km1,km2,E
        Fx = Fx.reshape(B, km1, km2, nbk, bsz)
        return torch.cat([Fx.real, Fx.imag], -1).reshape(B * km1 * km2, 2 * E)
    check(f"SpectralFn h{h} m{km}", lambda x: AG.SpectralFn.apply(x, B, h, nbk, km1, km2), spec_ref, [xs_])
    O2 = torch.randn(B * km1 * km2, 2 * E, device=dev); skip = torch.randn(B * nn_, E, device=dev)
    def inv_ref(O2, skip):
        Z = O2.view(B, km1, km2, nbk, 2, bsz)
        Zc = torch.complex(Z[..., 0, :], Z[..., 1, :]).reshape(B, km1, km2, E)
        full = torch.zeros(B, h, h // 2 + 1, E, dtype=Zc.dtype, device=dev)
        full[:, :km1, :km2] = Zc
        return torch.fft.irfft2(full, s=(h, h), dim=(1, 2), norm="ortho").reshape(B * nn_, E) + skip
    check(f"SpectralInvFn h{h} m{km}", lambda O2, s: AG.SpectralInvFn.apply(O2, s, B, h, nbk, km1, km2), inv_ref, [O2, skip])
# ---- PatchGemm
B, R, T, Cc, P, mid = 2, 16, 3, 2, 4, 5
h = R // P
xf = torch.randn(B, R, R, T, Cc, device=dev); W0p = torch.randn(mid, P * P * Cc, device=dev) / 4
rb0 = torch.randn(h * h * T, mid, device=dev)
def patch_ref(x, W0p, rb0):
    xp = x.view(B, h, P, h, P, T, Cc).permute(0, 1, 3, 5, 2, 4, 6).reshape(B * h * h * T, P * P * Cc)   # rows (b,p,q,t), k=(u,v,c)
    return F.gelu(xp @ W0p.t() + rb0.repeat(B, 1))
check("PatchGemmFn", lambda x, W, rb: AG.PatchGemmFn.apply(x, W, rb, P, gelu), patch_ref, [xf, W0p, rb0])
# ---- PixelShuffle
y = torch.randn(B * h * h * P * P, 3, device=dev)
def ps_ref(y):
    return y.view(B, h, h, P, P, 3).permute(0, 1, 3, 2, 4, 5).reshape(B, h * P, h * P, 3)
check("PixelShuffleFn", lambda y: AG.PixelShuffleFn.apply(y, B, h, P), ps_ref, [y])
# ---- SpatialMean
a = torch.randn(B * h * h, 8, device=dev)
check("SpatialMeanFn", lambda a: AG.SpatialMeanFn.apply(a, B, h * h), lambda a: a.view(B, h * h, 8).mean(1), [a])
