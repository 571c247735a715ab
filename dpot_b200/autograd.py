"""Training path of DPOTNet: forward with saved activations + hand-written backward.

``dpot_forward_train`` is what ``DPOTNet.forward`` runs when gradients are required
(train_temporal.py:206,227: ``im, cls_pred = model(xx) ... loss.backward()`` through all AR steps).
The graph is a chain of ``torch.autograd.Function``s; every tensor-sized arithmetic operation in
their forward AND backward is a kernel of libdpot_b200.so (GEMM engines, wgrad contraction, FFT
adjoints, GroupNorm backward, activation backward, pixel shuffle).  PyTorch contributes the autograd
tape, views/reshapes, and the weight-space re-parameterisations (a few MFLOP on parameter-sized
tensors: im2col weight permutation, coordinate-channel bias table, PatchEmbed-1x1 x TimeAggregator
folding), which are ordinary differentiable torch ops -- see DESIGN.md "training path".
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
from torch.autograd import Function

from . import _lib, ops
from ._lib import ACT_IDS, ACT_NONE, GEMM_AUTO, GemmArgs, WgradArgs, check, ptr

GROUPS = ops.GROUPS


def _s():
    return torch.cuda.current_stream().cuda_stream


def _lib_():
    return _lib.load()


# ------------------------------------------------------------------------------------------------ raw helpers
def _gemm(A, W, *, bias=None, act=ACT_NONE, residual=None, rowbias=None, out=None, pre=None, batch=1, strides=None,
          c_patch=None, engine=GEMM_AUTO):
    """A[M,K] (row stride lda), W[N,K]; batch/strides = (sA, sW, sC, sBias) for block-diagonal problems."""
    M = A.shape[0]
    g = GemmArgs()
    if batch == 1:
        N, K = W.shape
        ldw = W.stride(0)
    else:
        _, N, K = W.shape
        ldw = K
    if out is None:
        out = torch.empty((M, N * batch), device=A.device, dtype=torch.float32)
    g.A, g.lda, g.W, g.ldw, g.C, g.ldc = ptr(A), A.stride(0), ptr(W), ldw, ptr(out), out.stride(0) if c_patch is None else 0
    g.M, g.N, g.K = M, N, K
    g.bias, g.act = ptr(bias), act
    if rowbias is not None:
        g.rowbias, g.rowbias_period, g.ldrb = ptr(rowbias), rowbias.shape[0], rowbias.stride(0)
    if residual is not None:
        g.residual, g.ldr = ptr(residual), residual.stride(0)
    g.C_pre = ptr(pre)
    g.batch, g.engine, g.a_mode = batch, engine, _lib.A_PLAIN
    if batch > 1:
        g.strideA, g.strideW, g.strideC, g.strideBias = strides
    if c_patch is not None:
        g.c_mode, g.pX, g.pY, g.pT, g.pC, g.pP = _lib.A_PATCH, *c_patch
    check(_lib_().dpot_gemm(C.byref(g), _s()), "dpot_gemm")
    return out


def _wgrad(X, Y, N, K, *, batch=1, strides=None, patch=None):
    """dW[n,k] = sum_m X[m,n] Y[m,k]; batch/strides = (sX, sY, sW).

    Large plain problems (the channel-MLP / time-aggregation / ConvTranspose weights: 8.6 GFLOP each at B=16) run on
    the f16-split tcgen05 engine: the contraction index m must be the K-contiguous one, so both operands are transposed
    ([N, M], [K, M]) and split, then dW = gemm16(Xt, Yt) -- fp32-faithful like every other product (measured: 347 us on
    the CUDA-core dpot_wgrad -> ~60 us).  Batched / im2col forms stay on dpot_wgrad."""
    M = X.shape[0]
    if (batch == 1 and patch is None and M >= 256 and M % 8 == 0 and N >= 64 and K >= 64 and X.dim() == 2 and Y.dim() == 2
            and X.shape[1] == N and Y.shape[1] == K and _lib_().dpot_tc16_available()):
        from . import ops
        Xt, Yt = _transpose(X), _transpose(Y)        # [N, M], [K, M]
        # contraction chunks of <= 1024 tokens chained through the epilogue's residual input: the tensor core truncates
        # on every accumulate into TMEM, an error that grows with the chain length
        dW, CH = None, 1024
        for m0 in range(0, M, CH):
            m1 = min(M, m0 + CH)
            dW = ops.gemm16(ops.split_f16(Xt[:, m0:m1]), ops.split_f16(Yt[:, m0:m1]), residual=dW)
        return dW                                     # [N, K]
    a = WgradArgs()
    dW = torch.empty((batch, N, K) if batch > 1 else (N, K), device=X.device, dtype=torch.float32)
    a.X, a.ldx, a.Y, a.ldy, a.dW, a.ldw = ptr(X), X.stride(0), ptr(Y), (Y.stride(0) if patch is None else 0), ptr(dW), K
    a.M, a.N, a.K = X.shape[0], N, K
    a.batch = batch
    if batch > 1:
        a.strideX, a.strideY, a.strideW = strides
    if patch is not None:
        a.y_mode, a.pX, a.pY, a.pT, a.pC, a.pP = _lib.A_PATCH, *patch
    check(_lib_().dpot_wgrad(C.byref(a), _s()), "dpot_wgrad")
    return dW


def _colsum(X, N=None):
    M = X.shape[0]
    N = N or X.shape[1]
    out = torch.empty(N, device=X.device, dtype=torch.float32)
    check(_lib_().dpot_colsum(ptr(X), X.stride(0), M, N, ptr(out), 0, _s()), "dpot_colsum")
    return out


def _transpose(W, batch=1):
    if batch == 1:
        R, Cc = W.shape
        out = torch.empty((Cc, R), device=W.device, dtype=torch.float32)
        check(_lib_().dpot_transpose(ptr(W), W.stride(0), ptr(out), R, R, Cc, 1, 0, 0, _s()), "dpot_transpose")
    else:
        _, R, Cc = W.shape
        out = torch.empty((batch, Cc, R), device=W.device, dtype=torch.float32)
        check(_lib_().dpot_transpose(ptr(W), Cc, ptr(out), R, R, Cc, batch, R * Cc, R * Cc, _s()), "dpot_transpose")
    return out


def _act_bwd(dy, pre, act):
    out = torch.empty_like(dy)
    check(_lib_().dpot_act_bwd(ptr(dy), ptr(pre), act, dy.numel(), ptr(out), _s()), "dpot_act_bwd")
    return out


# ------------------------------------------------------------------------------------------------ autograd ops
class LinearFn(Function):
    """y = act(x @ W.T + bias + rowbias[m % period]) + residual."""

    @staticmethod
    def forward(ctx, x, W, bias, rowbias, residual, act):
        x = x.contiguous()
        W = W.contiguous()
        pre = torch.empty((x.shape[0], W.shape[0]), device=x.device) if act != ACT_NONE else None
        y = _gemm(x, W, bias=bias, act=act, residual=residual, rowbias=rowbias, pre=pre)
        ctx.save_for_backward(x, W, pre)
        ctx.act = act
        ctx.has = (bias is not None, rowbias is not None, residual is not None)
        ctx.period = rowbias.shape[0] if rowbias is not None else 0
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W, pre = ctx.saved_tensors
        dy = dy.contiguous()
        g = _act_bwd(dy, pre, ctx.act) if ctx.act != ACT_NONE else dy
        N, K = W.shape
        dx = _gemm(g, _transpose(W)) if ctx.needs_input_grad[0] else None
        dW = _wgrad(g, x, N, K) if ctx.needs_input_grad[1] else None
        db = _colsum(g) if (ctx.has[0] and ctx.needs_input_grad[2]) else None
        drb = None
        if ctx.has[1] and ctx.needs_input_grad[3]:
            B = g.shape[0] // ctx.period
            drb = _colsum(g.view(B, ctx.period * N)).view(ctx.period, N)
        dres = dy if (ctx.has[2] and ctx.needs_input_grad[4]) else None
        return dx, dW, db, drb, dres, None


class BlockDiagLinearFn(Function):
    """y[:, kap] = act(x[:, kap] @ Wc[kap].T + bc[kap]) for the nb blocks of the AFNO spectral MLP."""

    @staticmethod
    def forward(ctx, x, Wc, bc, act):
        nb, n, k = Wc.shape
        x = x.contiguous()
        pre = torch.empty((x.shape[0], nb * n), device=x.device) if act != ACT_NONE else None
        y = _gemm(x, Wc, bias=bc, act=act, pre=pre, batch=nb, strides=(k, n * k, n, n))
        ctx.save_for_backward(x, Wc, pre)
        ctx.act = act
        return y

    @staticmethod
    def backward(ctx, dy):
        x, Wc, pre = ctx.saved_tensors
        nb, n, k = Wc.shape
        dy = dy.contiguous()
        g = _act_bwd(dy, pre, ctx.act) if ctx.act != ACT_NONE else dy
        WcT = _transpose(Wc, batch=nb)                                    # [nb, k, n]
        dx = _gemm(g, WcT, batch=nb, strides=(n, n * k, k, 0))
        dWc = _wgrad(g, x, n, k, batch=nb, strides=(n, k, n * k))
        dbc = _colsum(g).view(nb, n)
        return dx, dWc, dbc, None


class PackAfnoFn(Function):
    """(w[2,nb,bs,bs], b[2,nb,bs]) -> real block form (Wc[nb,2bs,2bs], bc[nb,2bs]); models/dpot.py:72-94."""

    @staticmethod
    def forward(ctx, w, b):
        ctx.shape = w.shape
        return ops.pack_afno(w, b)

    @staticmethod
    def backward(ctx, dWc, dbc):
        _, nb, bs, _ = ctx.shape
        dw = torch.zeros(ctx.shape, device=dWc.device)
        db = torch.zeros((2, nb, bs), device=dWc.device)
        check(_lib_().dpot_unpack_afno_grad(ptr(dWc.contiguous()), ptr(dbc.contiguous()), nb, bs, ptr(dw), ptr(db), _s()),
              "dpot_unpack_afno_grad")
        return dw, db


class GroupNormFn(Function):
    """torch.nn.GroupNorm(8, E) on token-major x[B*n, E] (models/dpot.py:167,175)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, B, n, eps):
        x = x.contiguous()
        stats = ops.gn_stats(x, B, n)
        scale, shift = ops.gn_finalize(stats, gamma, beta, n, eps)
        y = torch.empty_like(x)
        check(_lib_().dpot_gn_apply(ptr(x), ptr(scale), ptr(shift), B, n, x.shape[1], ptr(y), _s()), "dpot_gn_apply")
        ctx.save_for_backward(x, stats, gamma)
        ctx.dims = (B, n, eps)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, stats, gamma = ctx.saved_tensors
        B, n, eps = ctx.dims
        E = x.shape[1]
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        dgamma = torch.zeros(E, device=x.device)
        dbeta = torch.zeros(E, device=x.device)
        scratch = torch.empty(2 * B * E + 2 * B * GROUPS, device=x.device)
        check(_lib_().dpot_gn_bwd(ptr(dy), ptr(x), ptr(stats), ptr(gamma), None, B, n, E, GROUPS, eps, ptr(scratch), ptr(dx),
                                  ptr(dgamma), ptr(dbeta), _s()), "dpot_gn_bwd")
        return dx, dgamma, dbeta, None, None, None


class SpectralFn(Function):
    """S = rfft2_ortho(x) restricted to the kept modes (models/dpot.py:59,62,70)."""

    @staticmethod
    def forward(ctx, x, B, h, nb, km1, km2):
        x = x.contiguous()
        E = x.shape[1]
        S = torch.empty((B * km1 * km2, 2 * E), device=x.device)
        check(_lib_().dpot_afno_fft_fwd(ptr(x), None, None, B, h, E, nb, km1, km2, ptr(S), 1.0, _s()), "dpot_afno_fft_fwd")
        ctx.dims = (B, h, nb, km1, km2, E)
        return S

    @staticmethod
    def backward(ctx, dS):
        B, h, nb, km1, km2, E = ctx.dims
        dx = torch.empty((B * h * h, E), device=dS.device)
        # adjoint(fwd) = inverse transform with interior columns weighted 1/2
        check(_lib_().dpot_afno_fft_inv(ptr(dS.contiguous()), None, None, None, B, h, E, nb, km1, km2, ptr(dx), None, GROUPS,
                                        0.5, _s()), "dpot_afno_fft_inv(adjoint)")
        return dx, None, None, None, None, None


class SpectralInvFn(Function):
    """f = irfft2_ortho(zero-padded O2) + skip (models/dpot.py:96-106)."""

    @staticmethod
    def forward(ctx, O2, skip, B, h, nb, km1, km2):
        E = skip.shape[1]
        f = torch.empty_like(skip)
        check(_lib_().dpot_afno_fft_inv(ptr(O2.contiguous()), ptr(skip.contiguous()), None, None, B, h, E, nb, km1, km2,
                                        ptr(f), None, GROUPS, 1.0, _s()), "dpot_afno_fft_inv")
        ctx.dims = (B, h, nb, km1, km2, E)
        return f

    @staticmethod
    def backward(ctx, df):
        B, h, nb, km1, km2, E = ctx.dims
        df = df.contiguous()
        dO2 = torch.empty((B * km1 * km2, 2 * E), device=df.device)
        # adjoint(inv) = forward transform with interior columns weighted 2
        check(_lib_().dpot_afno_fft_fwd(ptr(df), None, None, B, h, E, nb, km1, km2, ptr(dO2), 2.0, _s()),
              "dpot_afno_fft_fwd(adjoint)")
        return dO2, df, None, None, None, None, None


class PatchGemmFn(Function):
    """z1[(b,p,q,t), m] = act(im2col(x) @ W0p.T + rowbias0[(p,q,t)])  (PatchEmbed conv0, models/dpot.py:199-200)."""

    @staticmethod
    def forward(ctx, x, W0p, rowbias0, P, act):
        x = x.contiguous()
        B, X, Y, T, Cc = x.shape
        mid, K0 = W0p.shape
        h, w = X // P, Y // P
        M = B * h * w * T
        g = GemmArgs()
        z1 = torch.empty((M, mid), device=x.device)
        pre = torch.empty((M, mid), device=x.device)
        W0p = W0p.contiguous()
        rowbias0 = rowbias0.contiguous()
        g.A, g.W, g.ldw, g.C, g.ldc = ptr(x), ptr(W0p), K0, ptr(z1), mid
        g.M, g.N, g.K = M, mid, K0
        g.rowbias, g.rowbias_period, g.ldrb = ptr(rowbias0), h * w * T, mid
        g.act, g.C_pre = act, ptr(pre)
        g.batch, g.engine = 1, GEMM_AUTO
        g.a_mode, g.pX, g.pY, g.pT, g.pC, g.pP = _lib.A_PATCH, X, Y, T, Cc, P
        check(_lib_().dpot_gemm(C.byref(g), _s()), "dpot_gemm(patch)")
        ctx.save_for_backward(x, W0p, pre)
        ctx.geo = (X, Y, T, Cc, P)
        ctx.act = act
        return z1

    @staticmethod
    def backward(ctx, dz):
        x, W0p, pre = ctx.saved_tensors
        X, Y, T, Cc, P = ctx.geo
        mid, K0 = W0p.shape
        g = _act_bwd(dz.contiguous(), pre, ctx.act)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)   # stride == kernel: every input element belongs to exactly one patch
            _gemm(g, _transpose(W0p), out=dx, c_patch=(X, Y, T, Cc, P))
        dW0p = _wgrad(g, x, mid, K0, patch=(X, Y, T, Cc, P))
        period = (X // P) * (Y // P) * T
        B = g.shape[0] // period
        drb = _colsum(g.view(B, period * mid)).view(period, mid)
        return dx, dW0p, drb, None, None


class PixelShuffleFn(Function):
    """rows (b,p,q,u,v) x C -> field [B, X, Y, C] (ConvTranspose2d(k=s=P) output layout, models/dpot.py:316,397)."""

    @staticmethod
    def forward(ctx, y, B, h, P):
        Cc = y.shape[1]
        out = torch.empty((B, h * P, h * P, Cc), device=y.device)
        check(_lib_().dpot_pixel_shuffle(ptr(y.contiguous()), ptr(out), B, h, h, P, Cc, 1, _s()), "dpot_pixel_shuffle")
        ctx.dims = (B, h, P, Cc)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, h, P, Cc = ctx.dims
        dy = torch.empty((B * h * h * P * P, Cc), device=dout.device)
        check(_lib_().dpot_pixel_shuffle(ptr(dout.contiguous()), ptr(dy), B, h, h, P, Cc, 0, _s()), "dpot_pixel_shuffle")
        return dy, None, None, None


class SpatialMeanFn(Function):
    """cls token = mean over the latent grid (models/dpot.py:394)."""

    @staticmethod
    def forward(ctx, a, B, n):
        E = a.shape[1]
        tok = torch.empty((B, E), device=a.device)
        check(_lib_().dpot_spatial_mean(ptr(a.contiguous()), B, n, E, ptr(tok), _s()), "dpot_spatial_mean")
        ctx.dims = (B, n, E)
        return tok

    @staticmethod
    def backward(ctx, dtok):
        B, n, E = ctx.dims
        return (dtok / n).view(B, 1, E).expand(B, n, E).reshape(B * n, E), None, None


# ------------------------------------------------------------------------------------------------ model
def dpot_forward_train(net, x):
    """Differentiable DPOTNet.forward (models/dpot.py:364-403) on the kernels of libdpot_b200."""
    # the reference scripts address cuda:{args.gpu} without set_device (train_temporal.py:90): make x's device current
    # for every launch of the forward (backward runs on autograd's per-device thread, which already does)
    with torch.cuda.device(x.device):
        return _forward_train(net, x)


def _forward_train(net, x):
    if x.dtype != torch.float32:
        x = x.float()
    dev = x.device
    B, R, _, T, Cc = x.shape
    P, E, Co, To = net.patch_size, net.embed_dim, net.out_channels, net.out_timesteps
    h = R // P
    n = h * h
    act = ACT_IDS[net.act_name]
    km1, km2 = min(net.modes, h), min(net.modes, h // 2 + 1)

    # ---- input normalisation (normalize=True, models/dpot.py:366-370).  No shipped config trains with it, so the
    # per-sample statistics, the AdaIN affine below and the de-normalisation of the output are plain differentiable
    # torch ops (five elementwise / reduction kernels); the inference engine has its own fused kernels for them.
    mu = sigma = scale_mu = scale_sigma = None
    if net.normalize:
        mu = x.mean(dim=(1, 2, 3), keepdim=True)
        sigma = x.std(dim=(1, 2, 3), keepdim=True) + 1e-6
        x = (x - mu) / sigma
        ms = torch.cat([mu, sigma], dim=-1).reshape(B, 2 * Cc)
        scale_mu = LinearFn.apply(ms, net.scale_feats_mu.weight, net.scale_feats_mu.bias, None, None, ACT_NONE)
        scale_sigma = LinearFn.apply(ms, net.scale_feats_sigma.weight, net.scale_feats_sigma.bias, None, None, ACT_NONE)

    # ---- weight-space re-parameterisations (differentiable torch ops on parameter-sized tensors)
    pe0, pe2 = net.patch_embed.proj[0], net.patch_embed.proj[2]
    mid = pe0.out_channels
    W0 = pe0.weight                                                        # [mid, C+3, P, P]
    W0p = W0[:, :Cc].permute(0, 2, 3, 1).reshape(mid, P * P * Cc)          # k = (u, v, c)
    gx = torch.tensor(np.linspace(0, 1, R), dtype=torch.float, device=dev)  # get_grid_3d, :350-357
    gt = torch.tensor(np.linspace(0, 1, T), dtype=torch.float, device=dev)
    gpx = gx.view(h, P)                                                    # [p, u]
    bx = torch.einsum('muv,pu->mp', W0[:, Cc], gpx)                        # [mid, h]
    by = torch.einsum('muv,qv->mq', W0[:, Cc + 1], gpx)
    bt = W0[:, Cc + 2].sum(dim=(1, 2))[:, None] * gt[None, :]              # [mid, T]
    rowbias0 = (pe0.bias[None, None, None, :] + bx.t()[:, None, None, :] + by.t()[None, :, None, :] +
                bt.t()[None, None, :, :]).reshape(n * T, mid)
    ta = net.time_agg_layer
    if ta.type == 'exp_mlp':
        tt = torch.linspace(0, 1, T).unsqueeze(-1).to(dev)
        temb = torch.cos(tt * ta.gamma)                                    # K=1 matmul == product (:230-231)
    else:
        temb = torch.ones((T, E), device=dev)
    W2 = pe2.weight.reshape(E, mid)
    wt = ta.w * temb.unsqueeze(-1)                                         # [T, E(i), E(j)]
    WeffT = torch.einsum('im,tij->jtm', W2, wt).reshape(E, T * mid)       # folded conv1x1 x time aggregation
    bias_eff = torch.einsum('ip,ij->pj', pe2.bias[:, None] + net.pos_embed[0].reshape(E, n), wt.sum(0))

    # ---- PatchEmbed conv0 + folded aggregation
    z1 = PatchGemmFn.apply(x, W0p, rowbias0, P, act)                       # [(b,pq,t), mid]
    a = LinearFn.apply(z1.view(B * n, T * mid), WeffT, None, bias_eff, None, ACT_NONE)
    if net.normalize:                                                      # AdaIN, models/dpot.py:386-387
        a = (a.view(B, n, E) * scale_sigma.view(B, 1, E) + scale_mu.view(B, 1, E)).reshape(B * n, E)

    # ---- blocks (models/dpot.py:165-180, double_skip=False)
    for blk in net.blocks:
        n1 = GroupNormFn.apply(a, blk.norm1.weight, blk.norm1.bias, B, n, blk.norm1.eps)
        flt = blk.filter
        Wc1, bc1 = PackAfnoFn.apply(flt.w1, flt.b1)
        Wc2, bc2 = PackAfnoFn.apply(flt.w2, flt.b2)
        S = SpectralFn.apply(n1, B, h, flt.num_blocks, km1, km2)
        O1 = BlockDiagLinearFn.apply(S, Wc1, bc1, act)
        O2 = BlockDiagLinearFn.apply(O1, Wc2, bc2, ACT_NONE)
        f = SpectralInvFn.apply(O2, n1, B, h, flt.num_blocks, km1, km2)
        n2 = GroupNormFn.apply(f, blk.norm2.weight, blk.norm2.bias, B, n, blk.norm2.eps)
        fc1, fc2 = blk.mlp[0], blk.mlp[2]
        hid = fc1.out_channels
        hdn = LinearFn.apply(n2, fc1.weight.reshape(hid, E), fc1.bias, None, None, act)
        a = LinearFn.apply(hdn, fc2.weight.reshape(E, hid), fc2.bias, None, a, ACT_NONE)

    # ---- classification head (models/dpot.py:394-395)
    ch = net.cls_head
    tok = SpatialMeanFn.apply(a, B, n)
    c1 = LinearFn.apply(tok, ch[0].weight, ch[0].bias, None, None, act)
    c2 = LinearFn.apply(c1, ch[2].weight, ch[2].bias, None, None, act)
    cls_pred = LinearFn.apply(c2, ch[4].weight, ch[4].bias, None, None, ACT_NONE)

    # ---- output head (models/dpot.py:315-321,397-398)
    ol = net.out_layer
    old = ol[0].out_channels
    WtT = ol[0].weight.permute(2, 3, 1, 0).reshape(P * P * old, E)         # rows (u,v,o)
    bias_t = ol[0].bias.repeat(P * P)
    y1 = LinearFn.apply(a, WtT, bias_t, None, None, act).view(B * n * P * P, old)
    y2 = LinearFn.apply(y1, ol[2].weight.reshape(old, old), ol[2].bias, None, None, act)
    y3 = LinearFn.apply(y2, ol[4].weight.reshape(Co * To, old), ol[4].bias, None, None, ACT_NONE)
    out = PixelShuffleFn.apply(y3, B, h, P).view(B, R, R, To, Co)
    if net.normalize:                                                      # models/dpot.py:400-401
        out = out * sigma + mu
    return out, cls_pred


# ------------------------------------------------------------------------------------------------ standalone modules
# The reference exports AFNO2D / Block / Mlp / PatchEmbed / TimeAggregator as ordinary differentiable nn.Modules
# (models/dpot.py:29-234).  Under grad their forwards run on the same per-operator Functions as the generic training
# path above; layout changes (NCHW <-> token-major) and parameter re-shapes are differentiable torch views.
def _afno_tokens(flt, a, B, h, skip):
    """AFNO2D on token-major a[B*h*h, E] (models/dpot.py:51-110): irfft2(MLP(rfft2(a))) + skip."""
    km1, km2 = min(flt.modes, h), min(flt.modes, h // 2 + 1)
    act = ACT_IDS[flt.act_name]
    Wc1, bc1 = PackAfnoFn.apply(flt.w1, flt.b1)
    Wc2, bc2 = PackAfnoFn.apply(flt.w2, flt.b2)
    S = SpectralFn.apply(a, B, h, flt.num_blocks, km1, km2)
    O1 = BlockDiagLinearFn.apply(S, Wc1, bc1, act)
    O2 = BlockDiagLinearFn.apply(O1, Wc2, bc2, ACT_NONE)
    return SpectralInvFn.apply(O2, skip, B, h, flt.num_blocks, km1, km2)


def afno2d_train(flt, x):
    with torch.cuda.device(x.device):
        a = x.permute(0, 2, 3, 1) if flt.channel_first else x
        B, H, W, Cc = a.shape
        a = a.reshape(B * H * W, Cc).contiguous().float()
        f = _afno_tokens(flt, a, B, H, a).reshape(B, H, W, Cc)
        return f.permute(0, 3, 1, 2) if flt.channel_first else f


def block_train(blk, x):
    """Block.forward (models/dpot.py:165-180) on x[B, E, H, W]."""
    with torch.cuda.device(x.device):
        B, E, H, W = x.shape
        n = H * W
        act = ACT_IDS[blk.act_name] if hasattr(blk, "act_name") else ACT_IDS[blk.filter.act_name]
        a = x.permute(0, 2, 3, 1).reshape(B * n, E).contiguous().float()
        n1 = GroupNormFn.apply(a, blk.norm1.weight, blk.norm1.bias, B, n, blk.norm1.eps)
        f = _afno_tokens(blk.filter, n1, B, H, n1)
        res = a
        if blk.double_skip:                       # :171-173
            f = f + a
            res = f
        n2 = GroupNormFn.apply(f, blk.norm2.weight, blk.norm2.bias, B, n, blk.norm2.eps)
        fc1, fc2 = blk.mlp[0], blk.mlp[2]
        hid = fc1.out_channels
        hdn = LinearFn.apply(n2, fc1.weight.reshape(hid, E), fc1.bias, None, None, act)
        out = LinearFn.apply(hdn, fc2.weight.reshape(E, hid), fc2.bias, None, res, ACT_NONE)
        return out.reshape(B, H, W, E).permute(0, 3, 1, 2)


def mlp_train(mod, x):
    with torch.cuda.device(x.device):
        lead = x.shape[:-1]
        a = x.reshape(-1, x.shape[-1]).contiguous().float()
        hdn = LinearFn.apply(a, mod.fc1.weight, mod.fc1.bias, None, None, ACT_IDS[mod.act_name])
        return LinearFn.apply(hdn, mod.fc2.weight, mod.fc2.bias, None, None, ACT_NONE).reshape(*lead, -1)


def patch_embed_train(pe, x):
    """PatchEmbed.forward (models/dpot.py:203-209) on NCHW frames."""
    with torch.cuda.device(x.device):
        B, Cc, H, W = x.shape
        P = pe.patch_size[0]
        c0, c2 = pe.proj[0], pe.proj[2]
        mid = c0.out_channels
        h, w = H // P, W // P
        xf = x.permute(0, 2, 3, 1).reshape(B, H, W, 1, Cc).contiguous().float()
        W0p = c0.weight.permute(0, 2, 3, 1).reshape(mid, P * P * Cc)
        rb = c0.bias.reshape(1, mid).expand(h * w, mid)
        z1 = PatchGemmFn.apply(xf, W0p, rb, P, ACT_IDS[pe.act_name])
        out = LinearFn.apply(z1, c2.weight.reshape(c2.out_channels, mid), c2.bias, None, None, ACT_NONE)
        return out.reshape(B, h, w, -1).permute(0, 3, 1, 2)


def time_agg_train(ta, x):
    """TimeAggregator.forward (models/dpot.py:226-234): out[..., j] = sum_{t,i} w[t,i,j] x[..., t, i] temb[t,i]."""
    with torch.cuda.device(x.device):
        lead, T, E = x.shape[:-2], x.shape[-2], x.shape[-1]
        if ta.type == 'exp_mlp':
            t = torch.linspace(0, 1, T).unsqueeze(-1).to(x.device)
            temb = torch.cos(t * ta.gamma)
        else:
            temb = torch.ones((T, E), device=x.device)
        Wt = (ta.w * temb.unsqueeze(-1)).reshape(T * E, E).t()
        a = x.reshape(-1, T * E).contiguous().float()
        return LinearFn.apply(a, Wt, None, None, None, ACT_NONE).reshape(*lead, E)
