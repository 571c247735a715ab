"""Torch-tensor wrappers over the C ABI (one python function per entry point of
include/dpot_b200.h).  These only marshal pointers/shapes and enqueue on the current CUDA stream;
every arithmetic operation happens inside libdpot_b200.so.  There is no fallback: a CPU tensor
or a missing library raises."""
from __future__ import annotations

import ctypes as C
import functools
from typing import Optional

import torch

from . import _lib
from ._lib import ACT_IDS, ACT_NONE, GEMM_AUTO, GemmArgs, check, ptr

GROUPS = 8  # torch.nn.GroupNorm(8, width), models/dpot.py:142,152


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _first_cuda(args):
    for a in args:
        if isinstance(a, torch.Tensor):
            if a.is_cuda:
                return a
        elif isinstance(a, (list, tuple)):
            t = _first_cuda(a)
            if t is not None:
                return t
    return None


def _on_device(fn):
    """Run `fn` with the CUDA device of its first CUDA tensor argument current: the reference scripts address
    cuda:{args.gpu} without ever calling set_device (train_temporal.py:90), and a kernel launched while another
    device is current would run against foreign pointers on the wrong GPU's stream."""
    @functools.wraps(fn)
    def wrapped(*args, **kw):
        t = _first_cuda(args)
        if t is None:
            t = _first_cuda(tuple(kw.values()))
        if t is None or t.device.index == torch.cuda.current_device():
            return fn(*args, **kw)
        with torch.cuda.device(t.device):
            return fn(*args, **kw)
    return wrapped


def _need_cuda(*ts):
    for t in ts:
        if t is not None and (not t.is_cuda or t.dtype != torch.float32):
            raise RuntimeError("dpot_b200: tensors must be float32 CUDA tensors (the hot path has no CPU fallback); "
                               f"got device={t.device} dtype={t.dtype}")


def act_id(act) -> int:
    if act is None:
        return ACT_NONE
    if isinstance(act, int):
        return act
    return ACT_IDS[act]


@_on_device
def gemm(A: torch.Tensor, W: torch.Tensor, *, bias=None, act=None, residual=None, rowbias=None,
         a_scale=None, a_shift=None, a_rows_per_sample=0, c_scale=None, c_shift=None, c_rows_per_sample=0,
         out: Optional[torch.Tensor] = None, engine: int = GEMM_AUTO) -> torch.Tensor:
    """C = act(A' @ W.T + bias + rowbias[m % period]) * c_scale + c_shift + residual; A[M,K], W[N,K]."""
    _need_cuda(A, W, bias, residual, rowbias, a_scale, a_shift, c_scale, c_shift, out)
    assert A.dim() == 2 and W.dim() == 2 and A.shape[1] == W.shape[1]
    assert A.stride(1) == 1 and W.stride(1) == 1
    M, K = A.shape
    N = W.shape[0]
    if out is None:
        out = torch.empty((M, N), device=A.device, dtype=torch.float32)
    g = GemmArgs()
    g.A, g.lda, g.W, g.ldw, g.C, g.ldc = ptr(A), A.stride(0), ptr(W), W.stride(0), ptr(out), out.stride(0)
    g.M, g.N, g.K = M, N, K
    g.bias = ptr(bias)
    if rowbias is not None:
        g.rowbias, g.rowbias_period, g.ldrb = ptr(rowbias), rowbias.shape[0], rowbias.stride(0)
    if residual is not None:
        g.residual, g.ldr = ptr(residual), residual.stride(0)
    g.act = act_id(act)
    if a_scale is not None:
        g.a_scale, g.a_shift, g.a_rows_per_sample = ptr(a_scale), ptr(a_shift), a_rows_per_sample
    if c_scale is not None:
        g.c_scale, g.c_shift, g.c_rows_per_sample = ptr(c_scale), ptr(c_shift), c_rows_per_sample
    g.batch, g.engine, g.a_mode = 1, engine, _lib.A_PLAIN
    check(_lib.load().dpot_gemm(C.byref(g), _stream()), "dpot_gemm")
    return out


@_on_device
def split_f16(x: torch.Tensor, scale=None, shift=None, rows_per_sample: int = 0) -> torch.Tensor:
    """fp32 [M,K] -> split fp16 [M, 2K] (DPOT_FMT_HL16: columns [0,K) = hi, [K,2K) = lo * 2048)."""
    _need_cuda(x, scale, shift)
    assert x.dim() == 2 and x.stride(1) == 1
    M, K = x.shape
    out = torch.empty((M, 2 * K), device=x.device, dtype=torch.float16)
    check(_lib.load().dpot_split_f16(ptr(x), x.stride(0), M, K, ptr(scale), ptr(shift), rows_per_sample, ptr(out),
                                     2 * K, K, _stream()), "dpot_split_f16")
    return out


@_on_device
def gemm16(A16: torch.Tensor, W16: torch.Tensor, *, bias=None, act=None, residual=None, rowbias=None, c_scale=None,
           c_shift=None, c_rows_per_sample=0, out16: bool = False, nb: int = 1, stats=None):
    """The f16-split tcgen05 engine on pre-split operands A16[M, 2*Kt], W16[nb*N, 2*K] (see split_f16).
    nb > 1: block-diagonal form, A16 hi plane [M, nb*K], W16 [nb, N, 2K], bias [nb, N] -> C[M, nb*N].
    stats=(groups, rows_per_sample): also return the GroupNorm statistics [M/rps, groups, 2] (double) of the result."""
    _need_cuda(bias, residual, rowbias, c_scale, c_shift)
    assert A16.dtype == torch.float16 and W16.dtype == torch.float16 and A16.is_cuda and W16.is_cuda
    M, Kt = A16.shape[0], A16.shape[1] // 2
    K = W16.shape[-1] // 2
    N = W16.shape[-2] if nb > 1 else W16.shape[0]
    assert Kt == nb * K
    Nt = nb * N
    g = GemmArgs()
    if out16 == "g32":      # one [hi 32 | lo 32] record per (row, 32-column group): DPOT_FMT_HL16G32
        out = torch.empty((M, 2 * Nt), device=A16.device, dtype=torch.float16)
        g.C, g.ldc, g.c_fmt = ptr(out), 2 * Nt, _lib.FMT_HL16G32
    elif out16:
        out = torch.empty((M, 2 * Nt), device=A16.device, dtype=torch.float16)
        g.C, g.ldc, g.c_fmt, g.c_lo_off = ptr(out), 2 * Nt, _lib.FMT_HL16, Nt
    else:
        out = torch.empty((M, Nt), device=A16.device, dtype=torch.float32)
        g.C, g.ldc = ptr(out), Nt
    g.A, g.lda, g.a_fmt, g.a_lo_off = ptr(A16), 2 * Kt, _lib.FMT_HL16, Kt
    g.W, g.ldw, g.w_fmt, g.w_lo_off = ptr(W16), 2 * K, _lib.FMT_HL16, K
    g.M, g.N, g.K = M, N, K
    g.bias, g.act = ptr(bias), act_id(act)
    if rowbias is not None:
        g.rowbias, g.rowbias_period, g.ldrb = ptr(rowbias), rowbias.shape[0], rowbias.stride(0)
    if residual is not None:
        g.residual, g.ldr = ptr(residual), residual.stride(0)
    if c_scale is not None:
        g.c_scale, g.c_shift, g.c_rows_per_sample = ptr(c_scale), ptr(c_shift), c_rows_per_sample
    g.batch, g.engine, g.a_mode = nb, _lib.GEMM_TC16, _lib.A_PLAIN
    if nb > 1:
        g.strideA, g.strideW, g.strideC, g.strideBias = K, N * 2 * K, N, N
    st = None
    if stats is not None:
        groups, rps = stats
        st = torch.empty((M // rps, groups, 2), device=A16.device, dtype=torch.float64)
        g.out_stats, g.stats_groups, g.stats_rows_per_sample = ptr(st), groups, rps
    lib = _lib.load()
    kc = lib.dpot_tc16_set_chain(-1)
    if out16 and nb == 1 and 0 < kc < K:    # split result: the chained partial sums need an fp32 buffer of their own
        scratch = torch.empty((M, Nt), device=A16.device, dtype=torch.float32)
        check(lib.dpot_gemm_chained(C.byref(g), kc, ptr(scratch), Nt, _stream()), "dpot_gemm_chained(tc16)")
    else:
        check(lib.dpot_gemm(C.byref(g), _stream()), "dpot_gemm(tc16)")
    return out if st is None else (out, st)


@_on_device
def gemm16_bw(A16: torch.Tensor, W16: torch.Tensor, M: int, N: int, K: int, *, a_trans=False, w_trans=False, nb: int = 1,
              k_chunk: int = 0, out16: bool = False, dact_src=None, dact=None, pre: bool = False, act=None, bias=None):
    """The backward-pass forms of the f16-split engine (dpot_gemm_args a_trans / w_trans / k_split, ABI 2).
    Per problem C[M, N] = A[M, K] W[N, K]^T with A16 stored [M, 2*nb*K] (or transposed [K, 2*nb*M]), W16 stored
    [nb*N, 2K] (or transposed [K, 2*nb*N]); nb problems side by side (block-diagonal AFNO weights, except W16 plain:
    [nb, N, 2K]).  k_chunk > 0: the contraction is cut into ceil(K / k_chunk) chunks -> returns the partial results
    [chunks, ...] (the caller sums them).  dact_src [M, nb*N] fp32: result *= act'(dact_src).  pre: also return the
    fp32 pre-activation."""
    g = GemmArgs()
    dev = A16.device
    Nt = nb * N
    ks = (K + k_chunk - 1) // k_chunk if k_chunk > 0 else 1
    if w_trans and nb > 1 and not a_trans:      # dgrad of the block-diagonal layer: C[M, nb*N]
        shape = (M, Nt)
    elif a_trans:                                # wgrad: one [M, N] matrix per problem
        shape = (nb, M, N)
    else:
        shape = (M, Nt)
    lead = (ks,) if ks > 1 else ()
    if out16:
        out = torch.empty(lead + shape[:-1] + (2 * shape[-1],), device=dev, dtype=torch.float16)
        g.c_fmt, g.c_lo_off = _lib.FMT_HL16, shape[-1]
        g.ldc = 2 * shape[-1]
    else:
        out = torch.empty(lead + shape, device=dev, dtype=torch.float32)
        g.ldc = shape[-1]
    g.C = ptr(out)
    g.M, g.N, g.K = M, N, K
    g.a_fmt = g.w_fmt = _lib.FMT_HL16
    g.A, g.W = ptr(A16), ptr(W16)
    g.lda, g.a_lo_off = A16.stride(0), A16.shape[1] // 2
    g.ldw, g.w_lo_off = W16.stride(0), W16.shape[-1] // 2
    g.a_trans, g.w_trans = int(a_trans), int(w_trans)
    g.batch, g.engine, g.a_mode = nb, _lib.GEMM_TC16, _lib.A_PLAIN
    g.act = act_id(act)
    g.bias = ptr(bias)
    if nb > 1:
        g.strideA = M if a_trans else K
        g.strideW = N if (w_trans and a_trans) else (K * 2 * N if w_trans else N * 2 * K)   # activation columns / stacked weights
        g.strideC = M * N * (2 if out16 else 1) if a_trans else N
        g.strideBias = N
    if ks > 1:
        g.k_split, g.k_chunk, g.strideC_split = ks, k_chunk, out[0].numel()
    pre_t = None
    if pre:
        pre_t = torch.empty(shape, device=dev, dtype=torch.float32)
        g.C_pre, g.ld_pre, g.stride_pre = ptr(pre_t), shape[-1], (M * N if a_trans else N)
    if dact_src is not None:
        g.dact_src, g.ld_dact, g.stride_dact, g.dact = ptr(dact_src), dact_src.stride(0), N, act_id(dact)
    check(_lib.load().dpot_gemm(C.byref(g), _stream()), "dpot_gemm(tc16 backward form)")
    return (out, pre_t) if pre else out


def unsplit_f16(x16: torch.Tensor) -> torch.Tensor:
    """Inverse of split_f16 (test helper: plain torch arithmetic on the stored halves)."""
    K = x16.shape[1] // 2
    return x16[:, :K].float() + x16[:, K:].float() / 2048.0


@_on_device
def gemm_batched_cols(A: torch.Tensor, W: torch.Tensor, bias: torch.Tensor, nb: int, *, act=None,
                      out: Optional[torch.Tensor] = None, engine: int = GEMM_AUTO) -> torch.Tensor:
    """Block-diagonal GEMM of the AFNO spectral MLP: A[M, nb*k], W[nb, n, k], bias[nb, n] -> C[M, nb*n]."""
    _need_cuda(A, W, bias, out)
    M = A.shape[0]
    _, n, k = W.shape
    assert A.shape[1] == nb * k and W.is_contiguous() and bias.is_contiguous() and A.stride(1) == 1
    if out is None:
        out = torch.empty((M, nb * n), device=A.device, dtype=torch.float32)
    g = GemmArgs()
    g.A, g.lda, g.W, g.ldw, g.C, g.ldc = ptr(A), A.stride(0), ptr(W), k, ptr(out), out.stride(0)
    g.M, g.N, g.K = M, n, k
    g.bias, g.act = ptr(bias), act_id(act)
    g.batch, g.strideA, g.strideW, g.strideC, g.strideBias = nb, k, n * k, n, n
    g.engine, g.a_mode = engine, _lib.A_PLAIN
    check(_lib.load().dpot_gemm(C.byref(g), _stream()), "dpot_gemm(batched)")
    return out


@_on_device
def patch_gemm(x: torch.Tensor, W0p: torch.Tensor, rowbias0: torch.Tensor, P: int, act, Kp: int, *,
               a_scale=None, a_shift=None, engine: int = GEMM_AUTO) -> torch.Tensor:
    """PatchEmbed conv0 + act as an im2col GEMM: x[B,X,Y,T,C] -> z1[B*n, Kp] (column t*mid+m)."""
    _need_cuda(x, W0p, rowbias0, a_scale, a_shift)
    B, X, Y, T, Cc = x.shape
    assert x.is_contiguous()
    mid, K0 = W0p.shape
    h, w = X // P, Y // P
    z1 = torch.zeros((B * h * w, Kp), device=x.device, dtype=torch.float32)
    g = GemmArgs()
    g.A, g.W, g.ldw, g.C, g.ldc = ptr(x), ptr(W0p), K0, ptr(z1), mid
    g.M, g.N, g.K = B * h * w * T, mid, K0
    g.rowbias, g.rowbias_period, g.ldrb = ptr(rowbias0), h * w * T, mid
    g.act = act_id(act)
    g.c_group, g.c_group_stride = T, Kp
    if a_scale is not None:
        g.a_scale, g.a_shift, g.a_rows_per_sample = ptr(a_scale), ptr(a_shift), h * w * T
    g.batch, g.engine = 1, engine
    g.a_mode, g.pX, g.pY, g.pT, g.pC, g.pP = _lib.A_PATCH, X, Y, T, Cc, P
    check(_lib.load().dpot_gemm(C.byref(g), _stream()), "dpot_gemm(patch)")
    return z1


@_on_device
def patch_embed(x: torch.Tensor, W0p: torch.Tensor, rowbias0: torch.Tensor, P: int, act, Kp: int, *, t0: int = 0,
                a_scale=None, a_shift=None, out16: bool = False) -> torch.Tensor:
    """dpot_patch_embed: x[B,X,Y,T,C] (a ring in time, logical frame t = slot (t+t0)%T) -> z1[B*n, Kp] fp32, or the
    split-fp16 form [B*n, 2*Kp] halves (out16)."""
    _need_cuda(x, W0p, rowbias0, a_scale, a_shift)
    B, X, Y, T, Cc = x.shape
    assert x.is_contiguous()
    mid = W0p.shape[0]
    n = (X // P) * (Y // P)
    if out16:
        z1 = torch.zeros((B * n, 2 * Kp), device=x.device, dtype=torch.float16)
    else:
        z1 = torch.zeros((B * n, Kp), device=x.device, dtype=torch.float32)
    check(_lib.load().dpot_patch_embed(ptr(x), t0, ptr(W0p), ptr(rowbias0), ptr(a_scale), ptr(a_shift), B, X, Y, T, Cc, P,
                                       mid, act_id(act), ptr(z1), Kp, _lib.FMT_HL16 if out16 else _lib.FMT_F32, _stream()),
          "dpot_patch_embed")
    return z1


@_on_device
def out_tail(Y1: torch.Tensor, w2, b2, w4, b4, B: int, h: int, w: int, P: int, act, *, mu=None, sigma=None,
             Co: int = 0) -> torch.Tensor:
    """dpot_out_tail: Y1[(b,p,q), (u,v,o)] -> per-pixel act(W2 y + b2) -> W4 . + b4 -> out[B, h*P, w*P, nout]."""
    _need_cuda(Y1, w2, b2, w4, b4, mu, sigma)
    nout, old = w4.shape[0], w4.shape[1]
    out = torch.empty((B, h * P, w * P, nout), device=Y1.device, dtype=torch.float32)
    check(_lib.load().dpot_out_tail(ptr(Y1), ptr(w2), ptr(b2), ptr(w4), ptr(b4), B, h, w, P, old, nout, act_id(act),
                                    ptr(mu), ptr(sigma), Co or nout, ptr(out), _stream()), "dpot_out_tail")
    return out


def to_g32(y: torch.Tensor) -> torch.Tensor:
    """fp32 [R, 32*G] -> DPOT_FMT_HL16G32 halves [R, 64*G] (test helper, plain torch arithmetic)."""
    R, N = y.shape
    hi = y.half()
    lo = ((y - hi.float()) * 2048.0).half()
    return torch.cat([hi.reshape(R, N // 32, 32), lo.reshape(R, N // 32, 32)], dim=2).reshape(R, 2 * N).contiguous()


def from_g32(g: torch.Tensor) -> torch.Tensor:
    R, N2 = g.shape
    r = g.reshape(R, N2 // 64, 64)
    return (r[:, :, :32].float() + r[:, :, 32:].float() / 2048.0).reshape(R, N2 // 2)


@_on_device
def out_tail_tc(Y1g: torch.Tensor, w2, b2, w4, b4, B: int, h: int, w: int, P: int, act, *, mu=None, sigma=None, Co: int = 0,
                ring=None, pred=None, slot0: int = 0, step: int = 0):
    """dpot_out_tail_tc: Y1g = per-pixel [hi 32 | lo 32] records (DPOT_FMT_HL16G32) -> out[B, h*P, w*P, nout], or into the
    ring window / prediction tensor when `ring` is given."""
    _need_cuda(w2, b2, w4, b4, mu, sigma, ring, pred)
    assert Y1g.is_cuda and Y1g.dtype == torch.float16 and Y1g.is_contiguous()
    nout, old = w4.shape[0], w4.shape[1]
    out = None if ring is not None else torch.empty((B, h * P, w * P, nout), device=Y1g.device, dtype=torch.float32)
    T = ring.shape[-2] if ring is not None else 0
    Ttot = pred.shape[-2] if pred is not None else 0
    check(_lib.load().dpot_out_tail_tc(ptr(Y1g), ptr(w2), ptr(b2), ptr(w4), ptr(b4), B, h, w, P, old, nout, act_id(act),
                                       ptr(mu), ptr(sigma), Co or nout, ptr(out), ptr(ring), ptr(pred), T, slot0, Ttot, step,
                                       _stream()), "dpot_out_tail_tc")
    return out


@_on_device
def gn_stats(x: torch.Tensor, B: int, n: int, groups: int = GROUPS) -> torch.Tensor:
    _need_cuda(x)
    E = x.shape[-1]
    stats = torch.empty((B, groups, 2), device=x.device, dtype=torch.float64)
    check(_lib.load().dpot_gn_stats(ptr(x), B, n, E, groups, ptr(stats), _stream()), "dpot_gn_stats")
    return stats


@_on_device
def gn_finalize(stats: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, n: int, eps: float = 1e-5):
    B, groups, _ = stats.shape
    E = gamma.numel()
    scale = torch.empty((B, E), device=stats.device, dtype=torch.float32)
    shift = torch.empty_like(scale)
    check(_lib.load().dpot_gn_finalize(ptr(stats), ptr(gamma), ptr(beta), B, n, E, groups, eps, ptr(scale), ptr(shift),
                                       _stream()), "dpot_gn_finalize")
    return scale, shift


@_on_device
def afno_fft_fwd(a, scale, shift, B, h, nb, km1, km2):
    _need_cuda(a, scale, shift)
    E = a.shape[-1]
    S = torch.empty((B * km1 * km2, 2 * E), device=a.device, dtype=torch.float32)
    check(_lib.load().dpot_afno_fft_fwd(ptr(a), ptr(scale), ptr(shift), B, h, E, nb, km1, km2, ptr(S), 1.0, _stream()),
          "dpot_afno_fft_fwd")
    return S


@_on_device
def afno_fft_inv(O2, a, scale, shift, B, h, nb, km1, km2, want_stats=True, groups: int = GROUPS):
    _need_cuda(O2, a, scale, shift)
    E = a.shape[-1]
    f = torch.empty_like(a)
    stats = torch.zeros((B, groups, 2), device=a.device, dtype=torch.float64) if want_stats else None
    check(_lib.load().dpot_afno_fft_inv(ptr(O2), ptr(a), ptr(scale), ptr(shift), B, h, E, nb, km1, km2, ptr(f),
                                        ptr(stats), groups, 1.0, _stream()), "dpot_afno_fft_inv")
    return f, stats


@_on_device
def afno_fused(lat: torch.Tensor, stats1: torch.Tensor, gamma1, beta1, w1, b1, w2, b2, B: int, h: int, act="gelu",
               eps: float = 1e-5, groups: int = GROUPS, want_stats: bool = True, debug: bool = False, gn2=None):
    """dpot_afno_fused: the whole AFNO2D mixer of one block (GroupNorm-1 by reference -> rfft2 -> complex block MLP ->
    irfft2 -> + skip) in one kernel.  lat[B*h*h, E]; stats1[B, groups, 2] (double); w1/w2 (2, nb, bs, bs), b1/b2 (2, nb, bs).
    Returns (f, stats2[, dbg]) -- dbg = per-unit images of the X and O1 operand tiles (test hook)."""
    _need_cuda(lat, gamma1, beta1, w1, b1, w2, b2)
    lib = _lib.load()
    E = lat.shape[1]
    nb, bs = w1.shape[1], w1.shape[2]
    if not lib.dpot_afno_fused_supported(h, E, nb, h, h // 2 + 1, groups):
        raise RuntimeError(f"dpot_afno_fused: geometry h={h} E={E} nb={nb} is not served by the fused mixer")
    packed = torch.empty(lib.dpot_afno_fused_packed_floats(nb), device=lat.device, dtype=torch.float32)
    check(lib.dpot_afno_fused_pack(ptr(w1.contiguous()), ptr(b1.contiguous()), ptr(w2.contiguous()), ptr(b2.contiguous()),
                                   nb, bs, ptr(packed), _stream()), "dpot_afno_fused_pack")
    f = torch.empty_like(lat)
    stats2 = torch.zeros((B, groups, 2), device=lat.device, dtype=torch.float64) if want_stats else None
    dbg = torch.zeros((B * nb, 2 * 147456 // 4), device=lat.device, dtype=torch.float32) if debug else None
    if gn2 is not None:      # (gamma2, beta2): GroupNorm-2 applied in the kernel -> also returns n2 as split fp16 [M, 2E]
        n2 = torch.empty((lat.shape[0], 2 * E), device=lat.device, dtype=torch.float16)
        check(lib.dpot_afno_fused_gn2(ptr(lat), ptr(stats1), ptr(gamma1), ptr(beta1), groups, eps, B, h, E, nb, ptr(packed),
                                      act_id(act), ptr(f), ptr(stats2), ptr(dbg), ptr(n2), ptr(gn2[0]), ptr(gn2[1]), eps,
                                      _stream()), "dpot_afno_fused_gn2")
        return f, stats2, n2
    check(lib.dpot_afno_fused(ptr(lat), ptr(stats1), ptr(gamma1), ptr(beta1), groups, eps, B, h, E, nb, ptr(packed),
                              act_id(act), ptr(f), ptr(stats2), ptr(dbg), _stream()), "dpot_afno_fused")
    return (f, stats2, dbg) if debug else (f, stats2)


@_on_device
def pack_afno(w: torch.Tensor, b: torch.Tensor):
    _need_cuda(w, b)
    _, nb, bs, _ = w.shape
    Wc = torch.empty((nb, 2 * bs, 2 * bs), device=w.device, dtype=torch.float32)
    bc = torch.empty((nb, 2 * bs), device=w.device, dtype=torch.float32)
    check(_lib.load().dpot_pack_afno(ptr(w.contiguous()), ptr(b.contiguous()), nb, bs, ptr(Wc), ptr(bc), _stream()),
          "dpot_pack_afno")
    return Wc, bc


@_on_device
def window_advance(xx, im, xx_next, pred=None, step=0):
    _need_cuda(xx, im, xx_next, pred)
    B, X, Y, T, Cc = xx.shape
    Tb = im.shape[-2]
    Ttot = pred.shape[-2] if pred is not None else 0
    check(_lib.load().dpot_window_advance(ptr(xx), ptr(im), ptr(xx_next), ptr(pred), B * X * Y, T, Tb, Cc, Ttot, step,
                                          _stream()), "dpot_window_advance")
    return xx_next


@_on_device
def ring_insert(im, ring, pred=None, slot0=0, step=0):
    """ring[..., (slot0+j) % T, :] = im[..., j, :]; pred[..., step*Tb+j, :] = im[..., j, :]."""
    _need_cuda(im, ring, pred)
    B, X, Y, T, Cc = ring.shape
    Tb = im.shape[-2]
    Ttot = pred.shape[-2] if pred is not None else 0
    check(_lib.load().dpot_ring_insert(ptr(im), ptr(ring), ptr(pred), B * X * Y, T, Tb, Cc, Ttot, slot0, step, _stream()),
          "dpot_ring_insert")
    return ring


@_on_device
def grad_sqnorm(grads, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Sum of squares over a list of gradient tensors -> one device double (no host sync): the norm half of
    torch.nn.utils.clip_grad_norm_ (train_temporal.py:228)."""
    grads = [g for g in grads if g is not None]
    _need_cuda(*grads)
    n = len(grads)
    if out is None:
        out = torch.zeros(1, device=grads[0].device if n else "cuda", dtype=torch.float64)
    if n == 0:
        return out.zero_()
    ga = (C.c_void_p * n)(*[g.data_ptr() for g in grads])
    na = (C.c_int64 * n)(*[g.numel() for g in grads])
    check(_lib.load().dpot_grad_sqnorm(ga, na, n, ptr(out), _stream()), "dpot_grad_sqnorm")
    return out


@_on_device
def noise_inject(x: torch.Tensor, scale: float, seed: int, offset: int, out: Optional[torch.Tensor] = None):
    """out = x + scale * ||x||_{(X,Y,T)} * randn  (train_temporal.py:205); returns (out, sumsq[B, C] double)."""
    _need_cuda(x)
    assert x.dim() == 5 and x.is_contiguous()
    B, X, Y, T, Cc = x.shape
    if out is None:
        out = torch.empty_like(x)
    sumsq = torch.empty((B, Cc), device=x.device, dtype=torch.float64)
    check(_lib.load().dpot_noise_inject(ptr(x), B, X * Y * T, Cc, float(scale), seed, offset, ptr(sumsq), ptr(out), _stream()),
          "dpot_noise_inject")
    return out, sumsq


@_on_device
def noise_inject_bwd(x, dy, scale, seed, offset, sumsq):
    B, X, Y, T, Cc = x.shape
    dx = torch.empty_like(x)
    dot = torch.empty((B, Cc), device=x.device, dtype=torch.float64)
    check(_lib.load().dpot_noise_inject_bwd(ptr(x), ptr(dy.contiguous()), B, X * Y * T, Cc, float(scale), seed, offset,
                                            ptr(sumsq), ptr(dot), ptr(dx), _stream()), "dpot_noise_inject_bwd")
    return dx


@_on_device
def lp_loss(x: torch.Tensor, y: torch.Tensor, mask: Optional[torch.Tensor] = None, loss: Optional[torch.Tensor] = None,
            accumulate: bool = False):
    """SimpleLpLoss(size_average=False)(x, y, mask) (utils/criterion.py:38-59) -> (loss[1] on the device, coef[B, C])."""
    _need_cuda(x, y, mask)
    assert x.shape == y.shape and x.dim() == 5 and x.is_contiguous() and y.is_contiguous()
    B, X, Y, T, Cc = x.shape
    if mask is not None:
        assert tuple(mask.shape) == (B, X, Y, 1, Cc) and mask.is_contiguous()
    partial = torch.empty((B, Cc, 3), device=x.device, dtype=torch.float64)
    coef = torch.empty((B, Cc), device=x.device, dtype=torch.float32)
    if loss is None:
        loss = torch.zeros(1, device=x.device, dtype=torch.float32)
    check(_lib.load().dpot_lp_loss(ptr(x), ptr(y), ptr(mask), B, X * Y, T, Cc, ptr(partial), ptr(coef), ptr(loss),
                                   1 if accumulate else 0, _stream()), "dpot_lp_loss")
    return loss, coef


@_on_device
def lp_loss_bwd(x, y, mask, coef, gscale: Optional[torch.Tensor] = None):
    B, X, Y, T, Cc = x.shape
    dx = torch.empty_like(x)
    check(_lib.load().dpot_lp_loss_bwd(ptr(x), ptr(y), ptr(mask), ptr(coef), ptr(gscale), B, X * Y, T, Cc, ptr(dx), _stream()),
          "dpot_lp_loss_bwd")
    return dx


@_on_device
def adam_step_multi(params, grads, ms, vs, vmaxs, steps, *, lr, beta1, beta2, eps, weight_decay, decoupled,
                    grad_scale=1.0, grad_sqnorm=None, max_norm=0.0):
    n = len(params)
    if n == 0:
        return
    _need_cuda(*params, *grads, *ms, *vs)
    VP = C.c_void_p * n
    pa = VP(*[p.data_ptr() for p in params])
    ga = VP(*[g.data_ptr() for g in grads])
    ma = VP(*[m.data_ptr() for m in ms])
    va = VP(*[v.data_ptr() for v in vs])
    xa = VP(*[x.data_ptr() for x in vmaxs]) if vmaxs else None
    na = (C.c_int64 * n)(*[p.numel() for p in params])
    sa = (C.c_int32 * n)(*steps)
    check(_lib.load().dpot_adam_step_multi_clip(pa, ga, ma, va, xa, na, n, lr, beta1, beta2, eps, weight_decay, sa,
                                                1 if decoupled else 0, grad_scale, ptr(grad_sqnorm), float(max_norm),
                                                _stream()), "dpot_adam_step_multi")


@_on_device
def lamb_step_multi(params, grads, ms, vs, steps, *, lr, beta1, beta2, eps, weight_decay, clamp_value, debias, adam,
                    norms: torch.Tensor, info: torch.Tensor):
    """Lamb.step (utils/optimizer.py:421-499) over a list of tensors: dpot_lamb_step_multi.  norms [2n] double and
    info [3n] float (weight_norm, adam_norm, trust_ratio per tensor) are caller-owned device buffers."""
    n = len(params)
    if n == 0:
        return
    _need_cuda(*params, *grads, *ms, *vs, info)
    assert norms.is_cuda and norms.dtype == torch.float64 and norms.numel() >= 2 * n and info.numel() >= 3 * n
    VP = C.c_void_p * n
    pa = VP(*[p.data_ptr() for p in params])
    ga = VP(*[g.data_ptr() for g in grads])
    ma = VP(*[m.data_ptr() for m in ms])
    va = VP(*[v.data_ptr() for v in vs])
    na = (C.c_int64 * n)(*[p.numel() for p in params])
    sa = (C.c_int32 * n)(*steps)
    check(_lib.load().dpot_lamb_step_multi(pa, ga, ma, va, na, n, lr, beta1, beta2, eps, weight_decay, float(clamp_value), sa,
                                           1 if debias else 0, 1 if adam else 0, ptr(norms), ptr(info), _stream()),
          "dpot_lamb_step_multi")
