"""GPU probe of the tcgen05 GEMM engine: correctness diagnostics + timing.  Run on a B200:
    python tools/tc_probe.py
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpot_b200 import _lib, ops  # noqa: E402


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def run(M, N, K, engine, seed=0, **kw):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    out = ops.gemm(torch.from_numpy(A).cuda(), torch.from_numpy(W).cuda(), engine=engine, **kw)
    torch.cuda.synchronize()
    return A, W, out.cpu().numpy()


def main():
    lib = _lib.load()
    print("tc_available:", lib.dpot_tc_available(), "device:", torch.cuda.get_device_name(0))
    # (a) identity weight: C must reproduce A exactly
    M, N, K = 256, 128, 128
    A = np.random.default_rng(0).standard_normal((M, K)).astype(np.float32)
    W = np.eye(N, K, dtype=np.float32)
    C = ops.gemm(torch.from_numpy(A).cuda(), torch.from_numpy(W).cuda(), engine=2).cpu().numpy()
    print(f"(a) identity  M{M} N{N} K{K}: rel={rel(C, A):.3e} max|d|={np.abs(C - A).max():.3e}")
    if rel(C, A) > 1e-5:
        print("   C[0,:8]  =", C[0, :8]); print("   A[0,:8]  =", A[0, :8])
        print("   C[:8,0]  =", C[:8, 0]); print("   A[:8,0]  =", A[:8, 0])
        # is C a permutation of A's columns?
        for j in range(4):
            hits = np.where(np.abs(A[0] - C[0, j]) < 1e-6)[0]
            print(f"   C[0,{j}] matches A[0,{hits.tolist()}]")
    # (b..) random problems
    for (M, N, K) in [(256, 128, 32), (256, 128, 64), (256, 128, 256), (300, 200, 96), (64, 32, 32), (144, 256, 64),
                      (1024, 1024, 1024), (8192, 1024, 1024), (4608, 256, 256)]:
        A, W, C = run(M, N, K, 2)
        ref = A.astype(np.float64) @ W.T.astype(np.float64)
        A1, W1, C1 = run(M, N, K, 1)
        print(f"(b) M{M} N{N} K{K}: tc rel={rel(C, ref):.3e}  simt rel={rel(C1, ref):.3e}")
    # (c) epilogue options
    rng = np.random.default_rng(5)
    M, N, K = 512, 256, 128
    A = rng.standard_normal((M, K)).astype(np.float32); W = (rng.standard_normal((N, K)) / 11).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32); R = rng.standard_normal((M, N)).astype(np.float32)
    rb = rng.standard_normal((64, N)).astype(np.float32)
    asc = rng.standard_normal((4, K)).astype(np.float32); ash = rng.standard_normal((4, K)).astype(np.float32)
    csc = rng.standard_normal((2, N)).astype(np.float32); csh = rng.standard_normal((2, N)).astype(np.float32)
    t = lambda x: torch.from_numpy(x).cuda()
    for eng in (1, 2):
        C = ops.gemm(t(A), t(W), bias=t(b), act="gelu", residual=t(R), rowbias=t(rb), a_scale=t(asc), a_shift=t(ash),
                     a_rows_per_sample=128, c_scale=t(csc), c_shift=t(csh), c_rows_per_sample=256, engine=eng).cpu().numpy()
        Ap = A.astype(np.float64) * np.repeat(asc, 128, 0) + np.repeat(ash, 128, 0)
        v = Ap @ W.T.astype(np.float64) + b + np.tile(rb, (8, 1))
        from math import sqrt
        from scipy.special import erf
        v = 0.5 * v * (1 + erf(v / sqrt(2)))
        v = v * np.repeat(csc, 256, 0) + np.repeat(csh, 256, 0) + R
        print(f"(c) full epilogue engine={eng}: rel={rel(C, v):.3e}")
    # (d) batched block-diagonal (AFNO shape)
    nb, bs2, Ms = 8, 256, 4608
    S = rng.standard_normal((Ms, nb * bs2)).astype(np.float32)
    Wc = (rng.standard_normal((nb, bs2, bs2)) / 16).astype(np.float32); bc = rng.standard_normal((nb, bs2)).astype(np.float32)
    for eng in (1, 2):
        O1 = ops.gemm_batched_cols(t(S), t(Wc), t(bc), nb, act=None, engine=eng).cpu().numpy()
        ref = np.concatenate([S[:, k * bs2:(k + 1) * bs2].astype(np.float64) @ Wc[k].T.astype(np.float64) + bc[k] for k in range(nb)], 1)
        print(f"(d) batched AFNO engine={eng}: rel={rel(O1, ref):.3e}")
    # (e) timing
    for (M, N, K) in [(8192, 1024, 1024), (8192, 4096, 1024), (8192, 1024, 4096), (8192, 1024, 352), (8192, 2048, 1024)]:
        nrot = 4
        As = [torch.randn((M, K), device="cuda") for _ in range(nrot)]
        Ws = [torch.randn((N, K), device="cuda") / 32 for _ in range(nrot)]
        Cs = [torch.empty((M, N), device="cuda") for _ in range(nrot)]
        bias = torch.randn(N, device="cuda")
        for eng in (1, 2):
            for i in range(nrot):
                ops.gemm(As[i], Ws[i], bias=bias, act="gelu", out=Cs[i], engine=eng)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            e0.record()
            for i in range(reps):
                ops.gemm(As[i % nrot], Ws[i % nrot], bias=bias, act="gelu", out=Cs[i % nrot], engine=eng)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            print(f"(e) M{M} N{N} K{K} engine={eng}: {ms*1e3:.1f} us  {2.0*M*N*K/ms/1e9:.1f} TFLOP/s")
    # (f) accumulation-chain length: accuracy and speed vs flush cadence
    M, N, K = 8192, 1024, 1024
    rngf = np.random.default_rng(9)
    Af = rngf.standard_normal((M, K)).astype(np.float32); Wf = (rngf.standard_normal((N, K)) / 32).astype(np.float32)
    reff = Af.astype(np.float64) @ Wf.T.astype(np.float64)
    At, Wtt = t(Af), t(Wf); Ct = torch.empty((M, N), device="cuda")
    for fl in (1, 2, 4, 8, 16, 32):
        lib.dpot_tc_set_flush(fl)
        ops.gemm(At, Wtt, out=Ct, engine=2); torch.cuda.synchronize()
        err = rel(Ct.cpu().numpy(), reff)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(10):
            ops.gemm(At, Wtt, out=Ct, engine=2)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"(f) flush={fl:2d}: rel={err:.3e}  {ms*1e3:.1f} us  {2.0*M*N*K/ms/1e9:.1f} TFLOP/s")
    lib.dpot_tc_set_flush(2)
    for tr in (0, 1):
        lib.dpot_tc_set_trunc(tr)
        ops.gemm(At, Wtt, out=Ct, engine=2); torch.cuda.synchronize()
        err = rel(Ct.cpu().numpy(), reff)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(10):
            ops.gemm(At, Wtt, out=Ct, engine=2)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"(g) trunc={tr}: rel={err:.3e}  {ms*1e3:.1f} us  {2.0*M*N*K/ms/1e9:.1f} TFLOP/s")
    lib.dpot_tc_set_trunc(0)
    # batched AFNO timing
    St = t(S); Wt = t(Wc); bt = t(bc)
    for eng in (1, 2):
        O = ops.gemm_batched_cols(St, Wt, bt, nb, act="gelu", engine=eng)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20):
            ops.gemm_batched_cols(St, Wt, bt, nb, act="gelu", out=O, engine=eng)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"(e) AFNO batched Ms{Ms} nb{nb} 256x256 engine={eng}: {ms*1e3:.1f} us  {2.0*Ms*nb*bs2*bs2/ms/1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
