#!/usr/bin/env python
"""Accumulation-chain experiment for the f16-split engine: one long contraction (K = 4096 / 6144 / 8192, the fc2 of
DPOT-M / L / H) in one launch versus K-chunks chained through the residual input, against fp64; fp32 cuBLAS beside it."""
import sys
import torch
sys.path.insert(0, ".")
from dpot_b200 import ops

torch.backends.cuda.matmul.allow_tf32 = False
g = torch.Generator(device="cuda").manual_seed(0)
M, N = 2048, 2048
for K in (1024, 4096, 6144, 8192):
    A = torch.nn.functional.gelu(torch.randn((M, K), device="cuda", generator=g))
    W = (torch.rand((N, K), device="cuda", generator=g) * 2 - 1) / K ** 0.5
    R = torch.randn((M, N), device="cuda", generator=g)
    ref = A.double() @ W.double().t()
    nrm = ref.norm()
    out = {}
    out["cublas fp32"] = float(((A @ W.t()).double() - ref).norm() / nrm)
    A16, W16 = ops.split_f16(A), ops.split_f16(W)
    out["one launch"] = float((ops.gemm16(A16, W16).double() - ref).norm() / nrm)
    for ch in (2048, 1024, 512):
        if ch >= K:
            continue
        acc = None
        for k0 in range(0, K, ch):
            a16 = torch.cat([A16[:, k0:k0 + ch], A16[:, K + k0:K + k0 + ch]], 1).contiguous()
            w16 = torch.cat([W16[:, k0:k0 + ch], W16[:, K + k0:K + k0 + ch]], 1).contiguous()
            acc = ops.gemm16(a16, w16, residual=acc)
        out[f"chunks of {ch}"] = float((acc.double() - ref).norm() / nrm)
    print(f"K={K}: " + "  ".join(f"{k} {v:.2e}" for k, v in out.items()), flush=True)
