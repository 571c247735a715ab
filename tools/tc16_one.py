"""Run one GEMM shape on the f16-split tcgen05 engine a few times (target for ncu)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpot_b200 import ops
M, N, K = [int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (8192, 1024, 1024))]
act = sys.argv[4] if len(sys.argv) > 4 else "gelu"
out16 = (sys.argv[5] == "1") if len(sys.argv) > 5 else True
A = ops.split_f16(torch.randn((M, K), device="cuda")); W = ops.split_f16(torch.randn((N, K), device="cuda") / 32)
b = torch.randn(N, device="cuda")
for _ in range(4):
    ops.gemm16(A, W, bias=b, act=None if act == "none" else act, out16=out16)
torch.cuda.synchronize()
print("ok")
