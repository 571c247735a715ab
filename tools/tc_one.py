"""Run one GEMM shape on the tcgen05 engine a few times (target for ncu)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpot_b200 import ops
M, N, K = [int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (8192, 1024, 1024))]
eng = int(sys.argv[4]) if len(sys.argv) > 4 else 2
A = torch.randn((M, K), device="cuda"); W = torch.randn((N, K), device="cuda") / 32
b = torch.randn(N, device="cuda"); C = torch.empty((M, N), device="cuda")
for _ in range(4):
    ops.gemm(A, W, bias=b, act="gelu", out=C, engine=eng)
torch.cuda.synchronize()
print("ok")
