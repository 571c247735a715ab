"""Pipeline timeline of CTA 0 of the tcgen05 GEMM engine (clock64 cycles)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpot_b200 import _lib, ops
lib = _lib.load()
M, N, K = [int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (8192, 1024, 1024))]
act = sys.argv[4] if len(sys.argv) > 4 else "gelu"
A = torch.randn((M, K), device="cuda"); W = torch.randn((N, K), device="cuda") / 32
b = torch.randn(N, device="cuda"); C = torch.empty((M, N), device="cuda")
for _ in range(3):
    ops.gemm(A, W, bias=b, act=act if act != "none" else None, out=C, engine=2)
buf = torch.zeros(7 * 64, dtype=torch.int64, device="cuda")
lib.dpot_tc_set_trace(buf.data_ptr())
ops.gemm(A, W, bias=b, act=act if act != "none" else None, out=C, engine=2)
torch.cuda.synchronize()
lib.dpot_tc_set_trace(None)
t = buf.cpu().numpy().reshape(7, 64)
t0 = t[0, 0]
names = ["prod_issue", "mma_start", "conv_start", "conv_done", "flush_start", "tile_acc_done", "tile_epi_done"]
np.set_printoptions(linewidth=200)
for r, nm in enumerate(names):
    row = t[r]; row = row[row > 0] - t0
    print(f"{nm:14s}", row[:40])
pi, ms, cs, cd = t[0] - t0, t[1] - t0, t[2] - t0, t[3] - t0
n = 24
print("per k-block: tma latency (conv_start - prod_issue):", (cs[:n] - pi[:n]))
print("per k-block: convert time (conv_done - conv_start):", (cd[:n] - cs[:n]))
print("per k-block: mma wake (mma_start - conv_done):     ", (ms[:n] - cd[:n]))
print("per k-block: mma_start deltas:                      ", np.diff(ms[:n + 1]))
