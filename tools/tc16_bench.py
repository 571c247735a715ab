"""Time the f16-split tcgen05 engine on the DPOT-S GEMM shapes (operands rotated through > L2)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpot_b200 import ops
def bench(M, N, K, act, out16, nb=1, nrot=6, reps=30, side=False):
    As = [ops.split_f16(torch.randn((M, nb * K), device="cuda")) for _ in range(nrot)]
    Ws = [ops.split_f16(torch.randn((nb * N, K), device="cuda") / 32).reshape(nb, N, 2 * K) if nb > 1 else
          ops.split_f16(torch.randn((N, K), device="cuda") / 32) for _ in range(nrot)]
    bias = torch.randn(nb * N, device="cuda").reshape(nb, N) if nb > 1 else torch.randn(N, device="cuda")
    kw = {}
    if side:   # the fc2 configuration: residual + fused GroupNorm statistics
        kw = dict(residual=torch.randn((M, N), device="cuda"), stats=(8, 256))
    for i in range(nrot):
        ops.gemm16(As[i], Ws[i], bias=bias, act=act, out16=out16, nb=nb, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        ops.gemm16(As[i % nrot], Ws[i % nrot], bias=bias, act=act, out16=out16, nb=nb, **kw)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3 / reps
    print(f"M={M} N={N} K={K} nb={nb} act={act} out16={out16} side={side}: {t*1e6:8.1f} us  {2.0*M*N*K*nb/t/1e12:7.1f} TFLOP/s", flush=True)
from dpot_b200 import _lib
for mode in ([int(a) for a in sys.argv[1:]] or [0, 1, -1]):
  _lib.load().dpot_tc16_set_pair(mode)
  print("pair mode", mode)
  bench(8192, 1024, 1024, "gelu", True)
  bench(8192, 1024, 1024, None, False)
  bench(8192, 1024, 1024, None, False, side=True)
  bench(8192, 1024, 352, None, False, side=True)
  bench(8192, 1024, 1024, None, True)
  bench(8192, 2048, 1024, "gelu", False)
  bench(8192, 1024, 352, None, False)
  bench(4608, 256, 256, "gelu", True, nb=8)
  bench(4608, 256, 256, None, False, nb=8)
  bench(8192, 4096, 1024, "gelu", True)
  bench(8192, 1024, 4096, None, False)
