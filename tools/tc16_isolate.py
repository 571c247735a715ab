"""Pipeline-isolation timings of the f16-split engine: full kernel vs loads-only / MMA-only / no-epilogue variants."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpot_b200 import ops, _lib
lib = _lib.load()
def bench(M, N, K, act, out16, nb=1, nrot=6, reps=20, rowbias=None, stats=None, residual=None):
    As = [ops.split_f16(torch.randn((M, nb * K), device="cuda")) for _ in range(nrot)]
    Ws = [ops.split_f16(torch.randn((nb * N, K), device="cuda") / 32).reshape(nb, N, 2 * K) if nb > 1 else
          ops.split_f16(torch.randn((N, K), device="cuda") / 32) for _ in range(nrot)]
    bias = torch.randn(nb * N, device="cuda").reshape(nb, N) if nb > 1 else torch.randn(N, device="cuda")
    res = []
    for pair in (0, -1):
        lib.dpot_tc16_set_pair(pair)
        for dbg in (0, 1, 2, 4, 5, 6, 3):
            lib.dpot_tc16_set_debug(dbg)
            for i in range(3):
                ops.gemm16(As[i], Ws[i], bias=bias, act=act, out16=out16, nb=nb, rowbias=rowbias, stats=stats, residual=residual)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(reps):
                ops.gemm16(As[i % nrot], Ws[i % nrot], bias=bias, act=act, out16=out16, nb=nb, rowbias=rowbias, stats=stats, residual=residual)
            e1.record(); torch.cuda.synchronize()
            res.append((pair, dbg, e0.elapsed_time(e1) * 1e3 / reps))
    lib.dpot_tc16_set_debug(0); lib.dpot_tc16_set_pair(-1)
    print(f"M={M} N={N} K={K} nb={nb} act={act} out16={out16} rowbias={rowbias is not None} stats={stats} residual={residual is not None}")
    names = {0: "full", 1: "noTMA", 2: "noMMA", 4: "noEPI", 5: "MMA only", 6: "TMA only", 3: "EPI only"}
    for pair in (0, -1):
        print(f"  pair={pair:2d}: " + "  ".join(f"{names[d]}={t:6.1f}" for pp, d, t in res if pp == pair), flush=True)
if len(sys.argv) > 1 and sys.argv[1] == "tagg":
    # the folded time-aggregation contraction of DPOT-S (K = 352, row-periodic bias, GroupNorm statistics) and fc2
    rb = torch.randn((256, 1024), device="cuda")
    bench(8192, 1024, 352, None, False, rowbias=rb, stats=(8, 256))
    bench(8192, 1024, 352, None, False, rowbias=rb)
    bench(8192, 1024, 352, None, False)
    bench(8192, 1024, 1024, None, False, residual=torch.randn((8192, 1024), device="cuda"), stats=(8, 256))
    sys.exit(0)
bench(8192, 2048, 1024, "gelu", False)
bench(8192, 1024, 1024, "gelu", True)
bench(8192, 1024, 1024, None, False)
bench(4608, 256, 256, "gelu", True, nb=8)
bench(4608, 256, 256, None, False, nb=8)
