"""Phase timeline of the fused AFNO mixer kernel (CTA 0, thread 0; dpot_afno_fused_set_trace) at the BASELINE shape
(B=32, E=1024, nb=8 -> 256 units on 148 SMs), and its CUDA-event time over rotated inputs (> L2)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpot_b200 import _lib, ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
nb, E = 8, 1024
lib = _lib.load()
g = torch.Generator(device="cuda").manual_seed(0)
lats = [torch.randn((B * 256, E), device="cuda", generator=g) for _ in range(6)]
w1 = torch.randn((2, nb, 128, 128), device="cuda", generator=g) / 11.3
w2 = torch.randn((2, nb, 128, 128), device="cuda", generator=g) / 11.3
b1 = 0.1 * torch.randn((2, nb, 128), device="cuda", generator=g)
b2 = 0.1 * torch.randn((2, nb, 128), device="cuda", generator=g)
gamma, beta = torch.ones(E, device="cuda"), torch.zeros(E, device="cuda")
packed = torch.empty(lib.dpot_afno_fused_packed_floats(nb), device="cuda")
_lib.check(lib.dpot_afno_fused_pack(w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), nb, 128, packed.data_ptr(), None))
stats = [ops.gn_stats(l, B, 256) for l in lats]
f = torch.empty_like(lats[0])
st2 = torch.zeros((B, 8, 2), device="cuda", dtype=torch.float64)
st = torch.cuda.current_stream().cuda_stream
n2 = torch.empty((B * 256, 2 * E), device="cuda", dtype=torch.float16)
GN2 = False
def run(i):
    if GN2:    # GroupNorm-2 inside the kernel, f stays on chip (the inference forward's default)
        _lib.check(lib.dpot_afno_fused_gn2(lats[i].data_ptr(), stats[i].data_ptr(), gamma.data_ptr(), beta.data_ptr(), 8, 1e-5, B, 16,
                                           E, nb, packed.data_ptr(), 0, None, None, None, n2.data_ptr(), gamma.data_ptr(),
                                           beta.data_ptr(), 1e-5, st))
    else:
        _lib.check(lib.dpot_afno_fused(lats[i].data_ptr(), stats[i].data_ptr(), gamma.data_ptr(), beta.data_ptr(), 8, 1e-5, B, 16, E, nb,
                                       packed.data_ptr(), 0, f.data_ptr(), st2.data_ptr(), None, st))
for GN2 in (False, True):
    for i in range(6):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(30):
        run(i % 6)
    e1.record()
    torch.cuda.synchronize()
    print(f"fused AFNO mixer B={B}, GroupNorm-2 {'inside (f on chip)' if GN2 else 'outside (f + statistics written)'}: "
          f"{e0.elapsed_time(e1) / 30 * 1e3:.1f} us per launch")
    trace = torch.zeros(8 * 8, device="cuda", dtype=torch.int64)
    lib.dpot_afno_fused_set_trace(trace.data_ptr())
    run(0)
    torch.cuda.synchronize()
    lib.dpot_afno_fused_set_trace(None)
    t = trace.cpu().numpy().reshape(8, 8)
    names = ["A: GN1+rfft2 -> X", "wait layer-1 MMAs", "E1: act -> O1", "wait layer-2 MMAs", "E2a: col inverse",
             "E2b: rows+skip+store" + (" + GN2" if GN2 else "")]
    for u in range(2):
        if t[u, 0] == 0:
            continue
        d = np.diff(t[u, :7])
        print(f"unit {u} of CTA 0: total {t[u, 6] - t[u, 0]} clk")
        for n, v in zip(names, d):
            print(f"   {n:32s} {v:7d} clk")
