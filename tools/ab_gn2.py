#!/usr/bin/env python
"""A/B of the rollout (DPOT-S, B = 32, 10 AR steps, CUDA graph) with GroupNorm-2 inside the fused AFNO mixer on / off."""
import sys
import torch
sys.path.insert(0, ".")
from dpot_b200 import _lib, zoo
from dpot_b200.models.dpot import DPOTNet
from dpot_b200.rollout import RolloutEngine

lib = _lib.load()
m = zoo.synthetic_weights_(DPOTNet(**zoo.zoo_cfg("S")), seed=0).cuda().eval()
xs = [torch.randn(32, 128, 128, 10, 4, device="cuda") for _ in range(4)]
res = {}
for rep in range(2):
    for on in (0, 1):
        lib.dpot_afno_set_fused_gn2(on)
        eng = RolloutEngine(m, 32, 10, use_graph=True)
        for i in range(3):
            out = eng.run(xs[i % 4])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(10):
            out = eng.run(xs[i % 4])
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        res.setdefault(on, []).append((ms, out.float().norm().item()))
        print(f"fused_gn2={on}: {ms:.3f} ms / rollout = {320 / ms * 1e3:.0f} field-steps/s, |pred| = {res[on][-1][1]:.6f}", flush=True)
lib.dpot_afno_set_fused_gn2(1)
