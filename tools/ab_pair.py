#!/usr/bin/env python
"""A/B of the rollout (DPOT-S, B = 32, 10 AR steps, CUDA graph, cls head on) with the f16-split engine's tile plan:
-1 = cost model, 1 = CTA pairs wherever legal, 0 = single-CTA tiles only."""
import sys
import torch
sys.path.insert(0, ".")
from dpot_b200 import _lib, zoo
from dpot_b200.models.dpot import DPOTNet
from dpot_b200.rollout import RolloutEngine

lib = _lib.load()
m = zoo.synthetic_weights_(DPOTNet(**zoo.zoo_cfg("S")), seed=0).cuda().eval()
xs = [torch.randn(32, 128, 128, 10, 4, device="cuda") for _ in range(4)]
for rep in range(2):
    for mode in (-1, 1, 0):
        lib.dpot_tc16_set_pair(mode)
        eng = RolloutEngine(m, 32, 10, use_graph=True, want_cls=True)
        for i in range(3):
            out = eng.run(xs[i % 4])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(10):
            out = eng.run(xs[i % 4])
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"pair mode {mode:2d}: {ms:.3f} ms / rollout = {320 / ms * 1e3:.0f} field-steps/s", flush=True)
lib.dpot_tc16_set_pair(-1)
