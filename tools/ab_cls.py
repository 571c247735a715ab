#!/usr/bin/env python
"""A/B of the rollout (DPOT-S, B = 32, 10 AR steps, CUDA graph, cls head ON) with the classification head on the f16-split
tensor-core engine (1) or on the exact-fp32 CUDA-core skinny kernels (0); and without the head for reference."""
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
    for tag, eng_id, want in (("no cls head", 1, False), ("cls on tcgen05", 1, True), ("cls on CUDA cores", 0, True)):
        lib.dpot_set_cls_engine(eng_id)
        eng = RolloutEngine(m, 32, 10, use_graph=True, want_cls=want)
        for i in range(3):
            out = eng.run(xs[i % 4])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(10):
            out = eng.run(xs[i % 4])
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        cls = getattr(eng, "cls", None)
        print(f"{tag:18s}: {ms:.3f} ms / rollout = {320 / ms * 1e3:.0f} field-steps/s" +
              (f", |cls| = {cls.float().norm().item():.6f}" if want and cls is not None else ""), flush=True)
