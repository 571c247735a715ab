"""Training-step timing through the drop-in API (train_temporal.py:201-230, T_ar = 1, noise off): forward + SimpleLpLoss +
backward + clip + Adam.step on one GPU, CUDA events, synthetic data.

    python tools/train_step_bench.py [S|M|Ti] [batch] [path: auto|generic|both] [steps] [fp32|half]

Prints one JSON line: ms per step for the one-call training step (dpot_train_*) and/or the per-operator path."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpot_b200 import _lib, zoo
from dpot_b200.models.dpot import DPOTNet
from dpot_b200.train import ar_train_step
from dpot_b200.utils.optimizer import Adam

name = sys.argv[1] if len(sys.argv) > 1 else "S"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
which = sys.argv[3] if len(sys.argv) > 3 else "both"
n = int(sys.argv[4]) if len(sys.argv) > 4 else 10
precision = sys.argv[5] if len(sys.argv) > 5 else "fp32"
import dpot_b200
dpot_b200.set_precision(precision)
cfg = zoo.zoo_cfg(name)
dev = torch.device("cuda")
x = torch.randn(B, 128, 128, 10, 4, device=dev)
y = torch.randn(B, 128, 128, 1, 4, device=dev)
msk = torch.ones(B, 128, 128, 1, 4, device=dev)
out = {"model": name, "batch": B, "steps": n, "precision": precision}
for path in (["auto", "generic"] if which == "both" else [which]):
    m = zoo.synthetic_weights_(DPOTNet(**cfg), seed=0).to(dev).train()
    m.train_path = path
    opt = Adam(m.parameters(), lr=1e-4, betas=(0.9, 0.9), weight_decay=1e-6)
    it = {"i": 0}

    def step():
        it["i"] += 1
        return ar_train_step(m, opt, x, y, msk, T_bundle=1, noise_scale=0.0, grad_clip=1e4, step=it["i"])

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    l0 = _lib.load().dpot_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    out[path] = {"ms_per_step": ms, "field_steps_per_s": B / ms * 1e3, "loss": float(loss),
                 "launches_per_step": (_lib.load().dpot_launch_count() - l0) / n,
                 "fused": bool(m._train_eng is not None and m._train_eng.supported)}
    del m, opt
    torch.cuda.empty_cache()
print(json.dumps(out))
