#!/usr/bin/env python
"""Where the optimizer step's time goes (DPOT-S parameter set): gradient norm, Adam kernel, the pair as one CUDA graph."""
import sys
import torch
sys.path.insert(0, ".")
from dpot_b200 import ops, zoo
from dpot_b200.models.dpot import DPOTNet
from dpot_b200.utils.clip import clip_grad_norm_
from dpot_b200.utils.optimizer import Adam

name = sys.argv[1] if len(sys.argv) > 1 else "S"
m = zoo.synthetic_weights_(DPOTNet(**zoo.zoo_cfg(name)), seed=0).cuda()
ps = [p for p in m.parameters()]
for p in ps:
    p.grad = torch.randn_like(p) * 1e-3
opt = Adam(ps, lr=1e-4, betas=(0.9, 0.9), weight_decay=1e-6)
n = sum(p.numel() for p in ps)


def timed(fn, it=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3


grads = [p.grad for p in ps]
print(f"{name}: {n / 1e6:.1f} M parameters in {len(ps)} tensors")
t = timed(lambda: ops.grad_sqnorm(grads)); print(f"grad_sqnorm eager      {t:7.1f} us  {4 * n / t / 1e3:7.0f} GB/s")
t = timed(lambda: opt.step()); print(f"Adam.step eager        {t:7.1f} us  {28 * n / t / 1e3:7.0f} GB/s (28 B/param)")
t = timed(lambda: (clip_grad_norm_(ps, 1e4, optimizer=opt), opt.step())); print(f"clip + Adam eager      {t:7.1f} us")
for what in ("adam", "clip+adam", "sqnorm"):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            if what == "sqnorm":
                ops.grad_sqnorm(grads)
            else:
                if what == "clip+adam":
                    clip_grad_norm_(ps, 1e4, optimizer=opt)
                opt.step()
    torch.cuda.current_stream().wait_stream(side)
    t = timed(g.replay)
    print(f"{what:10s} graph replay {t:7.1f} us  {(28 if 'adam' in what else 4) * n / t / 1e3:7.0f} GB/s")
