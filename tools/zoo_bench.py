"""Every published DPOT size at FULL depth on one GPU (BASELINE.json configs 1-5): inference forward time and one
training step (forward + SimpleLpLoss + backward + clip + Adam) through the drop-in API, with the path that served it.

    python tools/zoo_bench.py [Ti S M L H] -> one JSON line per model

L / H use 256^2 / 128^2 fields as BASELINE.json configs 4 / 5 name them (L: patch 16, modes 64)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpot_b200 import zoo
from dpot_b200.models.dpot import DPOTNet
from dpot_b200.train import ar_train_step
from dpot_b200.utils.optimizer import Adam

SHAPES = {"Ti": dict(img_size=128, patch_size=8), "S": dict(img_size=128, patch_size=8), "M": dict(img_size=128, patch_size=8),
          "L": dict(img_size=256, patch_size=16, modes=64), "H": dict(img_size=128, patch_size=8)}
BATCH = {"Ti": (32, 16), "S": (32, 16), "M": (32, 16), "L": (8, 4), "H": (8, 4)}      # (inference, training)
dev = torch.device("cuda")


def timed(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for name in (sys.argv[1:] or ["Ti", "S", "M", "L", "H"]):
    cfg = zoo.zoo_cfg(name, **SHAPES[name])
    R = cfg["img_size"]
    Bi, Bt = BATCH[name]
    out = {"model": name, "img": R, "depth": cfg["depth"], "embed_dim": cfg["embed_dim"]}
    try:
        m = zoo.synthetic_weights_(DPOTNet(**cfg), seed=0).to(dev)
        out["params_M"] = sum(p.numel() for p in m.parameters()) / 1e6
        x = torch.randn(Bi, R, R, 10, 4, device=dev)
        m.eval()
        with torch.no_grad():
            for _ in range(2):
                y, _ = m(x)
            ms = timed(lambda: m(x), 5)
        out["inference"] = {"batch": Bi, "ms_per_forward": ms, "field_steps_per_s": Bi / ms * 1e3, "finite": bool(torch.isfinite(y).all())}
        m.train()
        opt = Adam(m.parameters(), lr=1e-4, betas=(0.9, 0.9), weight_decay=1e-6)
        xt = torch.randn(Bt, R, R, 10, 4, device=dev)
        yt = torch.randn(Bt, R, R, 1, 4, device=dev)
        msk = torch.ones(Bt, R, R, 1, 4, device=dev)
        it = [0]

        def step():
            it[0] += 1
            return ar_train_step(m, opt, xt, yt, msk, grad_clip=1e4, step=it[0])
        for _ in range(2):
            loss = step()
        ms = timed(step, 4)
        out["training"] = {"batch": Bt, "ms_per_step": ms, "field_steps_per_s": Bt / ms * 1e3, "loss": float(loss),
                           "path": "one-call step (dpot_train_*)" if (m._train_eng is not None and m._train_eng.supported) else
                           "per-operator path (autograd.py)"}
        out["peak_mem_GB"] = torch.cuda.max_memory_allocated() / 2 ** 30
    except Exception as e:      # noqa: BLE001  (report, keep going with the next size)
        out["error"] = f"{type(e).__name__}: {e}"[:300]
    print(json.dumps(out), flush=True)
    del m
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
