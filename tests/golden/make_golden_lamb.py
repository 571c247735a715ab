#!/usr/bin/env python
"""Reference trajectories of Lamb (utils/optimizer.py:359-499) -> tests/golden/lamb.npz.

    python tests/golden/make_golden_lamb.py        # needs /root/reference (build container only)

Three parameters of ragged sizes (one all-zero: the trust_ratio = 1 branch, :480-481) stepped four times with changing
learning rates by the UNMODIFIED reference optimizer on CPU, for the constructor variants the scripts use
(evaluate.py:135-136: adam=True, debias=False, weight_decay=1e-4) and the general trust-ratio / debias / clamp paths."""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, os.environ.get("DPOT_REFERENCE", "/root/reference"))
from utils.optimizer import Lamb  # noqa: E402  (reference)

VARIANTS = {
    "script": dict(betas=(0.9, 0.9), adam=True, debias=False, weight_decay=1e-4),          # evaluate.py:136
    "trust": dict(betas=(0.9, 0.999), adam=False, debias=False, weight_decay=1e-2),
    "debias": dict(betas=(0.9, 0.99), adam=False, debias=True, weight_decay=0.0),
    "clamp": dict(betas=(0.9, 0.999), adam=False, debias=True, weight_decay=1e-3, clamp_value=0.5),
}


def main():
    rng = np.random.default_rng(5)
    sizes, nsteps = [1031, 257, 64], 4
    p0 = [rng.standard_normal(n).astype(np.float32) for n in sizes]
    p0[2][:] = 0.0                                              # weight_norm == 0 on the first step
    grads = [[rng.standard_normal(n).astype(np.float32) for n in sizes] for _ in range(nsteps)]
    lrs = [1e-3, 5e-4, 2e-3, 1e-3]
    out = {"lrs": np.array(lrs), "kw": json.dumps({k: {a: (list(b) if isinstance(b, tuple) else b) for a, b in v.items()}
                                                   for k, v in VARIANTS.items()})}
    for i in range(len(sizes)):
        out[f"p0.{i}"] = p0[i]
        out[f"grads.{i}"] = np.stack([grads[s][i] for s in range(nsteps)])
    for tag, kw in VARIANTS.items():
        ps = [torch.nn.Parameter(torch.from_numpy(a.copy())) for a in p0]
        opt = Lamb(ps, lr=1e-3, eps=1e-6, **kw)
        traj = [[] for _ in sizes]
        info = [[] for _ in sizes]
        for s in range(nsteps):
            opt.param_groups[0]["lr"] = lrs[s]
            for i, p in enumerate(ps):
                p.grad = torch.from_numpy(grads[s][i].copy())
            opt.step()
            for i, p in enumerate(ps):
                st = opt.state[p]
                traj[i].append(p.detach().numpy().copy())
                info[i].append([float(st["weight_norm"]), float(st["adam_norm"]), float(st["trust_ratio"])])
        for i, p in enumerate(ps):
            out[f"{tag}.p.{i}"] = np.stack(traj[i])
            out[f"{tag}.info.{i}"] = np.array(info[i], np.float64)
            out[f"{tag}.m.{i}"] = opt.state[p]["exp_avg"].numpy()
            out[f"{tag}.v.{i}"] = opt.state[p]["exp_avg_sq"].numpy()
    np.savez_compressed(os.path.join(HERE, "lamb.npz"), **out)
    print("lamb ok")


if __name__ == "__main__":
    main()
