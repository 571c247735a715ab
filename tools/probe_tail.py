import json, os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from oracle import dpot_oracle as O
from test_train_step_gpu import _run, G
from test_parity_r2_gpu import sample_index
for fx in ("train_grads_fused3.npz", "train_grads_fused4.npz"):
    z = np.load(os.path.join(G, fx)); ns = int(z["nsample"])
    res = {}
    for path in ("auto", "generic"):
        m, x_in, loss = _run(z, path)
        errs = {}
        for k, p in m.named_parameters():
            if not bool(z["hasgrad." + k]): continue
            g = p.grad.reshape(-1).cpu().numpy()
            errs[k] = O.rel_l2(g[sample_index(k, g.size, ns)], z["sample." + k])
        res[path] = (float(loss), errs, {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None})
    print(fx, "loss ref", float(z["loss"]), "auto", res["auto"][0], "generic", res["generic"][0])
    for k in ("out_layer.4.weight", "out_layer.2.weight", "out_layer.0.weight", "blocks.0.mlp.2.weight", "patch_embed.proj.0.weight"):
        a, g = res["auto"][2][k], res["generic"][2][k]
        print(f"  {k:28s} auto-vs-ref {res['auto'][1][k]:.2e} generic-vs-ref {res['generic'][1][k]:.2e} auto-vs-generic {float((a-g).norm()/g.norm()):.2e}")
