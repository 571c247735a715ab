"""Generate golden fixtures by running the UNMODIFIED reference (imported read-only from
/root/reference) on seeded synthetic weights/inputs from oracle.dpot_oracle.make_params /
make_input.  Run in the build container only:  python tests/golden/make_golden.py

Weights and inputs are NOT stored (they are re-derived from numpy Generator seeds, which are
platform independent); each fixture stores the config, the seeds, a float64 checksum of the
weights/inputs, and the reference outputs (fp32, CPU).
"""
import json
import os
import sys

import numpy as np

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import torch  # noqa: E402

from models.dpot import AFNO2D, DPOTNet  # noqa: E402  (reference)
from utils.criterion import SimpleLpLoss  # noqa: E402  (reference)
from utils.optimizer import Adam, AdamW  # noqa: E402  (reference)

from oracle import dpot_oracle as O  # noqa: E402

torch.set_num_threads(8)

CASES = {
    # name: (cfg, B, input kind, rollout steps)
    "tiny_trunc": (O.make_cfg(img_size=32, patch_size=4, in_channels=3, out_channels=3, in_timesteps=5,
                              out_timesteps=1, n_blocks=4, embed_dim=32, out_layer_dim=16, depth=2, modes=3,
                              mlp_ratio=2, n_cls=5), 3, "randn", 2),
    "tiny_bundle": (O.make_cfg(img_size=32, patch_size=4, in_channels=3, out_channels=3, in_timesteps=5,
                               out_timesteps=2, n_blocks=4, embed_dim=32, out_layer_dim=16, depth=2, modes=32,
                               mlp_ratio=2, n_cls=5), 2, "randn", 2),
    "tiny_norm_tanh_mlp": (O.make_cfg(img_size=32, patch_size=4, in_channels=3, out_channels=3, in_timesteps=5,
                                      out_timesteps=1, n_blocks=2, embed_dim=32, out_layer_dim=16, depth=2,
                                      modes=32, mlp_ratio=1, n_cls=5, normalize=True, act="tanh",
                                      time_agg="mlp"), 2, "randn", 1),
    "smoke_20": (O.make_cfg(img_size=20, patch_size=5, in_channels=3, out_channels=3, in_timesteps=6,
                            out_timesteps=1, embed_dim=32, normalize=True), 4, "randn", 1),  # models/dpot.py:462-468
    "c1_ti64": (O.zoo_cfg("Ti", img_size=64), 1, "ns2d", 1),
    "c2_s128": (O.zoo_cfg("S"), 1, "randn", 3),
}
for _a in ["tanh", "sigmoid", "relu", "leaky_relu", "softplus", "ELU", "silu"]:
    CASES[f"act_{_a}"] = (O.make_cfg(img_size=16, patch_size=4, in_channels=2, out_channels=2, in_timesteps=3,
                                     out_timesteps=1, n_blocks=2, embed_dim=16, out_layer_dim=8, depth=1,
                                     modes=32, mlp_ratio=1, n_cls=3, act=_a), 2, "randn", 1)


def ref_model(cfg, params):
    m = DPOTNet(**cfg)
    sd = {k: torch.from_numpy(v.copy()) for k, v in params.items()}
    m.load_state_dict(sd, strict=True)
    return m.eval()


def checksum(d):
    if isinstance(d, dict):
        return float(sum(np.asarray(v, dtype=np.float64).sum() for v in d.values()))
    return float(np.asarray(d, dtype=np.float64).sum())


def gen_forward(name):
    cfg, B, kind, nsteps = CASES[name]
    params = O.make_params(cfg, seed=0)
    x = O.make_input(cfg, B, seed=0, kind=kind)
    m = ref_model(cfg, params)
    Tb = cfg["out_timesteps"]
    with torch.no_grad():
        xx = torch.from_numpy(x)
        y1, cls1 = m(xx)
        preds = []
        for _ in range(nsteps):
            im, _ = m(xx)
            preds.append(im)
            xx = torch.cat((xx[..., Tb:, :], im), dim=-2)
        pred = torch.cat(preds, dim=-2)
    np.savez_compressed(
        os.path.join(HERE, f"fwd_{name}.npz"), cfg=json.dumps(cfg), B=B, kind=kind, nsteps=nsteps,
        params_checksum=checksum(params), input_checksum=checksum(x),
        y=y1.numpy(), cls=cls1.numpy(), pred=pred.numpy())
    print(name, "y", tuple(y1.shape), "pred", tuple(pred.shape))


def gen_afno():
    """AFNO2D(x) - x per module with re-scaled weights (SURVEY.md finding 3), incl. real truncation."""
    out = {}
    for tag, (E, nb, H, modes) in {"e32_h8_m3": (32, 4, 8, 3), "e64_h16_m32": (64, 2, 16, 32),
                                   "e32_h4_m2": (32, 4, 4, 2)}.items():
        rng = np.random.default_rng(7)
        bs = E // nb
        w = {k: (rng.standard_normal(s) * sc).astype(np.float32) for k, s, sc in [
            ("w1", (2, nb, bs, bs), bs ** -0.5), ("b1", (2, nb, bs), 0.1),
            ("w2", (2, nb, bs, bs), bs ** -0.5), ("b2", (2, nb, bs), 0.1)]}
        x = rng.standard_normal((2, E, H, H)).astype(np.float32)
        f = AFNO2D(width=E, num_blocks=nb, channel_first=True, modes=modes)
        f.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
        with torch.no_grad():
            y = f(torch.from_numpy(x)).numpy()
        for k, v in w.items():
            out[f"{tag}.{k}"] = v
        out[f"{tag}.x"] = x
        out[f"{tag}.delta"] = y - x
        out[f"{tag}.meta"] = np.array([E, nb, H, modes])
    np.savez_compressed(os.path.join(HERE, "afno2d.npz"), **out)
    print("afno2d ok")


def gen_adam():
    out = {}
    rng = np.random.default_rng(3)
    n, nsteps = 257, 4
    p0 = rng.standard_normal(n).astype(np.float32)
    grads = rng.standard_normal((nsteps, n)).astype(np.float32)
    lrs = [1e-3, 5e-4, 2e-3, 1e-3]
    out["p0"], out["grads"], out["lrs"] = p0, grads, np.array(lrs)
    for tag, cls, kw in [("adam", Adam, dict(betas=(0.9, 0.9), weight_decay=1e-6)),
                         ("adam_wd0", Adam, dict(betas=(0.9, 0.999), weight_decay=0.0)),
                         ("adam_ams", Adam, dict(betas=(0.9, 0.99), weight_decay=1e-2, amsgrad=True)),
                         ("adamw", AdamW, dict(betas=(0.9, 0.999), weight_decay=1e-2)),
                         ("adamw_ams", AdamW, dict(betas=(0.9, 0.9), weight_decay=1e-1, amsgrad=True))]:
        p = torch.nn.Parameter(torch.from_numpy(p0.copy()))
        opt = cls([p], lr=1e-3, eps=1e-8, **kw)
        traj = []
        for s in range(nsteps):
            opt.param_groups[0]["lr"] = lrs[s]
            p.grad = torch.from_numpy(grads[s].copy())
            opt.step()
            traj.append(p.detach().numpy().copy())
        st = opt.state[p]
        out[f"{tag}.p"] = np.stack(traj)
        out[f"{tag}.m"] = st["exp_avg"].numpy()
        out[f"{tag}.v"] = st["exp_avg_sq"].numpy()
        out[f"{tag}.kw"] = json.dumps({k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()})
    np.savez_compressed(os.path.join(HERE, "adam.npz"), **out)
    print("adam ok")


def gen_loss_and_grads():
    """SimpleLpLoss values and reference autograd gradients of a 2-step AR training loss
    (train_temporal.py:201-227 with noise_scale=0) for the tiny_trunc config."""
    cfg, B, kind, nsteps = CASES["tiny_trunc"]
    params = O.make_params(cfg, seed=0)
    x = O.make_input(cfg, B, seed=0, kind=kind)
    rng = np.random.default_rng(5)
    yy = rng.standard_normal((B, cfg["img_size"], cfg["img_size"], nsteps, cfg["out_channels"])).astype(np.float32)
    msk = np.ones((B, cfg["img_size"], cfg["img_size"], 1, cfg["out_channels"]), dtype=np.float32)
    msk[1, ..., 2] = 0.0  # one inactive channel in sample 1
    m = ref_model(cfg, params).train()
    myloss = SimpleLpLoss(size_average=False)
    xx = torch.from_numpy(x).requires_grad_(True)
    x_in = xx
    loss = 0.0
    Tb = cfg["out_timesteps"]
    yt, mt = torch.from_numpy(yy), torch.from_numpy(msk)
    for t in range(0, nsteps, Tb):
        im, _ = m(xx)
        loss = loss + myloss(im, yt[..., t:t + Tb, :], mask=mt)
        xx = torch.cat((xx[..., Tb:, :], im), dim=-2)
    loss.backward()
    out = dict(cfg=json.dumps(cfg), B=B, yy=yy, msk=msk, loss=np.float32(loss.item()), dx=x_in.grad.numpy())
    for k, v in m.named_parameters():
        out["grad." + k] = np.zeros(v.shape, np.float32) if v.grad is None else v.grad.numpy()
        out["hasgrad." + k] = np.array(v.grad is not None)
    # plain loss fixtures
    a = rng.standard_normal((3, 8, 8, 2, 4)).astype(np.float32)
    b = rng.standard_normal((3, 8, 8, 2, 4)).astype(np.float32)
    mk = np.ones((3, 8, 8, 1, 4), np.float32)
    mk[0, ..., 1:] = 0
    out["loss.a"], out["loss.b"], out["loss.mask"] = a, b, mk
    out["loss.masked"] = np.float32(myloss(torch.from_numpy(a), torch.from_numpy(b), mask=torch.from_numpy(mk)).item())
    out["loss.nomask"] = np.float32(myloss(torch.from_numpy(a), torch.from_numpy(b)).item())
    np.savez_compressed(os.path.join(HERE, "train_grads_tiny.npz"), **out)
    print("grads ok, loss", loss.item())


if __name__ == "__main__" and "--traj" not in sys.argv:
    for name in CASES:
        gen_forward(name)
    gen_afno()
    gen_adam()
    gen_loss_and_grads()


def gen_train_trajectory():
    """3 optimizer steps of the loop of train_temporal.py:189-230 (noise_scale=0, Adam(betas=(0.9,0.9), wd=1e-6),
    clip_grad_norm_(1e4), OneCycle-like lr changes) on the tiny_trunc config: losses and final weights."""
    cfg, B, kind, nsteps = CASES["tiny_trunc"]
    params = O.make_params(cfg, seed=0)
    m = ref_model(cfg, params).train()
    opt = Adam(m.parameters(), lr=1e-3, betas=(0.9, 0.9), weight_decay=1e-6)
    myloss = SimpleLpLoss(size_average=False)
    rng = np.random.default_rng(11)
    Tb = cfg["out_timesteps"]
    losses, lrs = [], [1e-3, 2e-3, 5e-4]
    xs, ys = [], []
    msk = torch.ones((B, cfg["img_size"], cfg["img_size"], 1, cfg["out_channels"]))
    for it in range(3):
        x = rng.standard_normal((B, cfg["img_size"], cfg["img_size"], cfg["in_timesteps"], cfg["in_channels"])).astype(np.float32)
        yy = rng.standard_normal((B, cfg["img_size"], cfg["img_size"], nsteps, cfg["out_channels"])).astype(np.float32)
        xs.append(x); ys.append(yy)
        xx, yt = torch.from_numpy(x), torch.from_numpy(yy)
        loss = 0.0
        for t in range(0, nsteps, Tb):
            im, _ = m(xx)
            loss = loss + myloss(im, yt[..., t:t + Tb, :], mask=msk)
            xx = torch.cat((xx[..., Tb:, :], im), dim=-2)
        opt.param_groups[0]["lr"] = lrs[it]
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), 10000.0)
        opt.step()
        losses.append(loss.item())
    out = dict(cfg=json.dumps(cfg), B=B, losses=np.array(losses, np.float32), lrs=np.array(lrs), xs=np.stack(xs), ys=np.stack(ys))
    for k, v in m.state_dict().items():
        out["final." + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "train_traj_tiny.npz"), **out)
    print("trajectory ok", losses)


if __name__ == "__main__" and "--traj" in sys.argv:
    gen_train_trajectory()
