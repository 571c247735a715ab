"""Kernels either side of model(xx) in the training loop (csrc/train_fused.cu, adam.cu): global gradient norm + clip folded
into the fused Adam (train_temporal.py:228-229), noise injection (:205), SimpleLpLoss forward / backward
(utils/criterion.py:38-59).  Oracles: torch CPU / numpy restatements of the reference lines, the golden loss fixtures."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import dpot_oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def test_grad_sqnorm_ragged_tensors():
    from dpot_b200 import ops
    rng = np.random.default_rng(0)
    sizes = [1, 3, 7, 4096, 32768 + 5, 1 << 20, 0, 513] + [17 * k + 1 for k in range(70)]
    gs = [torch.from_numpy(rng.standard_normal(n).astype(np.float32) * (1 + i % 5)).cuda() for i, n in enumerate(sizes)]
    want = sum(float((g.double() ** 2).sum()) for g in gs)
    got = float(ops.grad_sqnorm(gs))
    assert got == pytest.approx(want, rel=1e-6)
    # unaligned views
    base = torch.from_numpy(rng.standard_normal(1000).astype(np.float32)).cuda()
    views = [base[1:400], base[401:999]]
    assert float(ops.grad_sqnorm(views)) == pytest.approx(sum(float((v.double() ** 2).sum()) for v in views), rel=1e-6)


@pytest.mark.parametrize("max_norm", [1e4, 0.05])
def test_clip_folded_into_adam_matches_torch_clip_then_reference_adam(max_norm):
    """clip_grad_norm_ + Adam.step of the reference loop vs. our deferred clip inside the fused Adam kernel."""
    from dpot_b200.utils.clip import clip_grad_norm_
    from dpot_b200.utils.optimizer import Adam
    rng = np.random.default_rng(1)
    shapes = [(33,), (64, 7), (1024,), (5, 5, 5), (2, 300)]
    p0 = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    ps = [torch.nn.Parameter(torch.from_numpy(a.copy()).cuda()) for a in p0]
    opt = Adam(ps, lr=2e-3, betas=(0.9, 0.9), weight_decay=1e-6)
    pr = [a.copy() for a in p0]
    ms = [np.zeros_like(a) for a in p0]
    vs = [np.zeros_like(a) for a in p0]
    for step in (1, 2, 3):
        gs = [rng.standard_normal(s).astype(np.float32) * 0.01 * step for s in shapes]
        for p, g in zip(ps, gs):
            p.grad = torch.from_numpy(g.copy()).cuda()
        total = clip_grad_norm_(ps, max_norm, optimizer=opt)
        opt.step()
        tn = np.sqrt(sum(float((g.astype(np.float64) ** 2).sum()) for g in gs))
        assert float(total) == pytest.approx(tn, rel=1e-6)
        coef = min(1.0, max_norm / (np.float32(tn) + np.float32(1e-6)))
        for a, g, m, v in zip(pr, gs, ms, vs):
            O.adam_step(a, (g * np.float32(coef)).astype(np.float32), m, v, step, lr=2e-3, beta1=0.9, beta2=0.9, eps=1e-8,
                        weight_decay=1e-6)
        for p, g in zip(ps, gs):      # deferred clipping leaves .grad untouched
            assert torch.equal(p.grad.cpu(), torch.from_numpy(g))
    for p, a in zip(ps, pr):
        np.testing.assert_allclose(p.detach().cpu().numpy(), a, rtol=3e-6, atol=3e-7)
    # without an optimizer: torch semantics (in-place scaling)
    for p, s in zip(ps, shapes):
        p.grad = torch.ones(s, device="cuda")
    n_el = sum(int(np.prod(s)) for s in shapes)
    total = clip_grad_norm_(ps, 1.0)
    assert float(total) == pytest.approx(np.sqrt(n_el), rel=1e-6)
    assert float(ps[0].grad[0]) == pytest.approx(1.0 / (np.sqrt(n_el) + 1e-6), rel=1e-5)


def test_noise_injection_statistics_and_backward():
    """xx + s * ||xx||_{(X,Y,T)} * randn (train_temporal.py:205): moments per (sample, channel), determinism in
    (seed, offset), independence across offsets, and the backward through the norm against torch autograd."""
    from dpot_b200 import ops
    rng = np.random.default_rng(2)
    B, X, Y, T, C = 3, 64, 64, 10, 4
    x = torch.from_numpy((rng.standard_normal((B, X, Y, T, C)) * np.array([1.0, 2.0, 0.5, 3.0])).astype(np.float32)).cuda()
    s = 0.05
    y, sumsq = ops.noise_inject(x, s, seed=1234, offset=7)
    y2, _ = ops.noise_inject(x, s, seed=1234, offset=7)
    y3, _ = ops.noise_inject(x, s, seed=1234, offset=8)
    assert torch.equal(y, y2) and not torch.equal(y, y3)
    nrm = torch.sum(x.double() ** 2, dim=(1, 2, 3), keepdim=True) ** 0.5
    np.testing.assert_allclose(sumsq.cpu().numpy(), (nrm ** 2).reshape(B, C).cpu().numpy(), rtol=1e-6)
    eps = ((y.double() - x.double()) / (s * nrm))                       # should be N(0, 1)
    eps3 = ((y3.double() - x.double()) / (s * nrm))
    n = X * Y * T
    m = eps.mean(dim=(1, 2, 3)).abs().max().item()
    v = eps.var(dim=(1, 2, 3)).sub(1).abs().max().item()
    assert m < 5 / np.sqrt(n) and v < 8 * np.sqrt(2 / n), (m, v)
    k4 = (eps ** 4).mean().item()
    assert abs(k4 - 3.0) < 0.05                                        # Gaussian kurtosis
    assert abs((eps * eps3).mean().item()) < 5 / np.sqrt(B * n * C)    # streams of different offsets are uncorrelated
    assert abs((eps[..., 0] * eps[..., 1]).mean().item()) < 5 / np.sqrt(B * n)
    # backward: d/dx of x + s * ||x|| * eps with eps constant
    dy = torch.from_numpy(rng.standard_normal((B, X, Y, T, C)).astype(np.float32)).cuda()
    dx = ops.noise_inject_bwd(x, dy, s, 1234, 7, sumsq)
    xr = x.double().requires_grad_(True)
    yr = xr + s * torch.sum(xr ** 2, dim=(1, 2, 3), keepdim=True) ** 0.5 * eps.detach()
    yr.backward(dy.double())
    assert O.rel_l2(dx.cpu().numpy(), xr.grad.cpu().numpy()) < 2e-6


def test_simple_lp_loss_forward_backward():
    from dpot_b200 import ops
    z = np.load(os.path.join(G, "train_grads_tiny.npz"))
    a, b, mk = (torch.from_numpy(z[k]).cuda() for k in ("loss.a", "loss.b", "loss.mask"))
    loss, _ = ops.lp_loss(a, b, mk)
    assert float(loss) == pytest.approx(float(z["loss.masked"]), rel=2e-6)
    loss, _ = ops.lp_loss(a, b, None)
    assert float(loss) == pytest.approx(float(z["loss.nomask"]), rel=2e-6)
    # accumulate over AR steps + backward against torch autograd of the restated formula
    rng = np.random.default_rng(3)
    B, X, T, C = 4, 32, 2, 4
    x = torch.from_numpy(rng.standard_normal((B, X, X, T, C)).astype(np.float32)).cuda()
    y = torch.from_numpy(rng.standard_normal((B, X, X, T, C)).astype(np.float32)).cuda()
    msk = torch.ones((B, X, X, 1, C), device="cuda")
    msk[1, ..., 2] = 0
    msk[3, ..., :3] = 0

    def ref(xx):
        n = xx.shape[0]
        xm, ym = xx * msk, y.double() * msk
        nch = msk.sum(dim=(1, 2, 3)).count_nonzero(dim=-1)
        d = torch.norm(xm.reshape(n, -1, C) - ym.reshape(n, -1, C), 2, dim=1)
        yn = torch.norm(ym.reshape(n, -1, C), 2, dim=1) + 1e-8
        return torch.sum(torch.sum(d / yn, dim=-1) / nch)
    xr = x.double().requires_grad_(True)
    lr = ref(xr)
    lr.backward()
    loss, coef = ops.lp_loss(x, y, msk)
    assert float(loss) == pytest.approx(float(lr), rel=2e-6)
    dx = ops.lp_loss_bwd(x, y, msk, coef)
    assert O.rel_l2(dx.cpu().numpy(), xr.grad.cpu().numpy()) < 2e-6
    loss2, _ = ops.lp_loss(x, y, msk, loss=loss, accumulate=True)
    assert float(loss2) == pytest.approx(2 * float(lr), rel=2e-6)
    g = torch.full((1,), 0.5, device="cuda")
    assert O.rel_l2(ops.lp_loss_bwd(x, y, msk, coef, g).cpu().numpy(), 0.5 * xr.grad.cpu().numpy()) < 2e-6


def test_ar_train_step_fast_path_matches_reference_trajectory():
    """dpot_b200.train.ar_train_step (loss kernel, deferred clip inside the fused Adam, no host sync) over the 3-step
    reference trajectory fixture (train_temporal.py:189-230 with noise off): losses and final weights."""
    from dpot_b200.models.dpot import DPOTNet
    from dpot_b200.train import ar_train_step
    from dpot_b200.utils.optimizer import Adam
    z = np.load(os.path.join(G, "train_traj_tiny.npz"))
    cfg = json.loads(str(z["cfg"]))
    m = DPOTNet(**cfg)
    m.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in O.make_params(cfg, seed=0).items()})
    m = m.cuda().train()
    opt = Adam(m.parameters(), lr=1e-3, betas=(0.9, 0.9), weight_decay=1e-6)
    msk = torch.ones((int(z["B"]), cfg["img_size"], cfg["img_size"], 1, cfg["out_channels"]), device="cuda")
    for it in range(3):
        opt.param_groups[0]["lr"] = float(z["lrs"][it])
        loss = ar_train_step(m, opt, torch.from_numpy(z["xs"][it]).cuda(), torch.from_numpy(z["ys"][it]).cuda(), msk,
                             T_bundle=cfg["out_timesteps"], noise_scale=0.0, grad_clip=10000.0, step=it)
        assert float(loss) == pytest.approx(float(z["losses"][it]), rel=5e-5), it
    for k, v in m.state_dict().items():
        ref = z["final." + k]
        err = np.abs(v.cpu().numpy() - ref).max()
        assert err < 2e-4 * max(1.0, np.abs(ref).max()), (k, err)
    # with noise: the step runs, the loss stays finite and differs from the noise-free one
    l_noise = ar_train_step(m, opt, torch.from_numpy(z["xs"][0]).cuda(), torch.from_numpy(z["ys"][0]).cuda(), msk,
                            T_bundle=cfg["out_timesteps"], noise_scale=5e-4, grad_clip=10000.0, seed=3, step=9)
    assert torch.isfinite(l_noise)


def test_lamb_matches_reference_golden():
    """The fused Lamb (csrc/lamb.cu) against trajectories of the unmodified reference Lamb, utils/optimizer.py:359-499:
    the scripts' variant (adam=True), the trust-ratio path, debias and clamp; one all-zero parameter (trust_ratio = 1)."""
    import json
    from dpot_b200.utils.optimizer import Lamb
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "lamb.npz"))
    kws = json.loads(str(z["kw"]))
    for tag, kw in kws.items():
        kw = dict(kw, betas=tuple(kw["betas"]))
        ps = [torch.nn.Parameter(torch.from_numpy(z[f"p0.{i}"].copy()).cuda()) for i in range(3)]
        opt = Lamb(ps, lr=1e-3, eps=1e-6, **kw)
        for s in range(len(z["lrs"])):
            opt.param_groups[0]["lr"] = float(z["lrs"][s])
            for i, p in enumerate(ps):
                p.grad = torch.from_numpy(z[f"grads.{i}"][s].copy()).cuda()
            opt.step()
            for i, p in enumerate(ps):
                st = opt.state[p]
                np.testing.assert_allclose(p.detach().cpu().numpy(), z[f"{tag}.p.{i}"][s], rtol=3e-6, atol=3e-7,
                                           err_msg=f"{tag} tensor {i} step {s}")
                got = [float(st["weight_norm"]), float(st["adam_norm"]), float(st["trust_ratio"])]
                np.testing.assert_allclose(got, z[f"{tag}.info.{i}"][s], rtol=3e-6, err_msg=f"{tag} info {i} step {s}")
        for i, p in enumerate(ps):
            st = opt.state[p]
            assert st["step"] == len(z["lrs"])
            np.testing.assert_allclose(st["exp_avg"].cpu().numpy(), z[f"{tag}.m.{i}"], rtol=3e-6, atol=1e-7)
            np.testing.assert_allclose(st["exp_avg_sq"].cpu().numpy(), z[f"{tag}.v.{i}"], rtol=3e-6, atol=1e-7)


def test_lamb_many_ragged_tensors_and_bandwidth():
    """98 tensors (two launches per stage), sizes from 1 element to several blocks incl. odd and unaligned views, against
    the numpy oracle of the reference's step; then the step's HBM rate on a DPOT-S-sized parameter set (40 B / param)."""
    from dpot_b200.utils.optimizer import Lamb
    from oracle import dpot_oracle as O
    rng = np.random.default_rng(9)
    sizes = [1, 3, 5, 31, 257, 4096, 4097, 12289, 70001] * 11
    sizes = sizes[:98]
    flat = torch.from_numpy(rng.standard_normal(sum(sizes) + 1).astype(np.float32)).cuda()
    ps, off = [], 1                                             # views at odd offsets: the unaligned scalar path
    for n in sizes:
        ps.append(torch.nn.Parameter(flat[off:off + n])); off += n
    ref = [p.detach().cpu().numpy().copy() for p in ps]
    rm, rv = [np.zeros_like(a) for a in ref], [np.zeros_like(a) for a in ref]
    opt = Lamb(ps, lr=2e-3, betas=(0.9, 0.99), weight_decay=1e-2, debias=True)
    for s in range(1, 4):
        for i, p in enumerate(ps):
            g = rng.standard_normal(sizes[i]).astype(np.float32)
            p.grad = torch.from_numpy(g).cuda()
            O.lamb_step(ref[i], g, rm[i], rv[i], s, lr=2e-3, beta1=0.9, beta2=0.99, eps=1e-6, weight_decay=1e-2, debias=True)
        opt.step()
    worst = max(float(np.abs(p.detach().cpu().numpy() - r).max() / (np.abs(r).max() + 1e-30)) for p, r in zip(ps, ref))
    assert worst < 5e-6, worst
    # bandwidth: 30.8 M parameters in 91 tensors
    big = [torch.nn.Parameter(torch.randn(n, device="cuda")) for n in [338461] * 91]
    for p in big:
        p.grad = torch.randn_like(p)
    opt = Lamb(big, lr=1e-3, weight_decay=1e-4, adam=True)
    for _ in range(3):
        opt.step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        opt.step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    n = sum(p.numel() for p in big)
    print(f"fused Lamb: {n / 1e6:.1f} M parameters, {ms * 1e3:.0f} us / step (python loop included) = {40 * n / ms / 1e6:.0f} GB/s of 40 B / parameter")
