"""The one-call training step (dpot_train_forward / dpot_train_backward, csrc/train_step.cu) against the reference's
autograd: loss, dL/dx and every parameter gradient of a 2-step autoregressive training loss (train_temporal.py:201-227)
on two geometries -- the DPOT-S width (fixture train_grads_swidth.npz) and a truncated-mode / 2-frame-bundle / SiLU /
time_agg='mlp' / cls-in-the-loss variant (train_grads_fused2.npz) and an out_layer_dim = 128 head as DPOT-L/H have it
(train_grads_fused3.npz: the generic tail on batched contractions; train_grads_fused4.npz: patch 16, block size 96 -- DPOT-L) -- and against the per-operator path of autograd.py.
Fixtures: tests/golden/make_golden_r2.py (imports the unmodified reference in the build container)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import dpot_oracle as O
from test_parity_r2_gpu import _simple_lp_loss, build_model, sample_index

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
# bar: 2e-5 on the rel-L2 of the sampled entries of every parameter gradient (see test_parity_r2_gpu.py: two correct
# fp32 evaluations of a sum over 2 * 32768 tokens differ by ~1e-5); the forward / loss bar stays 1e-5
GRAD_TOL = 2e-5


def _run(z, path):
    cfg = json.loads(str(z["cfg"]))
    B, nsteps = int(z["B"]), int(z["nsteps"])
    cw = float(z["cls_weight"]) if "cls_weight" in z else 0.0
    params = O.make_params(cfg, seed=0)
    x = O.make_input(cfg, B, seed=int(z["seed_x"]))
    rng = np.random.default_rng(int(z["seed_y"]))
    R, Co, Tb = cfg["img_size"], cfg["out_channels"], cfg["out_timesteps"]
    yy = rng.standard_normal((B, R, R, nsteps * Tb, Co)).astype(np.float32)
    msk = np.ones((B, R, R, 1, Co), dtype=np.float32)
    msk[1, ..., int(z["mask_channel"])] = 0.0
    m = build_model(cfg, params).train()
    m.train_path = path
    xx = torch.from_numpy(x).cuda().requires_grad_(True)
    x_in = xx
    yt, mt = torch.from_numpy(yy).cuda(), torch.from_numpy(msk).cuda()
    loss = 0.0
    for t in range(0, nsteps * Tb, Tb):
        im, cls = m(xx)
        loss = loss + _simple_lp_loss(im, yt[..., t:t + Tb, :], mt)
        if cw:
            loss = loss + cw * (cls * cls).sum()
        xx = torch.cat((xx[..., Tb:, :], im), dim=-2)
    loss.backward()
    return m, x_in, loss


def _check_against_fixture(z, m, x_in, loss, tol):
    ns = int(z["nsample"])
    assert float(loss) == pytest.approx(float(z["loss"]), rel=1e-5)
    errs = {}
    dx = x_in.grad.reshape(-1).cpu().numpy()
    errs["dx"] = O.rel_l2(dx[sample_index("dx", dx.size, ns)], z["dx.sample"])
    for k, p in m.named_parameters():
        if not bool(z["hasgrad." + k]):
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        assert p.grad is not None, k
        g = p.grad.reshape(-1).cpu().numpy()
        errs[k] = O.rel_l2(g[sample_index(k, g.size, ns)], z["sample." + k])
        nrm = float(np.linalg.norm(g.astype(np.float64)))
        if abs(nrm - float(z["norm." + k])) > 1e-4 * float(z["norm." + k]):
            errs[k] = max(errs[k], abs(nrm / float(z["norm." + k]) - 1.0))
    table = sorted(errs.items(), key=lambda kv: -kv[1])
    print("parameter-gradient rel-L2 against the reference autograd (worst first):")
    for k, e in table:
        print(f"  {e:.2e}  {k}")
    bad = [(k, e) for k, e in table if not e < tol]
    assert not bad, bad


@pytest.mark.parametrize("fixture", ["train_grads_swidth.npz", "train_grads_fused2.npz", "train_grads_fused3.npz", "train_grads_fused4.npz"])
def test_fused_training_step_matches_reference_autograd(fixture):
    from dpot_b200.train_engine import _TrainEngine
    z = np.load(os.path.join(G, fixture))
    m, x_in, loss = _run(z, "auto")
    assert m._train_eng is not None and m._train_eng.supported, "the one-call training step did not serve this geometry"
    _check_against_fixture(z, m, x_in, loss, GRAD_TOL)


def test_fused_training_step_matches_per_operator_path():
    """Same step through autograd.py's per-operator Functions: two implementations, one set of gradients."""
    z = np.load(os.path.join(G, "train_grads_fused2.npz"))
    m1, x1, l1 = _run(z, "auto")
    m2, x2, l2 = _run(z, "generic")
    assert m2._train_eng is None
    assert float(l1) == pytest.approx(float(l2), rel=2e-6)
    assert O.rel_l2(x1.grad.cpu().numpy(), x2.grad.cpu().numpy()) < GRAD_TOL
    for (k, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        assert (p1.grad is None) == (p2.grad is None), k
        if p1.grad is not None:
            assert O.rel_l2(p1.grad.cpu().numpy(), p2.grad.cpu().numpy()) < 5e-5, k


def test_gradients_written_into_the_exchange_arena():
    """FusedGradExchange (parallel.py): with a single process the exchange is the identity, the gradients must be the
    same numbers, live in the arena (no packing copy) and a second step must reuse it."""
    from dpot_b200.parallel import FusedGradExchange
    z = np.load(os.path.join(G, "train_grads_fused2.npz"))
    cfg = json.loads(str(z["cfg"]))
    params = O.make_params(cfg, seed=0)
    x = torch.from_numpy(O.make_input(cfg, 2, seed=5)).cuda()
    ref = build_model(cfg, params).train()
    im, _ = ref(x)
    im.square().sum().backward()
    m = build_model(cfg, params).train()
    ex = FusedGradExchange(m)
    for _ in range(2):
        for p in m.parameters():
            p.grad = None
        im, _ = m(x)
        im.square().sum().backward()
        n = ex.finish()
        lo, hi = ex.buf.data_ptr(), ex.buf.data_ptr() + 4 * ex.buf.numel()
        for (k, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
            assert (p.grad is None) == (q.grad is None), k
            if p.grad is not None:
                assert lo <= p.grad.data_ptr() < hi, k
                # (double atomics accumulate the GroupNorm statistics / bias gradients: last-bit differences between runs)
                assert O.rel_l2(p.grad.cpu().numpy(), q.grad.cpu().numpy()) < 1e-6, k
    ex.close()


def test_half_precision_operand_mode():
    """dpot_b200.set_precision("half"): fp16 operands (hi planes only), one MMA per product, fp32 accumulate -- the
    16-bit mixed-precision mode for the reference's bf16 configs (configs/pretrain_medium.yaml).  A contraction must
    equal the fp64 product of the fp16-rounded operands; the training step must stay within 1e-2 of the fp32 gradients
    (tolerance of a 16-bit mode: measured ~1e-3)."""
    import dpot_b200
    from dpot_b200 import ops
    rng = np.random.default_rng(3)
    A = rng.standard_normal((512, 320)).astype(np.float32)
    W = (rng.standard_normal((256, 320)) / 18).astype(np.float32)
    z = np.load(os.path.join(G, "train_grads_fused2.npz"))
    m32, x32, l32 = _run(z, "auto")
    prev = dpot_b200.set_precision("half")
    try:
        out = ops.gemm16(ops.split_f16(torch.from_numpy(A).cuda()), ops.split_f16(torch.from_numpy(W).cuda()))
        ref = A.astype(np.float16).astype(np.float64) @ W.astype(np.float16).astype(np.float64).T
        assert O.rel_l2(out.cpu().numpy(), ref) < 2e-6
        assert O.rel_l2(out.cpu().numpy(), A.astype(np.float64) @ W.astype(np.float64).T) > 1e-5     # it really is 16-bit
        m16, x16, l16 = _run(z, "auto")
    finally:
        dpot_b200.set_precision(prev)
    assert float(l16) == pytest.approx(float(l32), rel=2e-3)
    worst = 0.0
    for (k, p), (_, q) in zip(m16.named_parameters(), m32.named_parameters()):
        if q.grad is None:
            continue
        e = O.rel_l2(p.grad.cpu().numpy(), q.grad.cpu().numpy())
        worst = max(worst, e)
        assert e < 1e-2, (k, e)
    print("half-precision operand mode: worst parameter-gradient rel-L2 vs fp32 =", worst)


def test_standalone_modules_are_differentiable():
    """AFNO2D / Block / Mlp / PatchEmbed / TimeAggregator are ordinary differentiable nn.Modules in the reference
    (models/dpot.py:29-234).  Gradients of each against torch autograd over a plain-torch restatement of the same lines
    (fp64 on the GPU)."""
    import torch.nn.functional as F
    from dpot_b200.models.dpot import AFNO2D, Block, PatchEmbed, TimeAggregator
    torch.manual_seed(0)
    dev = "cuda"

    def afno_ref(x, w1, b1, w2, b2, nb, modes):            # models/dpot.py:51-110, channel-last x[B,H,W,C]
        B, H, W, Cc = x.shape
        bias = x
        xf = torch.fft.rfft2(x, dim=(1, 2), norm="ortho").reshape(B, H, W // 2 + 1, nb, Cc // nb)
        km = modes
        o1r = torch.zeros_like(xf.real); o1i = torch.zeros_like(xf.real)
        s = (slice(None), slice(0, km), slice(0, km))
        ein = lambda a, w: torch.einsum('...bi,bio->...bo', a, w)
        o1r[s] = F.gelu(ein(xf[s].real, w1[0]) - ein(xf[s].imag, w1[1]) + b1[0])
        o1i[s] = F.gelu(ein(xf[s].imag, w1[0]) + ein(xf[s].real, w1[1]) + b1[1])
        o2r = torch.zeros_like(o1r); o2i = torch.zeros_like(o1r)
        o2r[s] = ein(o1r[s], w2[0]) - ein(o1i[s], w2[1]) + b2[0]
        o2i[s] = ein(o1i[s], w2[0]) + ein(o1r[s], w2[1]) + b2[1]
        y = torch.fft.irfft2(torch.complex(o2r, o2i).reshape(B, H, W // 2 + 1, Cc), s=(H, W), dim=(1, 2), norm="ortho")
        return y + bias

    def check(mod, x, ref_fn, tol=2e-5):
        x1 = x.clone().requires_grad_(True)
        y = mod(x1)
        gy = torch.randn_like(y)
        (y * gy).sum().backward()
        p64 = [p.detach().double().requires_grad_(True) for p in mod.parameters()]
        x2 = x.detach().double().requires_grad_(True)
        y2 = ref_fn(x2, *p64)
        assert O.rel_l2(y.detach().cpu().numpy(), y2.detach().cpu().numpy()) < 1e-5
        (y2 * gy.double()).sum().backward()
        assert O.rel_l2(x1.grad.cpu().numpy(), x2.grad.cpu().numpy()) < tol, "dx"
        for (k, p), q in zip(mod.named_parameters(), p64):
            assert p.grad is not None, k
            assert O.rel_l2(p.grad.cpu().numpy(), q.grad.cpu().numpy()) < tol, k

    nb, E, H = 4, 64, 8
    flt = AFNO2D(width=E, num_blocks=nb, channel_first=False, modes=5, act='gelu').to(dev)
    with torch.no_grad():
        for p in flt.parameters():
            p.copy_(torch.randn_like(p) * 0.2)
    check(flt, torch.randn(2, H, H, E, device=dev), lambda x, w1, b1, w2, b2: afno_ref(x, w1, b1, w2, b2, nb, 5))

    blk = Block(width=E, n_blocks=nb, mlp_ratio=2, modes=32, double_skip=False).to(dev)
    with torch.no_grad():
        for k, p in blk.named_parameters():
            if "filter" in k:
                p.copy_(torch.randn_like(p) * 0.2)

    def block_ref(x, n1w, n1b, w1, b1, w2, b2, n2w, n2b, f1w, f1b, f2w, f2b):      # models/dpot.py:165-180
        r = x
        y = F.group_norm(x, 8, n1w, n1b, 1e-5)
        y = afno_ref(y.permute(0, 2, 3, 1), w1, b1, w2, b2, nb, 32).permute(0, 3, 1, 2)
        y = F.group_norm(y, 8, n2w, n2b, 1e-5)
        y = F.conv2d(F.gelu(F.conv2d(y, f1w, f1b)), f2w, f2b)
        return y + r
    check(blk, torch.randn(2, E, H, H, device=dev), block_ref)

    pe = PatchEmbed(img_size=32, patch_size=4, in_chans=6, embed_dim=15, out_dim=32).to(dev)
    check(pe, torch.randn(3, 6, 32, 32, device=dev),
          lambda x, w0, b0, w2, b2: F.conv2d(F.gelu(F.conv2d(x, w0, b0, stride=4)), w2, b2))

    ta = TimeAggregator(3, 5, 32, "exp_mlp").to(dev)
    with torch.no_grad():
        ta.gamma.copy_(torch.rand_like(ta.gamma) * 3)

    def ta_ref(x, w, gamma):                                # models/dpot.py:228-232
        t = torch.linspace(0, 1, x.shape[-2], device=x.device, dtype=x.dtype).unsqueeze(-1)
        return torch.einsum('tij,...ti->...j', w, x * torch.cos(t @ gamma))
    check(ta, torch.randn(2, 8, 8, 5, 32, device=dev), ta_ref)


def test_evaluation_loop_matches_the_reference_loop():
    """dpot_b200.evaluate.evaluate_loaders against evaluate.py:182-222 restated verbatim with torch ops on the same model
    (two loaders, different trajectory lengths / batch sizes, a masked channel, T_bundle = 1)."""
    from dpot_b200.evaluate import evaluate_loaders
    cfg = O.zoo_cfg("Ti", depth=2)
    m = build_model(cfg, O.make_params(cfg, seed=0))
    g = torch.Generator().manual_seed(5)
    loaders, ntests = [], []
    for B, T_ar, nb in ((3, 4, 2), (2, 3, 1)):
        batches = []
        for _ in range(nb):
            xx = torch.randn((B, 128, 128, 10, 4), generator=g)
            yy = torch.randn((B, 128, 128, T_ar, 4), generator=g)
            msk = torch.ones((B, 128, 128, 1, 4))
            msk[0, ..., 3] = 0.0
            batches.append((xx, yy, msk, torch.zeros(B, 1, dtype=torch.long)))
        loaders.append(batches)
        ntests.append(B * nb)
    fulls, steps = evaluate_loaders(m, loaders, ntests, T_bundle=1)
    with torch.no_grad():                                   # evaluate.py:182-217
        for lid, loader in enumerate(loaders):
            test_l2_full, test_l2_step = 0, 0
            for xx, yy, msk, _ in loader:
                loss = 0
                xx, yy, msk = xx.cuda(), yy.cuda(), msk.cuda()
                for t in range(0, yy.shape[-2], 1):
                    y = yy[..., t:t + 1, :]
                    im, _ = m(xx)
                    loss += _simple_lp_loss(im, y, msk)
                    pred = im if t == 0 else torch.cat((pred, im), -2)
                    xx = torch.cat((xx[..., 1:, :], im), dim=-2)
                test_l2_step += loss.item()
                test_l2_full += _simple_lp_loss(pred, yy, msk)
            assert steps[lid] == pytest.approx(test_l2_step / ntests[lid] / (yy.shape[-2] / 1), rel=2e-5)
            assert fulls[lid] == pytest.approx(float(test_l2_full) / ntests[lid], rel=2e-5)


def test_wide_mlp_training_step_with_chained_fc2():
    """mlp_ratio = 4 at the DPOT-S width (hid = 4096, as DPOT-M): the training forward chains fc2 over two <= 2048-deep
    launches -- in place for the inner block, through a scratch slot for the last block, whose result is stored split
    (csrc/train_step.cu).  The step must give the same loss and gradients with the chain limit on and off, and the same as
    the per-operator path of autograd.py (3xTF32 engine with register flushes: no chain-length effect at all)."""
    import dpot_b200
    z = np.load(os.path.join(G, "train_grads_swidth.npz"))
    cfg = dict(json.loads(str(z["cfg"])), mlp_ratio=4)
    zz = {k: z[k] for k in z.files}
    zz["cfg"] = np.array(json.dumps(cfg))
    runs = {}
    for tag, path, chain in (("chained", "auto", 2048), ("one launch", "auto", 0), ("per-operator", "generic", 2048)):
        prev = dpot_b200.set_chain(chain)
        try:
            m, x_in, loss = _run(zz, path)
        finally:
            dpot_b200.set_chain(prev)
        if path == "auto":
            assert m._train_eng is not None and m._train_eng.supported
        runs[tag] = (float(loss), x_in.grad.cpu().numpy(), {k: p.grad.cpu().numpy() for k, p in m.named_parameters() if p.grad is not None})
    l0, dx0, g0 = runs["chained"]
    for tag in ("one launch", "per-operator"):
        l1, dx1, g1 = runs[tag]
        assert l1 == pytest.approx(l0, rel=1e-5), tag
        assert set(g0) == set(g1), tag
        errs = sorted([(O.rel_l2(dx0, dx1), "dx")] + [(O.rel_l2(g0[k], g1[k]), k) for k in g0], reverse=True)
        worst, where = errs[0]
        print(f"chained vs {tag}: loss {l0:.6f} / {l1:.6f}, worst gradient rel-L2 {worst:.2e} ({where})")
        # measured: 3.8e-6 against the single launch; 9.3e-5 against the per-operator path, whose chunked weight gradients
        # on the 3xTF32 engine are the less accurate side (the one-call step holds 2e-5 against the reference autograd)
        assert worst < (2e-5 if tag == "one launch" else 2e-4), (tag, worst, where)
