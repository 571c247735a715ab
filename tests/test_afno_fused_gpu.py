"""The fused AFNO2D mixer kernel (dpot_afno_fused, csrc/afno_fused.cu) against float64 numpy restatements of
models/dpot.py:51-110 + the GroupNorm around it, stage by stage (operand-tile images through the kernel's test hook)
and end to end, and against the four-kernel path it replaces."""
import numpy as np
import pytest
import torch

from oracle import dpot_oracle as O

pytestmark = pytest.mark.gpu
PLANE = 144 * 128          # bytes of one operand plane of one k-block
OPER = 8 * PLANE


def _supported(h, E, nb):
    from dpot_b200 import _lib
    return bool(_lib.load().dpot_afno_fused_supported(h, E, nb, h, h // 2 + 1, 8))


def _make(B, nb, seed, wscale=None, bscale=0.1):
    rng = np.random.default_rng(seed)
    bs, E = 128, 128 * nb
    lat = (rng.standard_normal((B * 256, E)) * (1.0 + rng.random(E)) + 0.3 * rng.standard_normal(E)).astype(np.float32)
    if wscale is None:
        wscale = bs ** -0.5
    w = {k: (rng.standard_normal(s) * sc).astype(np.float32) for k, s, sc in [
        ("w1", (2, nb, bs, bs), wscale), ("b1", (2, nb, bs), bscale), ("w2", (2, nb, bs, bs), wscale), ("b2", (2, nb, bs), bscale)]}
    gamma = (1.0 + 0.1 * rng.standard_normal(E)).astype(np.float32)
    beta = (0.1 * rng.standard_normal(E)).astype(np.float32)
    return lat, w, gamma, beta


def _reference(lat, w, gamma, beta, B, nb, act="gelu"):
    """float64: (spectrum X[B,16,9,E] complex, hidden O1, f, n1)."""
    E = lat.shape[1]
    bs = E // nb
    a = lat.astype(np.float64).reshape(B, 256, E)
    g = a.reshape(B, 256, 8, E // 8)
    mean = g.mean(axis=(1, 3), keepdims=True)
    var = g.var(axis=(1, 3), keepdims=True)
    n1 = ((g - mean) / np.sqrt(var + 1e-5)).reshape(B, 256, E) * gamma.astype(np.float64) + beta.astype(np.float64)
    x = n1.reshape(B, 16, 16, E)
    X = np.fft.rfft2(x, axes=(1, 2), norm="ortho")                              # [B,16,9,E]
    Xb = X.reshape(B, 16, 9, nb, bs)
    w1 = w["w1"].astype(np.float64); w2 = w["w2"].astype(np.float64)
    b1 = w["b1"].astype(np.float64); b2 = w["b2"].astype(np.float64)
    ein = lambda v, m: np.einsum("...bi,bio->...bo", v, m)
    o1r = O.activation(ein(Xb.real, w1[0]) - ein(Xb.imag, w1[1]) + b1[0], act)
    o1i = O.activation(ein(Xb.imag, w1[0]) + ein(Xb.real, w1[1]) + b1[1], act)
    o2r = ein(o1r, w2[0]) - ein(o1i, w2[1]) + b2[0]
    o2i = ein(o1i, w2[0]) + ein(o1r, w2[1]) + b2[1]
    o2 = (o2r + 1j * o2i).reshape(B, 16, 9, E)
    y = O.irfft2_torch(o2, 16, 16) if hasattr(O, "irfft2_torch") else None
    if y is None:   # torch semantics: complex inverse along k1, then c2r along k2 (ignores Im of the k2 = 0 / 8 columns)
        t = np.fft.ifft(o2, axis=1, norm="ortho")
        y = np.fft.irfft(t, n=16, axis=2, norm="ortho")
    f = y.reshape(B, 256, E) + n1
    return X, (o1r + 1j * o1i), f.reshape(B * 256, E), n1.reshape(B * 256, E)


def _decode_operand(img_bytes: np.ndarray) -> np.ndarray:
    """One operand-tile image (OPER bytes) -> values [144 modes (m = k2*16 + k1), 256 k (re 128 | im 128)] (float64)."""
    h = img_bytes.view(np.float16)
    out = np.zeros((144, 256), dtype=np.float64)
    m = np.arange(144)[:, None]
    k = np.arange(64)[None, :]
    for kb in range(4):
        idx = m * 64 + (((k >> 3) ^ (m & 7)) << 3) + (k & 7)
        hi = h[(kb * 2) * (PLANE // 2):][idx].astype(np.float64)
        lo = h[(kb * 2 + 1) * (PLANE // 2):][idx].astype(np.float64)
        out[:, kb * 64:(kb + 1) * 64] = hi + lo / 2048.0
    return out


def _run(lat, w, gamma, beta, B, nb, act="gelu", debug=False):
    from dpot_b200 import ops
    t = lambda v: torch.from_numpy(v).cuda()
    latt = t(lat)
    stats1 = ops.gn_stats(latt, B, 256)
    return ops.afno_fused(latt, stats1, t(gamma), t(beta), t(w["w1"]), t(w["b1"]), t(w["w2"]), t(w["b2"]), B, 16, act=act,
                          debug=debug)


@pytest.mark.parametrize("B,nb", [(1, 2), (3, 2), (2, 8), (37, 4)], ids=["B1_E256", "B3_E256", "B2_E1024", "B37_E512_multi_unit"])
def test_fused_mixer_stage_by_stage(B, nb):
    if not _supported(16, 128 * nb, nb):
        pytest.skip("fused AFNO mixer not available on this device")
    lat, w, gamma, beta = _make(B, nb, seed=B * 10 + nb)
    X, O1, f_ref, n1 = _reference(lat, w, gamma, beta, B, nb)
    f, stats2, dbg = _run(lat, w, gamma, beta, B, nb, debug=True)
    torch.cuda.synchronize()
    dbg = dbg.cpu().numpy().view(np.uint8).reshape(B * nb, 2, OPER)
    # stage 1 / 2: the operand tiles of a few units
    for u in sorted({0, B * nb - 1, (B * nb) // 2}):
        b, kap = divmod(u, nb)
        for stage, ref in ((0, X), (1, O1.reshape(B, 16, 9, 128 * nb))):
            got = _decode_operand(dbg[u, stage])                                  # [m = k2*16 + k1, re | im]
            blk = ref[b][:, :, kap * 128:(kap + 1) * 128]                         # [k1, k2, 128]
            want = np.concatenate([blk.real, blk.imag], axis=-1).transpose(1, 0, 2).reshape(144, 256)
            e = O.rel_l2(got, want)
            assert e < 2e-6, (("X", "O1")[stage], u, e)
    # end to end: the spectral branch alone (f - skip) and f
    got = f.cpu().numpy().astype(np.float64)
    assert O.rel_l2(got - n1, f_ref - n1) < 1e-5
    assert O.rel_l2(got, f_ref) < 2e-6
    # GroupNorm-2 statistics of f
    fr = f_ref.reshape(B, 256, 8, -1)
    want = np.stack([fr.sum(axis=(1, 3)), (fr ** 2).sum(axis=(1, 3))], axis=-1)
    np.testing.assert_allclose(stats2.cpu().numpy(), want, rtol=2e-5, atol=1e-3)


@pytest.mark.parametrize("act", ["gelu", "tanh", "silu"])
def test_fused_mixer_reference_default_init_and_activations(act):
    """The reference's own initialisation (scale * rand, scale = 1 / bs^2 ~ 6e-5, models/dpot.py:41-48): the spectral
    branch is ~1e-4 of the signal; the per-layer power-of-two weight scale must keep it at full precision."""
    nb, B = 2, 2
    if not _supported(16, 128 * nb, nb):
        pytest.skip("fused AFNO mixer not available on this device")
    rng = np.random.default_rng(5)
    lat, w, gamma, beta = _make(B, nb, seed=77)
    sc = 1.0 / (128 * 128)
    for k in w:
        w[k] = (sc * rng.random(w[k].shape)).astype(np.float32)
    _, _, f_ref, n1 = _reference(lat, w, gamma, beta, B, nb, act=act)
    f, _ = _run(lat, w, gamma, beta, B, nb, act=act)
    got = f.cpu().numpy().astype(np.float64)
    # the branch itself sits ~1e-4 below the skip it is added to: the fp32 rounding of f (6e-8 relative) is the floor
    floor = 6e-8 * np.linalg.norm(f_ref) / np.linalg.norm(f_ref - n1)
    assert O.rel_l2(got - n1, f_ref - n1) < max(1e-5, 4 * floor), floor
    assert O.rel_l2(got, f_ref) < 2e-7


def test_fused_mixer_equals_four_kernel_path_in_the_model():
    """DPOT-S forward with the fused mixer vs. with the four-kernel path it replaces (dpot_afno_set_fused)."""
    from dpot_b200 import _lib
    from dpot_b200.models.dpot import DPOTNet
    lib = _lib.load()
    cfg = O.zoo_cfg("S", depth=2)
    if not _supported(16, 1024, 8):
        pytest.skip("fused AFNO mixer not available on this device")
    params = O.make_params(cfg, seed=0)
    x = torch.from_numpy(O.make_input(cfg, 3, seed=1)).cuda()
    outs = []
    for mode in (-1, 0):
        lib.dpot_afno_set_fused(mode)
        try:
            m = DPOTNet(**cfg)
            m.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in params.items()})
            m = m.cuda().eval()
            l0 = lib.dpot_launch_count()
            with torch.no_grad():
                y, cls = m(x)
            torch.cuda.synchronize()
            outs.append((y.cpu().numpy(), cls.cpu().numpy(), lib.dpot_launch_count() - l0))
        finally:
            lib.dpot_afno_set_fused(-1)
    yo, co = O.dpot_forward(x.cpu().numpy(), params, cfg)
    assert O.rel_l2(outs[0][0], yo) < 1e-5 and O.rel_l2(outs[0][1], co) < 1e-5
    assert O.rel_l2(outs[0][0], outs[1][0]) < 2e-6
    assert outs[0][2] < outs[1][2]          # fewer launches: 4 kernels -> 1 per block


@pytest.mark.parametrize("B,nb", [(3, 8), (2, 4)])
def test_fused_mixer_applies_groupnorm2_itself(B, nb):
    """dpot_afno_fused_gn2: the unit (sample, channel block) covers whole GroupNorm groups (128 channels at E = 1024,
    64 at E = 512), so the kernel normalises f itself and writes the channel MLP's split-fp16 operand: against
    GroupNorm-2 of the fp64 reference f, and against the separate pass it replaces."""
    from dpot_b200 import ops
    if not _supported(16, 128 * nb, nb):
        pytest.skip("fused AFNO mixer not available on this device")
    lat, w, gamma, beta = _make(B, nb, seed=21 + nb)
    E = 128 * nb
    rng = np.random.default_rng(5)
    g2 = (1.0 + 0.1 * rng.standard_normal(E)).astype(np.float32)
    b2 = (0.1 * rng.standard_normal(E)).astype(np.float32)
    dev = lambda a: torch.from_numpy(a).cuda()
    lt = dev(lat)
    st1 = ops.gn_stats(lt, B, 256)
    g2d, b2d = dev(g2), dev(b2)
    f, st2, n2 = ops.afno_fused(lt, st1, dev(gamma), dev(beta), dev(w["w1"]), dev(w["b1"]), dev(w["w2"]), dev(w["b2"]), B, 16,
                                gn2=(g2d, b2d))
    _, _, f_ref, _ = _reference(lat, w, gamma, beta, B, nb)
    fr = f_ref.reshape(B, 256, 8, E // 8)
    mean = fr.mean(axis=(1, 3), keepdims=True)
    var = fr.var(axis=(1, 3), keepdims=True)
    want = ((fr - mean) / np.sqrt(var + 1e-5)).reshape(B * 256, E) * g2.astype(np.float64) + b2.astype(np.float64)
    got = ops.unsplit_f16(n2).cpu().numpy()
    assert O.rel_l2(got, want) < 3e-6
    # the statistics the kernel still reports, and the separate GroupNorm-2 + split pass on the same f
    sep = torch.empty_like(n2)
    from dpot_b200 import _lib
    from dpot_b200._lib import check, ptr
    check(_lib.load().dpot_split_f16_gn(ptr(f), E, B * 256, E, ptr(st2), ptr(g2d), ptr(b2d), 8, 1e-5, 256, ptr(sep),
                                        2 * E, E, torch.cuda.current_stream().cuda_stream), "dpot_split_f16_gn")
    assert O.rel_l2(got, ops.unsplit_f16(sep).cpu().numpy()) < 1e-6


def test_model_forward_with_groupnorm2_inside_the_mixer():
    """dpot_afno_set_fused_gn2(1): the forward without the separate GroupNorm-2 + split launches gives the same output."""
    from dpot_b200 import _lib
    from dpot_b200.models.dpot import DPOTNet
    lib = _lib.load()
    if not _supported(16, 1024, 8):
        pytest.skip("fused AFNO mixer not available on this device")
    cfg = O.zoo_cfg("S", depth=2)
    params = O.make_params(cfg, seed=0)
    x = torch.from_numpy(O.make_input(cfg, 3, seed=1)).cuda()
    outs = []
    default = lib.dpot_afno_set_fused_gn2(-1)
    for on in (0, 1):
        lib.dpot_afno_set_fused_gn2(on)
        try:
            m = DPOTNet(**cfg)
            m.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in params.items()})
            m = m.cuda().eval()
            l0 = lib.dpot_launch_count()
            with torch.no_grad():
                y, cls = m(x)
            torch.cuda.synchronize()
            outs.append((y.cpu().numpy(), cls.cpu().numpy(), lib.dpot_launch_count() - l0))
        finally:
            lib.dpot_afno_set_fused_gn2(default)
    yo, _ = O.dpot_forward(x.cpu().numpy(), params, cfg)
    assert O.rel_l2(outs[1][0], yo) < 1e-5
    assert O.rel_l2(outs[0][0], outs[1][0]) < 2e-6 and O.rel_l2(outs[0][1], outs[1][1]) < 2e-6
    assert outs[1][2] == outs[0][2] - cfg["depth"]          # one launch fewer per block
