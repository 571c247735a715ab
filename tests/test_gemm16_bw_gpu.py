"""The backward-pass forms of the f16-split tcgen05 engine (dpot_gemm_args ABI 2: a_trans / w_trans = MN-major operand
tiles straight from the forward layouts, k_split, act' multiply and fp32 pre-activation with a split result) against
float64 numpy on the same fp32 inputs.  These are the contractions of autograd's backward of models/dpot.py:72-94,
157-161 (data gradient dx = g W, weight gradient dW = g^T x)."""
import numpy as np
import pytest
import torch

from oracle import dpot_oracle as O

pytestmark = pytest.mark.gpu
TOL = 3e-6


def _rand(shape, seed, scale=1.0):
    return (np.random.default_rng(seed).standard_normal(shape) * scale).astype(np.float32)


def _dev(a):
    return torch.from_numpy(a).cuda()


def _act_grad(x, name):
    from scipy.special import erf
    x = x.astype(np.float64)
    if name == "gelu":
        return 0.5 * (1 + erf(x / np.sqrt(2))) + x * np.exp(-0.5 * x * x) / np.sqrt(2 * np.pi)
    if name == "relu":
        return (x > 0).astype(np.float64)
    raise KeyError(name)


@pytest.mark.parametrize("M,N,K", [(4096, 1024, 1024), (512, 256, 320), (200, 136, 72), (4096, 352, 1024)])
@pytest.mark.parametrize("pair", [-1, 0, 1])
def test_data_gradient_w_trans(M, N, K, pair):
    """dx[M, N] = g[M, K] @ W[K, N] with W in its forward layout [K(out), N(in)] (w_trans)."""
    from dpot_b200 import _lib, ops
    g, W = _rand((M, K), 1), _rand((K, N), 2, K ** -0.5)
    ref = g.astype(np.float64) @ W.astype(np.float64)
    _lib.load().dpot_tc16_set_pair(pair)
    try:
        out = ops.gemm16_bw(ops.split_f16(_dev(g)), ops.split_f16(_dev(W)), M, N, K, w_trans=True)
    finally:
        _lib.load().dpot_tc16_set_pair(-1)
    assert O.rel_l2(out.cpu().numpy(), ref) < TOL


@pytest.mark.parametrize("Mtok,N,K", [(4096, 1024, 1024), (2304, 256, 512), (1000, 136, 72)])
@pytest.mark.parametrize("pair", [-1, 0, 1])
def test_weight_gradient_both_trans(Mtok, N, K, pair):
    """dW[N, K] = g[Mtok, N]^T @ x[Mtok, K] from the token-major tensors as stored (a_trans + w_trans), chunked."""
    from dpot_b200 import _lib, ops
    g, x = _rand((Mtok, N), 3), _rand((Mtok, K), 4)
    ref = g.astype(np.float64).T @ x.astype(np.float64)
    g16, x16 = ops.split_f16(_dev(g)), ops.split_f16(_dev(x))
    _lib.load().dpot_tc16_set_pair(pair)
    try:
        one = ops.gemm16_bw(g16, x16, N, K, Mtok, a_trans=True, w_trans=True)[0]
        parts = ops.gemm16_bw(g16, x16, N, K, Mtok, a_trans=True, w_trans=True, k_chunk=1024)
    finally:
        _lib.load().dpot_tc16_set_pair(-1)
    assert O.rel_l2(one.cpu().numpy(), ref) < 2 * TOL
    tot = parts.double().sum(0) if parts.dim() == 4 else parts.double()
    assert O.rel_l2(tot[0].cpu().numpy(), ref) < TOL


def test_block_diagonal_backward_forms():
    """The AFNO block MLP (models/dpot.py:72-94 in real block form): nb independent [2bs x 2bs] problems side by side."""
    from dpot_b200 import ops
    nb, n, Ms = 8, 256, 2304
    g, x = _rand((Ms, nb * n), 5), _rand((Ms, nb * n), 6)
    Wc = _rand((nb, n, n), 7, n ** -0.5)
    pre = _rand((Ms, nb * n), 8)
    g16, x16, W16 = ops.split_f16(_dev(g)), ops.split_f16(_dev(x)), ops.split_f16(_dev(Wc.reshape(nb * n, n)))
    # data gradient with the activation derivative of the previous layer, result split
    dx16 = ops.gemm16_bw(g16, W16, Ms, n, n, w_trans=True, nb=nb, out16=True, dact_src=_dev(pre), dact="gelu")
    ref = np.einsum("mbn,bnk->mbk", g.astype(np.float64).reshape(Ms, nb, n), Wc.astype(np.float64)).reshape(Ms, nb * n)
    ref = ref * _act_grad(pre, "gelu")
    assert O.rel_l2(ops.unsplit_f16(dx16).cpu().numpy(), ref) < TOL
    # weight gradient, contraction over the spectral rows in chunks
    parts = ops.gemm16_bw(g16, x16, n, n, Ms, a_trans=True, w_trans=True, nb=nb, k_chunk=768)
    refw = np.einsum("mbn,mbk->bnk", g.astype(np.float64).reshape(Ms, nb, n), x.astype(np.float64).reshape(Ms, nb, n))
    assert O.rel_l2(parts.double().sum(0).cpu().numpy(), refw) < TOL


def test_forward_with_preactivation_and_split_result():
    """Training forward: act(x W^T + b) stored split for the next contraction + the fp32 pre-activation for backward."""
    from dpot_b200 import ops
    M, N, K = 1024, 512, 256
    x, W, b = _rand((M, K), 9), _rand((N, K), 10, K ** -0.5), _rand((N,), 11)
    out16, pre = ops.gemm16_bw(ops.split_f16(_dev(x)), ops.split_f16(_dev(W)), M, N, K, out16=True, pre=True, act="gelu",
                               bias=_dev(b))
    refp = x.astype(np.float64) @ W.astype(np.float64).T + b
    assert O.rel_l2(pre.cpu().numpy(), refp) < TOL
    assert O.rel_l2(ops.unsplit_f16(out16).cpu().numpy(), O.activation(refp, "gelu")) < TOL


@pytest.mark.parametrize("K", [4096, 6144, 8192, 2112])
def test_long_contraction_runs_as_chained_launches(K):
    """fc2 of DPOT-M / L / H (models/dpot.py:160) is 4096 / 6144 / 8192 deep.  One tcgen05 accumulation chain of that
    length is 1.2e-9 * K off (truncating accumulator); dpot_gemm chains <= 2048-deep launches through the result buffer.
    Checked against fp64: plain fp32 result with bias + residual + GroupNorm statistics (chained in place), activation
    result (partial sums must enter the activation), split-fp16 result (dpot_gemm_chained with scratch), and the
    unchained launch for contrast."""
    import dpot_b200
    from dpot_b200 import ops
    M, N, rps = 1024, 512, 256
    A = np.maximum(_rand((M, K), 3), 0) + 0.1 * _rand((M, K), 4)
    W, b, R = _rand((N, K), 5, K ** -0.5), _rand((N,), 6), _rand((M, N), 7)
    pre = A.astype(np.float64) @ W.astype(np.float64).T + b
    A16, W16 = ops.split_f16(_dev(A)), ops.split_f16(_dev(W))
    rel = lambda x, ref: float(np.linalg.norm(x.double().cpu().numpy() - ref) / np.linalg.norm(ref))
    assert dpot_b200.set_chain(2048) == 2048      # the default
    out, st = ops.gemm16(A16, W16, bias=_dev(b), residual=_dev(R), stats=(8, rps))
    e_plain = rel(out, pre + R)
    ref_st = (pre + R).reshape(M // rps, rps, 8, N // 8)
    e_st = max(rel(st[..., 0], ref_st.sum((1, 3))), rel(st[..., 1], (ref_st ** 2).sum((1, 3))))
    e_act = rel(ops.gemm16(A16, W16, bias=_dev(b), act="gelu", residual=_dev(R)), O.activation(pre, "gelu") + R)
    e_16 = rel(ops.unsplit_f16(ops.gemm16(A16, W16, bias=_dev(b), residual=_dev(R), out16=True)), pre + R)
    try:
        dpot_b200.set_chain(0)
        e_one = rel(ops.gemm16(A16, W16, bias=_dev(b), residual=_dev(R)), pre + R)
    finally:
        dpot_b200.set_chain(2048)
    print(f"K={K}: chained {e_plain:.2e} (stats {e_st:.2e}, gelu {e_act:.2e}, split result {e_16:.2e}); one launch {e_one:.2e}")
    assert max(e_plain, e_act, e_16) < TOL and e_st < 2e-6      # statistics: fp32 partial sums per warp, double across tiles
    if K >= 4096:
        assert e_one > 1.5 * e_plain      # the reason the chained form exists
