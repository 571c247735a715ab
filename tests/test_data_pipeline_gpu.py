"""dpot_assemble_batch / dpot_b200.data.BatchAssembler against the reference's own data path restated with the same torch
calls: utils/griddataset.py:88-100 (pad_data: F.interpolate bilinear + channel padding with 1.0), :152-157 (training
window), :102-116 (get_target_mask)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def pad_data(x, res, n_channels):                  # utils/griddataset.py:88-100, line for line
    H, W, T, C = x.shape
    x = x.view(H, W, -1).permute(2, 0, 1)
    x = F.interpolate(x.unsqueeze(0), size=(res, res), mode='bilinear').squeeze(0).permute(1, 2, 0)
    x = x.view(*x.shape[:2], T, C)
    x_new = torch.ones([*x.shape[:-1], n_channels])
    x_new[..., :x.shape[-1]] = x
    return x_new


def target_mask(x, size_orig):                     # utils/griddataset.py:102-116
    msk = torch.zeros(*x.shape[:2], 1, x.shape[-1])
    kx, ky = x.shape[0] // size_orig[0], x.shape[1] // size_orig[1]
    kx = 1 if kx == 0 else kx
    ky = 1 if ky == 0 else ky
    msk[::kx, ::ky, :, :size_orig[-1]] = 1
    return msk


@pytest.mark.parametrize("H0,res,C0", [(64, 128, 1), (128, 128, 3), (96, 64, 2), (50, 128, 4)])
def test_training_batch_matches_reference_data_path(H0, res, C0):
    from dpot_b200.data import BatchAssembler
    rng = np.random.default_rng(H0 + res)
    B, T0, t_in, t_ar, C = 3, 14, 10, 2, 4
    raw = rng.standard_normal((B, H0, H0, T0, C0)).astype(np.float32)
    starts = [0, 1, 2]
    asm = BatchAssembler(res, t_in, t_ar, C)
    xx, yy, msk = asm(torch.from_numpy(raw).pin_memory(), starts)
    for b in range(B):
        s = pad_data(torch.from_numpy(raw[b]), res, C)
        x_ref, y_ref = s[..., starts[b]:starts[b] + t_in, :], s[..., starts[b] + t_in:starts[b] + t_in + t_ar, :]
        assert torch.allclose(xx[b].cpu(), x_ref, rtol=0, atol=2e-6), (b, float((xx[b].cpu() - x_ref).abs().max()))
        assert torch.allclose(yy[b].cpu(), y_ref, rtol=0, atol=2e-6)
    assert bool((msk == 1).all())
    # second call: the double-buffered staging path, random windows stay inside the sample
    xx2, yy2, _ = asm(torch.from_numpy(raw).pin_memory())
    assert torch.isfinite(xx2).all() and torch.isfinite(yy2).all()


def test_evaluation_mask_matches_get_target_mask():
    from dpot_b200.data import BatchAssembler
    rng = np.random.default_rng(1)
    raw = rng.standard_normal((2, 64, 64, 12, 3)).astype(np.float32)
    asm = BatchAssembler(128, 10, 2, 4, train=False, pred_channels=2)
    xx, yy, msk = asm(raw)
    s = pad_data(torch.from_numpy(raw[1]), 128, 4)
    assert torch.allclose(xx[1].cpu(), s[..., :10, :], rtol=0, atol=2e-6)
    assert torch.equal(msk[1].cpu(), target_mask(s, [64, 64, 12, 2]))
