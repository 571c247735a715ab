"""B200-native drop-in for the reference's ``models/dpot.py`` hot path.

Same public names, constructor signatures, parameter names/shapes/initialisation order and
``forward`` contracts as the reference (``models/dpot.py:22-110`` AFNO2D, ``:137-180`` Block,
``:183-209`` PatchEmbed, ``:213-234`` TimeAggregator, ``:245-420`` DPOTNet), so that
``train_temporal.py:118`` / ``evaluate.py:124`` construct it unchanged and reference checkpoints
load with ``strict=True``.  The arithmetic does not run in PyTorch: the modules are parameter
containers and every forward enqueues kernels of ``libdpot_b200.so`` (C ABI, sm_100a).  There
is no CPU and no eager-PyTorch fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
import logging
import math
import os
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib, ops
from .._lib import ACT_IDS, GEMM_AUTO, BlockParams, Config, Params, check, ptr

_logger = logging.getLogger(__name__)

# activation modules are parameter-free placeholders (the kernels apply them in their epilogues);
# keys as in models/dpot.py:19
ACTIVATION = {'gelu': nn.GELU(), 'tanh': nn.Tanh(), 'sigmoid': nn.Sigmoid(), 'relu': nn.ReLU(),
              'leaky_relu': nn.LeakyReLU(0.1), 'softplus': nn.Softplus(), 'ELU': nn.ELU(), 'silu': nn.SiLU()}

_SUPPORTED_LATENT = (2, 4, 8, 16, 32)


def _kept_modes(modes: int, h: int) -> Tuple[int, int]:
    """[:modes, :modes] slicing of the [h, h/2+1] spectrum clamps (models/dpot.py:70-94)."""
    return min(modes, h), min(modes, h // 2 + 1)


def _require_cuda(x: torch.Tensor, who: str) -> None:
    if not x.is_cuda:
        raise RuntimeError(f"{who}: dpot_b200 runs on a CUDA (sm_100a) device only; got a {x.device} tensor. "
                           "There is no CPU fallback for the hot path.")


def _grad_needed(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


class AFNO2D(nn.Module):
    """Fourier mixer, reference models/dpot.py:22-110.  Parameters: w1,b1,w2,b2 with the real/imag
    pair in dim 0.  ``sparsity_threshold`` is accepted and ignored: soft-shrink is commented
    out in the reference (:97-98)."""

    def __init__(self, width=32, num_blocks=8, channel_first=False, sparsity_threshold=0.01, modes=32,
                 hard_thresholding_fraction=1, hidden_size_factor=1, act='gelu'):
        super().__init__()
        if width % num_blocks != 0:
            raise AssertionError(f"hidden_size {width} should be divisble by num_blocks {num_blocks}")
        if hidden_size_factor != 1:
            raise NotImplementedError("dpot_b200.AFNO2D: hidden_size_factor != 1 is never used by the reference "
                                      "(models/dpot.py:149-150) and is not built")
        self.hidden_size = width
        self.sparsity_threshold = sparsity_threshold
        self.num_blocks = num_blocks
        self.block_size = width // num_blocks
        self.channel_first = channel_first
        self.modes = modes
        self.hidden_size_factor = hidden_size_factor
        self.scale = 1 / (self.block_size * self.block_size * hidden_size_factor)
        self.act_name = act
        self.act = ACTIVATION[act]
        bs, nb = self.block_size, num_blocks
        # same RNG consumption order as the reference (:45-48)
        self.w1 = nn.Parameter(self.scale * torch.rand(2, nb, bs, bs * hidden_size_factor))
        self.b1 = nn.Parameter(self.scale * torch.rand(2, nb, bs * hidden_size_factor))
        self.w2 = nn.Parameter(self.scale * torch.rand(2, nb, bs * hidden_size_factor, bs))
        self.b2 = nn.Parameter(self.scale * torch.rand(2, nb, bs))

    def spectral_branch(self, a: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, B: int, h: int,
                        want_stats: bool = False, engine: int = GEMM_AUTO):
        """a[B*h*h, E] token-major, (scale, shift)[B,E] the affine applied on load.  Returns
        f = irfft2(MLP(rfft2(a*scale+shift))) + (a*scale+shift) and, optionally, GroupNorm stats of f."""
        km1, km2 = _kept_modes(self.modes, h)
        nb = self.num_blocks
        Wc1, bc1 = ops.pack_afno(self.w1.detach(), self.b1.detach())
        Wc2, bc2 = ops.pack_afno(self.w2.detach(), self.b2.detach())
        S = ops.afno_fft_fwd(a, scale, shift, B, h, nb, km1, km2)
        O1 = ops.gemm_batched_cols(S, Wc1, bc1, nb, act=self.act_name, engine=engine)
        O2 = ops.gemm_batched_cols(O1, Wc2, bc2, nb, act=None, out=S, engine=engine)
        return ops.afno_fft_inv(O2, a, scale, shift, B, h, nb, km1, km2, want_stats=want_stats)

    def forward(self, x, spatial_size=None):
        _require_cuda(x, "AFNO2D.forward")
        if _grad_needed(x, self.w1, self.b1, self.w2, self.b2):
            from ..autograd import afno2d_train
            return afno2d_train(self, x)
        if self.channel_first:
            B, Cc, H, W = x.shape
            a = x.permute(0, 2, 3, 1)
        else:
            B, H, W, Cc = x.shape
            a = x
        if H != W or H not in _SUPPORTED_LATENT:
            raise RuntimeError(f"AFNO2D: latent grid {H}x{W} unsupported (square power of two in [2,32])")
        a = a.reshape(B * H * W, Cc).contiguous().float()
        one = torch.ones((B, Cc), device=x.device)
        f, _ = self.spectral_branch(a, one, torch.zeros_like(one), B, H)
        f = f.reshape(B, H, W, Cc)
        return f.permute(0, 3, 1, 2) if self.channel_first else f


class Mlp(nn.Module):
    """Token MLP, reference models/dpot.py:120-134 (unused by DPOTNet; kept for the namespace)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act='gelu', drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act_name = act
        self.act = ACTIVATION[act]
        self.fc2 = nn.Linear(hidden_features, out_features)

    def forward(self, x):
        _require_cuda(x, "Mlp.forward")
        if _grad_needed(x, *self.parameters()):
            from ..autograd import mlp_train
            return mlp_train(self, x)
        lead = x.shape[:-1]
        a = x.reshape(-1, x.shape[-1]).contiguous().float()
        hdn = ops.gemm(a, self.fc1.weight.detach(), bias=self.fc1.bias.detach(), act=self.act_name)
        out = ops.gemm(hdn, self.fc2.weight.detach(), bias=self.fc2.bias.detach())
        return out.reshape(*lead, -1)


class Block(nn.Module):
    """GroupNorm -> AFNO2D -> GroupNorm -> 1x1-conv MLP -> + residual, reference models/dpot.py:137-180."""

    def __init__(self, mixing_type='afno', double_skip=True, width=32, n_blocks=4, mlp_ratio=1., channel_first=True,
                 modes=32, drop=0., drop_path=0., act='gelu', h=14, w=8):
        super().__init__()
        if mixing_type != 'afno':
            raise NotImplementedError(f"dpot_b200.Block: mixing_type={mixing_type!r}; only 'afno' exists in the reference")
        self.norm1 = nn.GroupNorm(8, width)
        self.width = width
        self.modes = modes
        self.act_name = act
        self.act = ACTIVATION[act]
        self.filter = AFNO2D(width=width, num_blocks=n_blocks, sparsity_threshold=0.01, channel_first=channel_first,
                             modes=modes, hard_thresholding_fraction=1, hidden_size_factor=1, act=act)
        self.norm2 = nn.GroupNorm(8, width)
        hidden = int(width * mlp_ratio)
        self.mlp = nn.Sequential(
            nn.Conv2d(in_channels=width, out_channels=hidden, kernel_size=1, stride=1),
            self.act,
            nn.Conv2d(in_channels=hidden, out_channels=width, kernel_size=1, stride=1),
        )
        self.double_skip = double_skip

    def forward_tokens(self, a: torch.Tensor, B: int, h: int, engine: int = GEMM_AUTO) -> torch.Tensor:
        """a[B*h*h, E] token-major -> block output, same layout (inference)."""
        n = h * h
        E = self.width
        st1 = ops.gn_stats(a, B, n)
        sc1, sh1 = ops.gn_finalize(st1, self.norm1.weight.detach(), self.norm1.bias.detach(), n, self.norm1.eps)
        f, st2 = self.filter.spectral_branch(a, sc1, sh1, B, h, want_stats=not self.double_skip, engine=engine)
        res = a
        if self.double_skip:  # :171-173
            f = f + a
            res = f
            st2 = ops.gn_stats(f, B, n)
        sc2, sh2 = ops.gn_finalize(st2, self.norm2.weight.detach(), self.norm2.bias.detach(), n, self.norm2.eps)
        fc1, fc2 = self.mlp[0], self.mlp[2]
        hdn = ops.gemm(f, fc1.weight.detach().reshape(fc1.out_channels, E), bias=fc1.bias.detach(), act=self.act_name,
                       a_scale=sc2, a_shift=sh2, a_rows_per_sample=n, engine=engine)
        return ops.gemm(hdn, fc2.weight.detach().reshape(E, fc1.out_channels), bias=fc2.bias.detach(), residual=res,
                        engine=engine)

    def forward(self, x):
        _require_cuda(x, "Block.forward")
        if _grad_needed(x, *self.parameters()):
            from ..autograd import block_train
            return block_train(self, x)
        B, E, H, W = x.shape
        if H != W or H not in _SUPPORTED_LATENT:
            raise RuntimeError(f"Block: latent grid {H}x{W} unsupported (square power of two in [2,32])")
        a = x.permute(0, 2, 3, 1).reshape(B * H * W, E).contiguous().float()
        out = self.forward_tokens(a, B, H)
        return out.reshape(B, H, W, E).permute(0, 3, 1, 2)


class PatchEmbed(nn.Module):
    """Conv(k=s=P) -> act -> Conv 1x1, reference models/dpot.py:183-209."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, out_dim=128, act='gelu'):
        super().__init__()
        self.img_size = (img_size, img_size)
        self.patch_size = (patch_size, patch_size)
        self.out_size = (img_size // patch_size, img_size // patch_size)
        self.num_patches = self.out_size[0] * self.out_size[1]
        self.out_dim = out_dim
        self.act_name = act
        self.act = ACTIVATION[act]
        self.proj = nn.Sequential(
            nn.Conv2d(in_chans, embed_dim, kernel_size=self.patch_size, stride=self.patch_size),
            self.act,
            nn.Conv2d(embed_dim, out_dim, kernel_size=1, stride=1),
        )

    def forward(self, x):
        _require_cuda(x, "PatchEmbed.forward")
        if _grad_needed(x, *self.parameters()):
            from ..autograd import patch_embed_train
            return patch_embed_train(self, x)
        B, Cc, H, W = x.shape
        assert H == self.img_size[0] and W == self.img_size[1], \
            f"Input image size ({H}*{W}) doesn't match model ({self.img_size[0]}*{self.img_size[1]})."
        P = self.patch_size[0]
        c0, c2 = self.proj[0], self.proj[2]
        # view NCHW frames as the field layout [N, X, Y, T=1, C] the im2col GEMM reads
        xf = x.permute(0, 2, 3, 1).reshape(B, H, W, 1, Cc).contiguous().float()
        mid = c0.out_channels
        W0p = c0.weight.detach().permute(0, 2, 3, 1).reshape(mid, P * P * Cc).contiguous()
        h, w = H // P, W // P
        rb = c0.bias.detach().reshape(1, mid).expand(h * w, mid).contiguous()
        z1 = ops.patch_gemm(xf, W0p, rb, P, self.act_name, mid)
        out = ops.gemm(z1, c2.weight.detach().reshape(c2.out_channels, mid), bias=c2.bias.detach())
        return out.reshape(B, h, w, -1).permute(0, 3, 1, 2)


class TimeAggregator(nn.Module):
    """Temporal aggregation, reference models/dpot.py:213-234 ('mlp' / 'exp_mlp')."""

    def __init__(self, n_channels, n_timesteps, out_channels, type='mlp'):
        super().__init__()
        self.n_channels = n_channels
        self.n_timesteps = n_timesteps
        self.out_channels = out_channels
        self.type = type
        if type in ('mlp', 'exp_mlp'):
            self.w = nn.Parameter(1 / (n_timesteps * out_channels ** 0.5) *
                                  torch.randn(n_timesteps, out_channels, out_channels), requires_grad=True)
        if type == 'exp_mlp':
            self.gamma = nn.Parameter(2 ** torch.linspace(-10, 10, out_channels).unsqueeze(0), requires_grad=True)

    def time_embedding(self, T: int, device) -> torch.Tensor:
        """t_embed[T,E] = cos(linspace(0,1,T)[:,None] @ gamma) (:230-231); ones for 'mlp'.  K=1 matmul == product."""
        E = self.out_channels
        if self.type != 'exp_mlp':
            return torch.ones((T, E), device=device)
        t = torch.linspace(0, 1, T).unsqueeze(-1).to(device)
        return torch.cos(t * self.gamma.detach())

    def forward(self, x):
        _require_cuda(x, "TimeAggregator.forward")
        if _grad_needed(x, *self.parameters()):
            from ..autograd import time_agg_train
            return time_agg_train(self, x)
        lead, T, E = x.shape[:-2], x.shape[-2], x.shape[-1]
        temb = self.time_embedding(T, x.device)
        # out[..., j] = sum_{t,i} w[t,i,j] x[..., t, i] temb[t,i]  ==  A[M, T*E] @ Wt[E, T*E]^T
        Wt = (self.w.detach() * temb.unsqueeze(-1)).reshape(T * E, E).t().contiguous()
        a = x.reshape(-1, T * E).contiguous().float()
        return ops.gemm(a, Wt).reshape(*lead, E)


class DPOTNet(nn.Module):
    """Reference models/dpot.py:245-420.  forward(x[B,X,Y,T,C]) -> (y[B,X,Y,T_out,C_out], cls_pred[B,n_cls])."""

    def __init__(self, img_size=224, patch_size=16, mixing_type='afno', in_channels=1, out_channels=4, in_timesteps=1,
                 out_timesteps=1, n_blocks=4, embed_dim=768, out_layer_dim=32, depth=12, modes=32, mlp_ratio=1.,
                 n_cls=12, normalize=False, act='gelu', time_agg='exp_mlp'):
        super().__init__()
        if act not in ACT_IDS:
            raise KeyError(act)
        if time_agg not in ('mlp', 'exp_mlp'):
            raise NotImplementedError(f"time_agg={time_agg!r}: the reference defines 'mlp' and 'exp_mlp' only")
        h = img_size // patch_size
        if img_size % patch_size != 0 or h not in _SUPPORTED_LATENT:
            raise NotImplementedError(
                f"dpot_b200.DPOTNet: latent grid img_size/patch_size = {img_size}/{patch_size} is not built "
                f"(needs a power of two in {_SUPPORTED_LATENT}); refusing to fall back to PyTorch")
        if embed_dim % 8 != 0:
            raise ValueError("embed_dim must be divisible by the 8 GroupNorm groups")
        if out_channels * out_timesteps > 16:
            raise NotImplementedError("out_channels*out_timesteps > 16 is not built")
        if normalize and in_channels != out_channels:
            raise ValueError("normalize=True needs in_channels == out_channels (x*sigma+mu, models/dpot.py:401)")
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.in_timesteps = in_timesteps
        self.out_timesteps = out_timesteps
        self.n_blocks = n_blocks
        self.modes = modes
        self.num_features = self.embed_dim = embed_dim
        self.mlp_ratio = mlp_ratio
        self.act_name = act
        self.act = ACTIVATION[act]
        self.img_size = img_size
        self.patch_size = patch_size
        self.out_layer_dim = out_layer_dim
        self.depth = depth
        # construction order == reference order (:278-325) so a seeded init draws identical weights
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_channels + 3,
                                      embed_dim=out_channels * patch_size + 3, out_dim=embed_dim, act=act)
        self.latent_size = self.patch_embed.out_size
        self.pos_embed = nn.Parameter(torch.zeros(1, embed_dim, h, h))
        self.normalize = normalize
        self.time_agg = time_agg
        self.n_cls = n_cls
        self.blocks = nn.ModuleList([
            Block(mixing_type=mixing_type, modes=modes, width=embed_dim, mlp_ratio=mlp_ratio, channel_first=True,
                  n_blocks=n_blocks, double_skip=False, h=h, w=h // 2 + 1, act=act)
            for _ in range(depth)])
        if self.normalize:
            self.scale_feats_mu = nn.Linear(2 * in_channels, embed_dim)
            self.scale_feats_sigma = nn.Linear(2 * in_channels, embed_dim)
        self.cls_head = nn.Sequential(nn.Linear(embed_dim, embed_dim), self.act, nn.Linear(embed_dim, embed_dim),
                                      self.act, nn.Linear(embed_dim, n_cls))
        self.time_agg_layer = TimeAggregator(in_channels, in_timesteps, embed_dim, time_agg)
        self.out_layer = nn.Sequential(
            nn.ConvTranspose2d(in_channels=embed_dim, out_channels=out_layer_dim, kernel_size=patch_size,
                               stride=patch_size),
            self.act,
            nn.Conv2d(in_channels=out_layer_dim, out_channels=out_layer_dim, kernel_size=1, stride=1),
            self.act,
            nn.Conv2d(in_channels=out_layer_dim, out_channels=out_channels * out_timesteps, kernel_size=1, stride=1),
        )
        torch.nn.init.trunc_normal_(self.pos_embed, std=.02)
        self.mixing_type = mixing_type
        # engine state (not part of the state dict)
        self.gemm_engine = GEMM_AUTO
        self._eng: Optional[_InferenceEngine] = None
        self._train_eng = None
        self.train_path = os.environ.get('DPOT_TRAIN_PATH', 'auto')   # 'auto' | 'generic' (per-operator autograd only)

    # -- reference helpers kept for API compatibility ------------------------------------------
    def _init_weights(self, m):  # defined but never applied in the reference (:329-337)
        if isinstance(m, (nn.Linear, nn.Conv2d)):
            torch.nn.init.trunc_normal_(m.weight, std=.002)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def get_grid(self, x):
        B, X, Y = x.shape[0], x.shape[1], x.shape[2]
        gx = torch.tensor(np.linspace(0, 1, X), dtype=torch.float).reshape(1, X, 1, 1).repeat([B, 1, Y, 1])
        gy = torch.tensor(np.linspace(0, 1, Y), dtype=torch.float).reshape(1, 1, Y, 1).repeat([B, X, 1, 1])
        return torch.cat((gx, gy), dim=-1).to(x.device)

    def get_grid_3d(self, x):
        B, X, Y, Z = x.shape[0], x.shape[1], x.shape[2], x.shape[3]
        gx = torch.tensor(np.linspace(0, 1, X), dtype=torch.float).reshape(1, X, 1, 1, 1).to(x.device)
        gy = torch.tensor(np.linspace(0, 1, Y), dtype=torch.float).reshape(1, 1, Y, 1, 1).to(x.device)
        gz = torch.tensor(np.linspace(0, 1, Z), dtype=torch.float).reshape(1, 1, 1, Z, 1).to(x.device)
        return torch.cat((gx.expand(B, X, Y, Z, 1), gy.expand(B, X, Y, Z, 1), gz.expand(B, X, Y, Z, 1)), dim=-1)

    # -- forward ---------------------------------------------------------------------------------
    def forward(self, x):
        _require_cuda(x, "DPOTNet.forward")
        if x.dim() != 5:
            raise ValueError("DPOTNet.forward expects x[B, X, Y, T, C]")
        B, X, Y, T, Cc = x.shape
        assert X == self.img_size and Y == self.img_size, \
            f"Input image size ({X}*{Y}) doesn't match model ({self.img_size}*{self.img_size})."
        if T != self.in_timesteps or Cc != self.in_channels:
            raise ValueError(f"expected T={self.in_timesteps}, C={self.in_channels}; got T={T}, C={Cc}")
        if _grad_needed(x, *self.parameters()):
            from ..train_engine import fused_train_forward
            out = fused_train_forward(self, x)         # one-call training step (dpot_train_*) when the geometry is served
            if out is not None:
                return out
            from ..autograd import dpot_forward_train  # generic per-operator path
            return dpot_forward_train(self, x)
        return self.engine().forward(x)

    def engine(self) -> "_InferenceEngine":
        if self._eng is None:
            self._eng = _InferenceEngine(self)
        return self._eng

    def extra_repr(self) -> str:
        mods = {name for name, _ in self.named_modules()}
        out = ''
        for name, p in self.named_parameters():
            head = name.split('.')[0]
            if head not in mods:
                out += '(' + head + '): tensor(' + str(tuple(p.shape)) + ', requires_grad=' + str(p.requires_grad) + ')\n'
        return out

    def _apply(self, fn, *a, **k):
        self._eng = None  # parameter storage may move (.to(device))
        self._train_eng = None
        return super()._apply(fn, *a, **k)

    # the engine caches raw pointers: never pickle / deepcopy it
    def __getstate__(self):
        st = self.__dict__.copy()
        st['_eng'] = None
        st['_train_eng'] = None
        return st


class _InferenceEngine:
    """Owns what dpot_forward() needs besides the parameters: the C config/param structs, the
    packed-weight arena (re-derived when any parameter's version or storage changes) and a
    per-batch-size activation workspace.  All device memory is torch-allocated and only lent to
    the library for the duration of the enqueued work."""

    def __init__(self, net: DPOTNet):
        self.net = net
        self.lib = _lib.load()
        c = Config()
        c.img_size, c.patch_size = net.img_size, net.patch_size
        c.in_channels, c.out_channels = net.in_channels, net.out_channels
        c.in_timesteps, c.out_timesteps = net.in_timesteps, net.out_timesteps
        c.n_blocks, c.embed_dim, c.out_layer_dim, c.depth = net.n_blocks, net.embed_dim, net.out_layer_dim, net.depth
        c.modes, c.hidden_dim, c.n_cls = net.modes, int(net.embed_dim * net.mlp_ratio), net.n_cls
        c.normalize, c.act, c.time_agg = int(net.normalize), ACT_IDS[net.act_name], int(net.time_agg == 'exp_mlp')
        self.cfg = c
        self.packed_floats = self.lib.dpot_packed_floats(C.byref(c))
        if self.packed_floats <= 0:
            raise RuntimeError("dpot_b200: unsupported configuration: " + self.lib.dpot_last_error_string().decode())
        self.packed: Optional[torch.Tensor] = None
        self.key = None
        self.blocks_arr = (BlockParams * net.depth)()
        self.prm = Params()
        self.ws: Dict[int, torch.Tensor] = {}
        self.aux = None  # grid tables / temb keep-alive

    def _param_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.net.parameters())

    def refresh(self, device) -> None:
        key = self._param_key()
        if key == self.key and self.packed is not None and self.packed.device == device:
            return
        self.bind_params(device)
        if self.packed is None or self.packed.device != device:
            self.packed = torch.empty(self.packed_floats, device=device, dtype=torch.float32)
        check(self.lib.dpot_pack_weights(C.byref(self.cfg), C.byref(self.prm), ptr(self.packed),
                                         torch.cuda.current_stream().cuda_stream), "dpot_pack_weights")
        self.key = key

    def bind_params(self, device) -> None:
        """Point the C parameter struct at the current parameter storage (+ the coordinate / time-embedding tables)."""
        net = self.net
        for p in net.parameters():
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("dpot_b200: parameters must be contiguous float32 CUDA tensors (call model.cuda())")
        prm = self.prm
        pe0, pe2 = net.patch_embed.proj[0], net.patch_embed.proj[2]
        prm.pos_embed, prm.pe0_w, prm.pe0_b = ptr(net.pos_embed), ptr(pe0.weight), ptr(pe0.bias)
        prm.pe2_w, prm.pe2_b = ptr(pe2.weight), ptr(pe2.bias)
        prm.tagg_w = ptr(net.time_agg_layer.w)
        prm.tagg_gamma = ptr(net.time_agg_layer.gamma) if net.time_agg == 'exp_mlp' else None
        ch = net.cls_head
        prm.cls0_w, prm.cls0_b, prm.cls2_w, prm.cls2_b = ptr(ch[0].weight), ptr(ch[0].bias), ptr(ch[2].weight), ptr(ch[2].bias)
        prm.cls4_w, prm.cls4_b = ptr(ch[4].weight), ptr(ch[4].bias)
        ol = net.out_layer
        prm.out0_w, prm.out0_b, prm.out2_w, prm.out2_b = ptr(ol[0].weight), ptr(ol[0].bias), ptr(ol[2].weight), ptr(ol[2].bias)
        prm.out4_w, prm.out4_b = ptr(ol[4].weight), ptr(ol[4].bias)
        if net.normalize:
            prm.mu_w, prm.mu_b = ptr(net.scale_feats_mu.weight), ptr(net.scale_feats_mu.bias)
            prm.sigma_w, prm.sigma_b = ptr(net.scale_feats_sigma.weight), ptr(net.scale_feats_sigma.bias)
        for i, blk in enumerate(net.blocks):
            b = self.blocks_arr[i]
            b.norm1_w, b.norm1_b = ptr(blk.norm1.weight), ptr(blk.norm1.bias)
            b.w1, b.b1, b.w2, b.b2 = ptr(blk.filter.w1), ptr(blk.filter.b1), ptr(blk.filter.w2), ptr(blk.filter.b2)
            b.norm2_w, b.norm2_b = ptr(blk.norm2.weight), ptr(blk.norm2.bias)
            b.fc1_w, b.fc1_b = ptr(blk.mlp[0].weight), ptr(blk.mlp[0].bias)
            b.fc2_w, b.fc2_b = ptr(blk.mlp[2].weight), ptr(blk.mlp[2].bias)
        prm.blocks = C.cast(self.blocks_arr, C.POINTER(BlockParams))
        # coordinate tables exactly as get_grid_3d builds them (np.linspace float64 -> float32, :350-357)
        R, T = net.img_size, net.in_timesteps
        gx = torch.tensor(np.linspace(0, 1, R), dtype=torch.float).to(device)
        gt = torch.tensor(np.linspace(0, 1, T), dtype=torch.float).to(device)
        temb = net.time_agg_layer.time_embedding(T, device).contiguous()
        self.aux = (gx, gt, temb)
        prm.grid_x, prm.grid_y, prm.grid_t, prm.temb = ptr(gx), ptr(gx), ptr(gt), ptr(temb)

    def workspace(self, B: int, device) -> torch.Tensor:
        ws = self.ws.get(B)
        if ws is None or ws.device != device:
            nfl = self.lib.dpot_workspace_floats(C.byref(self.cfg), B)
            if nfl <= 0:
                raise RuntimeError("dpot_workspace_floats: " + self.lib.dpot_last_error_string().decode())
            ws = torch.empty(nfl, device=device, dtype=torch.float32)
            # a few batch sizes stay resident (train batch, eval batch, last partial batch); the oldest goes first.
            # Captured CUDA graphs never point here: RolloutEngine owns its workspace (new_workspace).
            self.ws = {k: v for k, v in list(self.ws.items())[-2:] if v.device == device}
            self.ws[B] = ws
        return ws

    def new_workspace(self, B: int, device) -> torch.Tensor:
        """A workspace the caller owns (RolloutEngine: its pointer is baked into a captured graph)."""
        nfl = self.lib.dpot_workspace_floats(C.byref(self.cfg), B)
        if nfl <= 0:
            raise RuntimeError("dpot_workspace_floats: " + self.lib.dpot_last_error_string().decode())
        return torch.empty(nfl, device=device, dtype=torch.float32)

    @torch.no_grad()
    def forward(self, x: torch.Tensor, out: Optional[torch.Tensor] = None, want_cls: bool = True, t0: int = 0):
        """t0: x is a ring in time whose logical frame t is slot (t + t0) % T (0 = the plain layout)."""
        net = self.net
        if x.dtype != torch.float32:
            x = x.float()
        x = x.contiguous()
        dev = x.device
        with torch.cuda.device(dev):
            self.refresh(dev)
            B = x.shape[0]
            ws = self.workspace(B, dev)
            if out is None:
                out = torch.empty((B, net.img_size, net.img_size, net.out_timesteps, net.out_channels), device=dev,
                                  dtype=torch.float32)
            cls = torch.empty((B, net.n_cls), device=dev, dtype=torch.float32) if want_cls else None
            check(self.lib.dpot_forward_ring(C.byref(self.cfg), C.byref(self.prm), ptr(self.packed), ptr(x), t0, B,
                                             ptr(out), ptr(cls), ptr(ws), net.gemm_engine,
                                             torch.cuda.current_stream().cuda_stream), "dpot_forward_ring")
        return out, cls

    @torch.no_grad()
    def rollout_step(self, ring: torch.Tensor, scratch: torch.Tensor, pred: Optional[torch.Tensor], t0: int, step: int,
                     ws: Optional[torch.Tensor] = None, want_cls: Optional[torch.Tensor] = None):
        """One autoregressive step on the ring window (dpot_rollout_step): the new frames overwrite the oldest slots
        (t0 + j) % T of `ring` and land in pred[..., step*T_out + j, :]; the caller advances t0 by T_out."""
        dev = ring.device
        with torch.cuda.device(dev):
            self.refresh(dev)
            B = ring.shape[0]
            if ws is None:
                ws = self.workspace(B, dev)
            check(self.lib.dpot_rollout_step(C.byref(self.cfg), C.byref(self.prm), ptr(self.packed), ptr(ring), t0, B,
                                             ptr(scratch), ptr(want_cls), ptr(ws), self.net.gemm_engine, ptr(pred),
                                             pred.shape[-2] if pred is not None else 0, step,
                                             torch.cuda.current_stream().cuda_stream), "dpot_rollout_step")


def resize_pos_embed(posemb, posemb_new):
    """Bilinear rescale of a ViT-style [1, 1+g*g, D] position table (reference models/dpot.py:424-441;
    checkpoint-loading helper, off the hot path -> plain torch)."""
    _logger.info('Resized position embedding: %s to %s', posemb.shape, posemb_new.shape)
    ntok_new = posemb_new.shape[1] - 1
    tok, grid = posemb[:, :1], posemb[0, 1:]
    gs_old, gs_new = int(math.sqrt(len(grid))), int(math.sqrt(ntok_new))
    _logger.info('Position embedding grid-size from %s to %s', gs_old, gs_new)
    grid = grid.reshape(1, gs_old, gs_old, -1).permute(0, 3, 1, 2)
    grid = F.interpolate(grid, size=(gs_new, gs_new), mode='bilinear')
    grid = grid.permute(0, 2, 3, 1).reshape(1, gs_new * gs_new, -1)
    return torch.cat([tok, grid], dim=1)


def checkpoint_filter_fn(state_dict, model):
    """Checkpoint key/shape fix-ups (reference models/dpot.py:444-459)."""
    if 'model' in state_dict:
        state_dict = state_dict['model']
    fixed = {}
    for k, v in state_dict.items():
        if 'patch_embed.proj.weight' in k and len(v.shape) < 4:
            O, I, H, W = model.patch_embed.proj.weight.shape
            v = v.reshape(O, -1, H, W)
        elif k == 'pos_embed' and v.shape != model.pos_embed.shape:
            v = resize_pos_embed(v, model.pos_embed)
        fixed[k] = v
    return fixed
