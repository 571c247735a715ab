"""The one-call training step: ``im, cls = model(xx)`` / ``loss.backward()`` of train_temporal.py:206,227 as ONE
autograd node whose forward and backward are each a single call into libdpot_b200 (dpot_train_forward /
dpot_train_backward, csrc/train_step.cu).  PyTorch contributes the tape edge and owns the memory; every kernel of the
step -- contractions on the f16-split tcgen05 engine, FFT adjoints, GroupNorm backward, the fold / un-fold of the
re-parameterised weights -- is the library's.  Configurations the C side does not serve (dpot_train_supported == 0)
train through the per-operator path in autograd.py."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import torch
from torch.autograd import Function

from . import _lib
from ._lib import BLOCK_FIELDS, BlockParams, Params, check, ptr


def _param_table(net) -> List[Tuple[str, Optional[int], torch.nn.Parameter]]:
    """(C field, block index or None, parameter) for every parameter the C structs name."""
    pe0, pe2 = net.patch_embed.proj[0], net.patch_embed.proj[2]
    ch, ol = net.cls_head, net.out_layer
    tab = [("pos_embed", None, net.pos_embed), ("pe0_w", None, pe0.weight), ("pe0_b", None, pe0.bias),
           ("pe2_w", None, pe2.weight), ("pe2_b", None, pe2.bias), ("tagg_w", None, net.time_agg_layer.w)]
    if net.time_agg == 'exp_mlp':
        tab.append(("tagg_gamma", None, net.time_agg_layer.gamma))
    tab += [("cls0_w", None, ch[0].weight), ("cls0_b", None, ch[0].bias), ("cls2_w", None, ch[2].weight),
            ("cls2_b", None, ch[2].bias), ("cls4_w", None, ch[4].weight), ("cls4_b", None, ch[4].bias),
            ("out0_w", None, ol[0].weight), ("out0_b", None, ol[0].bias), ("out2_w", None, ol[2].weight),
            ("out2_b", None, ol[2].bias), ("out4_w", None, ol[4].weight), ("out4_b", None, ol[4].bias)]
    for i, blk in enumerate(net.blocks):
        ps = dict(norm1_w=blk.norm1.weight, norm1_b=blk.norm1.bias, w1=blk.filter.w1, b1=blk.filter.b1, w2=blk.filter.w2,
                  b2=blk.filter.b2, norm2_w=blk.norm2.weight, norm2_b=blk.norm2.bias, fc1_w=blk.mlp[0].weight,
                  fc1_b=blk.mlp[0].bias, fc2_w=blk.mlp[2].weight, fc2_b=blk.mlp[2].bias)
        tab += [(f, i, ps[f]) for f in BLOCK_FIELDS]
    return tab


class _TrainEngine:
    """Weights prepared once per optimizer step (dpot_train_prepare, keyed on parameter versions), a scratch arena per
    batch size, and the gradient layout.  `grad_buffer`: optional callable(total_floats, device) -> flat fp32 tensor the
    parameter gradients are written into (parameter order of `self.table`), e.g. a data-parallel exchange arena."""

    def __init__(self, net):
        from .models.dpot import _InferenceEngine
        self.net = net
        self.lib = _lib.load()
        self.binder = _InferenceEngine(net)          # cfg + parameter struct; its inference arena is never packed
        self.cfg = self.binder.cfg
        self.supported = bool(self.lib.dpot_train_supported(C.byref(self.cfg))) and net.mixing_type == 'afno'
        self.table = _param_table(net)
        named = {id(p) for p in net.parameters()}
        self.complete = named == {id(p) for _, _, p in self.table}   # no parameter outside the C structs (normalize=False)
        # flat gradient layout in the order groups become final during backward (dpot_train_backward `events`):
        # out_layer | block depth-1 | ... | block 0 | front (pos_embed, PatchEmbed, time aggregation) + cls head
        def group(entry):
            f, bi, _ = entry
            if bi is not None:
                return 1 + (net.depth - 1 - bi)
            return 0 if f.startswith("out") else net.depth + 1
        order = sorted(range(len(self.table)), key=lambda i: (group(self.table[i]), i))
        self.offsets, o = [0] * len(self.table), 0
        self.buckets = []                             # (lo, hi) float ranges, one per event
        for gi in range(net.depth + 2):
            lo = o
            for i in order:
                if group(self.table[i]) == gi:
                    self.offsets[i] = o
                    o += (self.table[i][2].numel() + 63) // 64 * 64
            self.buckets.append((lo, o))
        self.total = o
        self.exchange = None                          # FusedGradExchange (parallel.py) when training data-parallel
        self.n_forward = 0                            # forwards since the last optimizer step (T_ar > 1: several backward calls)
        self.packed = self.wprep = None
        self.key = None
        self.scratch: Dict[int, torch.Tensor] = {}
        self.grad_buffer = None

    def _stream(self):
        return torch.cuda.current_stream().cuda_stream

    def refresh(self, dev) -> None:
        key = tuple((p.data_ptr(), p._version) for p in self.net.parameters())
        if key == self.key and self.packed is not None and self.packed.device == dev:
            return
        for p in self.net.parameters():
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("dpot_b200: parameters must be contiguous float32 CUDA tensors (call model.cuda())")
        self.binder.bind_params(dev)
        if self.packed is None or self.packed.device != dev:
            self.packed = torch.empty(self.binder.packed_floats, device=dev, dtype=torch.float32)
            self.wprep = torch.empty(self.lib.dpot_train_wprep_floats(C.byref(self.cfg)), device=dev, dtype=torch.float32)
            self.scratch = {}
        check(self.lib.dpot_train_prepare(C.byref(self.cfg), C.byref(self.binder.prm), ptr(self.packed), ptr(self.wprep),
                                          ptr(self.get_scratch(1, dev)), self._stream()), "dpot_train_prepare")
        self.key = key
        self.n_forward = 0                            # new weights = new optimizer step

    def get_scratch(self, B: int, dev) -> torch.Tensor:
        big = max(self.scratch) if self.scratch else 0
        if big >= B and self.scratch[big].device == dev:
            return self.scratch[big]                 # the layout of a smaller batch fits the larger arena
        n = self.lib.dpot_train_scratch_floats(C.byref(self.cfg), B)
        if n <= 0:
            raise RuntimeError("dpot_train_scratch_floats: " + self.lib.dpot_last_error_string().decode())
        self.scratch = {B: torch.empty(n, device=dev, dtype=torch.float32)}
        return self.scratch[B]

    def grads_struct(self, flat: torch.Tensor, want_cls: bool):
        blocks = (BlockParams * self.net.depth)()
        g = Params()
        views = []
        base = flat.data_ptr()
        for (f, bi, p), o in zip(self.table, self.offsets):
            if f.startswith("cls") and not want_cls:
                views.append(None)
                continue
            addr = C.cast(base + 4 * o, _lib.c_f32p)
            if bi is None:
                setattr(g, f, addr)
            else:
                setattr(blocks[bi], f, addr)
            views.append(flat[o:o + p.numel()].view(p.shape))
        g.blocks = C.cast(blocks, C.POINTER(BlockParams))
        return g, blocks, views


class FusedTrainFn(Function):
    @staticmethod
    def forward(ctx, eng: _TrainEngine, x: torch.Tensor, *params):
        net, lib = eng.net, eng.lib
        dev = x.device
        ctx.set_materialize_grads(False)
        with torch.cuda.device(dev):
            eng.refresh(dev)
            B = x.shape[0]
            tape = torch.empty(lib.dpot_train_tape_floats(C.byref(eng.cfg), B), device=dev, dtype=torch.float32)
            y = torch.empty((B, net.img_size, net.img_size, net.out_timesteps, net.out_channels), device=dev, dtype=torch.float32)
            cls = torch.empty((B, net.n_cls), device=dev, dtype=torch.float32)
            check(lib.dpot_train_forward(C.byref(eng.cfg), C.byref(eng.binder.prm), ptr(eng.packed), ptr(eng.wprep), ptr(x), B, ptr(y), ptr(cls),
                                         ptr(tape), ptr(eng.get_scratch(B, dev)), eng._stream()), "dpot_train_forward")
        eng.n_forward += 1
        ctx.eng = eng
        ctx.save_for_backward(x, tape)
        ctx.need_dx = x.requires_grad
        ctx.key = eng.key
        return y, cls

    @staticmethod
    def backward(ctx, dy, dcls):
        eng: _TrainEngine = ctx.eng
        lib = eng.lib
        x, tape = ctx.saved_tensors
        dev = x.device
        if ctx.key != eng.key:
            raise RuntimeError("dpot_b200: parameters changed between the training forward and its backward")
        with torch.cuda.device(dev):
            B = x.shape[0]
            if dy is None:
                dy = torch.zeros((B, eng.net.img_size, eng.net.img_size, eng.net.out_timesteps, eng.net.out_channels), device=dev)
            dy = dy.contiguous().float()
            dcls = dcls.contiguous().float() if dcls is not None else None
            ex = eng.exchange
            flat, events = (ex.begin(eng, dev) if ex is not None else (None, None))
            if flat is None:
                flat = eng.grad_buffer(eng.total, dev) if eng.grad_buffer is not None else torch.empty(eng.total, device=dev)
            g, keep, views = eng.grads_struct(flat, dcls is not None)
            dx = torch.empty_like(x) if ctx.need_dx else None
            ev = None
            if events is not None:
                ev = (C.c_void_p * len(events))(*[e.cuda_event for e in events])
            check(lib.dpot_train_backward(C.byref(eng.cfg), C.byref(eng.binder.prm), ptr(eng.packed), ptr(eng.wprep), ptr(x), B,
                                          ptr(dy), ptr(dcls), ptr(tape), ptr(eng.get_scratch(B, dev)), C.byref(g), ptr(dx),
                                          ev, eng._stream()), "dpot_train_backward")
            del keep
            if events is not None:
                ex.enqueued(eng, flat, events)
        by_id = {id(p): v for (_, _, p), v in zip(eng.table, views)}
        return (None, dx) + tuple(by_id.get(id(p)) if p.requires_grad else None for p in eng.net.parameters())


def fused_train_forward(net, x: torch.Tensor):
    """(im, cls) through the one-call training step, or None when this configuration / input is not served."""
    if getattr(net, "train_path", "auto") == "generic" or net.mixing_type != 'afno':
        return None
    eng = net._train_eng
    if eng is None:
        eng = net._train_eng = _TrainEngine(net)
    if not (eng.supported and eng.complete):
        return None
    if x.dtype != torch.float32:
        x = x.float()
    x = x.contiguous()
    return FusedTrainFn.apply(eng, x, *net.parameters())
