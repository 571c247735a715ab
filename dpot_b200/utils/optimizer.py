"""B200-native drop-in for the reference's ``utils/optimizer.py``.

``Adam`` / ``AdamW`` keep the reference's constructor arguments, ``param_groups`` and per-parameter
state layout (``step`` python int, ``exp_avg``, ``exp_avg_sq``[, ``max_exp_avg_sq``]) -- see
utils/optimizer.py:55-164 and :217-334 -- so ``optimizer.state_dict()`` round-trips with the
reference.  The update itself (utils/optimizer.py:9-52, :170-212) is ONE fused multi-tensor
kernel of libdpot_b200.so per 32 tensors instead of ~8 elementwise kernels per tensor.
The reference's ``grad * grad.conj()`` only differs from ``grad**2`` for complex parameters;
no parameter of the reference is complex (models/dpot.py:45-48), complex tensors raise here.

``Lamb`` (utils/optimizer.py:359-499, the ``--opt lamb`` branch of the scripts) keeps the reference's
constructor and state layout too; its step is the fused two-stage multi-tensor kernel pair of
csrc/lamb.cu (moments + per-tensor norms, then the trust-ratio update) without host synchronisation.
"""
from __future__ import annotations

import math

import torch
from torch.optim.optimizer import Optimizer

from .. import ops


class _FusedAdamBase(Optimizer):
    _decoupled = False

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if not 0.0 <= lr:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {}".format(eps))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError("Invalid beta parameter at index 0: {}".format(betas[0]))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter at index 1: {}".format(betas[1]))
        if not 0.0 <= weight_decay:
            raise ValueError("Invalid weight_decay value: {}".format(weight_decay))
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad))
        self.grad_scale = 1.0  # set to 1/world_size to fold the DDP average into the step
        # clip_grad_norm_ folded into the step: set by dpot_b200.utils.clip.clip_grad_norm_(..., optimizer=self) --
        # (device double holding sum g^2, max_norm); consumed (and cleared) by the next step()
        self._pending_clip = None

    def __setstate__(self, state):
        super().__setstate__(state)
        for group in self.param_groups:
            group.setdefault('amsgrad', False)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        clip, self._pending_clip = self._pending_clip, None
        for group in self.param_groups:
            ps, gs, ms, vs, xs, steps = [], [], [], [], [], []
            beta1, beta2 = group['betas']
            for p in group['params']:
                if p.grad is None:  # skipped exactly like the reference (:123)
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError('Adam does not support sparse gradients, please consider SparseAdam instead')
                if p.is_complex():
                    raise RuntimeError('dpot_b200 Adam: complex parameters are not supported (the reference has none)')
                if not p.is_cuda or p.dtype != torch.float32:
                    raise RuntimeError('dpot_b200 Adam: parameters must be float32 CUDA tensors (no CPU fallback)')
                st = self.state[p]
                if len(st) == 0:
                    st['step'] = 0
                    st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    if group['amsgrad']:
                        st['max_exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st['step'] += 1
                g = p.grad
                if not (p.is_contiguous() and g.is_contiguous()):
                    raise RuntimeError('dpot_b200 Adam: parameters and gradients must be contiguous')
                ps.append(p); gs.append(g); ms.append(st['exp_avg']); vs.append(st['exp_avg_sq'])
                if group['amsgrad']:
                    xs.append(st['max_exp_avg_sq'])
                steps.append(int(st['step']))
            if not ps:
                continue
            # one process may drive a GPU other than the current one (train_temporal.py --gpu N never calls set_device)
            with torch.cuda.device(ps[0].device):
                ops.adam_step_multi(ps, gs, ms, vs, xs if group['amsgrad'] else None, steps, lr=float(group['lr']),
                                    beta1=beta1, beta2=beta2, eps=group['eps'], weight_decay=group['weight_decay'],
                                    decoupled=self._decoupled, grad_scale=self.grad_scale,
                                    grad_sqnorm=clip[0] if clip else None, max_norm=clip[1] if clip else 0.0)
            # the kernel updates the parameters through raw pointers: bump their autograd version counters so that
            # everything keyed on (data_ptr, _version) -- the inference engine's packed-weight arena, captured rollout
            # graphs, autograd's saved-tensor checks -- sees the change, exactly as p.addcdiv_() would have
            torch.autograd.graph.increment_version(ps)
        return loss


class Adam(_FusedAdamBase):
    """Adam with L2-coupled weight decay, reference utils/optimizer.py:55-164 + adam() :9-52."""
    _decoupled = False


class AdamW(_FusedAdamBase):
    """AdamW (decoupled decay), reference utils/optimizer.py:217-334 + adamw() :170-212."""
    _decoupled = True

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad)


class Lamb(Optimizer):
    """Layer-wise adaptive moments, reference utils/optimizer.py:359-499 (`--opt lamb`, evaluate.py:135-136 constructs it
    with adam=True).  Same constructor, param_groups and per-parameter state (`step`, `exp_avg`, `exp_avg_sq`,
    `weight_norm`, `adam_norm`, `trust_ratio`) as the reference; the step itself is dpot_lamb_step_multi: two launches
    per 56 tensors, no host synchronisation (the reference synchronises twice per parameter: `weight_norm == 0` and the
    tensor-valued alpha of the final add_).  weight_norm / adam_norm / trust_ratio are 0-dim device tensors (views of
    one buffer per group), as the reference's are whenever both norms are non-zero."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0, clamp_value=10, adam=False,
                 debias=False):
        if lr <= 0.0:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if eps < 0.0:
            raise ValueError("Invalid epsilon value: {}".format(eps))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError("Invalid beta parameter at index 0: {}".format(betas[0]))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter at index 1: {}".format(betas[1]))
        if weight_decay < 0:
            raise ValueError("Invalid weight_decay value: {}".format(weight_decay))
        if clamp_value < 0.0:
            raise ValueError("Invalid clamp value: {}".format(clamp_value))
        self.clamp_value, self.adam, self.debias = clamp_value, adam, debias
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            ps, gs, ms, vs, steps = [], [], [], [], []
            for p in group['params']:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError('Lamb does not support sparse gradients, please consider SparseAdam instead')
                if not p.is_cuda or p.dtype != torch.float32:
                    raise RuntimeError('dpot_b200 Lamb: parameters must be float32 CUDA tensors (no CPU fallback)')
                if not (p.is_contiguous() and p.grad.is_contiguous()):
                    raise RuntimeError('dpot_b200 Lamb: parameters and gradients must be contiguous')
                st = self.state[p]
                if len(st) == 0:
                    st['step'] = 0
                    st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st['step'] += 1
                ps.append(p); gs.append(p.grad); ms.append(st['exp_avg']); vs.append(st['exp_avg_sq'])
                steps.append(int(st['step']))
            if not ps:
                continue
            dev, n = ps[0].device, len(ps)
            norms = torch.empty(2 * n, device=dev, dtype=torch.float64)
            info = torch.empty(3 * n, device=dev, dtype=torch.float32)
            beta1, beta2 = group['betas']
            with torch.cuda.device(dev):
                ops.lamb_step_multi(ps, gs, ms, vs, steps, lr=float(group['lr']), beta1=beta1, beta2=beta2, eps=group['eps'],
                                    weight_decay=group['weight_decay'], clamp_value=self.clamp_value, debias=self.debias,
                                    adam=self.adam, norms=norms, info=info)
            for k, p in enumerate(ps):
                st = self.state[p]
                st['weight_norm'], st['adam_norm'], st['trust_ratio'] = info[3 * k], info[3 * k + 1], info[3 * k + 2]
            torch.autograd.graph.increment_version(ps)      # raw-pointer update: see _FusedAdamBase.step
        return loss
