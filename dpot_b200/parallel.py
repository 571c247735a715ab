"""Data parallelism over the GPUs of one box: one process per GPU, NCCL over NVLink/NVSwitch.

The reference's only parallel strategy is DDP through HF accelerate (train_temporal_parallel.py:102,185,244):
every rank holds a full replica, the batch is sharded by rank, and the gradients of a SUM-over-local-batch
loss are averaged over ranks once per optimizer step.  That single exchange is reproduced here as ONE
all-reduce(AVG) over a flat fp32 gradient arena (S: 123 MB, H: 4.1 GB -> a handful of large NVLS-eligible
messages instead of ~90 small ones).  The inference rollout needs no exchange at all: ranks are independent
replicas (bench.py --gpus N).
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> tuple:
    """Join the process group described by RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun contract).
    Returns (rank, world, device)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    use_cuda = torch.cuda.is_available() and backend != "gloo"
    dev = torch.device("cuda", local) if use_cuda else torch.device("cpu")
    if use_cuda:
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kw = {"device_id": dev} if use_cuda else {}
        dist.init_process_group(backend or ("nccl" if use_cuda else "gloo"), rank=rank, world_size=world, **kw)
    return rank, world, dev


def shard_batch(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Contiguous per-rank slice of the leading (sample) axis; samples are independent end to end."""
    n = t.shape[0]
    if n % world != 0:
        raise ValueError(f"batch {n} is not divisible by world size {world}")
    per = n // world
    return t[rank * per:(rank + 1) * per]


def max_over_ranks(value: float, device) -> float:
    """Timing rule of bench.py: a multi-GPU number is the MAX over ranks of the device-side time."""
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def broadcast_parameters(params: Iterable[torch.Tensor], src: int = 0) -> None:
    if dist.is_initialized() and dist.get_world_size() > 1:
        for p in params:
            dist.broadcast(p.data, src=src)


class GradArena:
    """Flat gradient arena: pack -> one all-reduce(AVG) -> parameters' .grad become views of the arena.

    DDP semantics of the reference: grad_i <- (1/world) * sum_ranks grad_i, for every parameter that
    received a gradient on this rank (train_temporal_parallel.py:243-244; parameters with grad None are
    skipped exactly like utils/optimizer.py:123 skips them -- all ranks run the same graph, so the set is
    identical on every rank)."""

    def __init__(self, params: Iterable[torch.Tensor]):
        self.params: List[torch.Tensor] = [p for p in params if p.requires_grad]
        self.buf: Optional[torch.Tensor] = None

    def _layout(self, live):
        offs, o = [], 0
        for p in live:
            offs.append(o)
            o += (p.numel() + 63) // 64 * 64          # 256 B granules keep every view 16 B aligned
        return offs, o

    @torch.no_grad()
    def allreduce(self) -> int:
        """Average the gradients over ranks; returns the number of fp32 elements exchanged."""
        live = [p for p in self.params if p.grad is not None]
        if not live:
            return 0
        offs, total = self._layout(live)
        dev = live[0].grad.device
        if self.buf is None or self.buf.numel() != total or self.buf.device != dev:
            self.buf = torch.zeros(total, device=dev, dtype=torch.float32)
        views = [self.buf[o:o + p.numel()].view_as(p) for o, p in zip(offs, live)]
        torch._foreach_copy_(views, [p.grad for p in live])
        world = dist.get_world_size() if dist.is_initialized() else 1
        if world > 1:
            if dist.get_backend() == "nccl":
                dist.all_reduce(self.buf, op=dist.ReduceOp.AVG)
            else:                                          # gloo has no AVG
                dist.all_reduce(self.buf, op=dist.ReduceOp.SUM)
                self.buf.div_(world)
        for p, v in zip(live, views):
            p.grad = v
        return total
