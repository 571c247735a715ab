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


class OverlappedGradArena:
    """The gradient exchange of GradArena, overlapped with backward (reverse-layer buckets on a
    communication stream): parameters are laid out in REVERSE registration order -- the order their gradients become
    ready -- in buckets of ~bucket_mb; a post-accumulate-grad hook copies each fresh gradient into its arena slot,
    makes .grad a view of the arena and, when the last expected gradient of a bucket has arrived, starts that bucket's
    asynchronous all-reduce while autograd keeps running the earlier layers.  finish() (after loss.backward(), before
    clip_grad_norm_ / optimizer.step()) starts whatever is left, waits and leaves DDP-averaged gradients behind.

    The set of parameters that receive gradients is learned on the first step (e.g. cls_head gets none in
    train_temporal.py:226): from then on buckets do not wait for them.  A gradient that shows up after its bucket has
    been launched is exchanged on its own in finish().  All ranks run the same graph, so every rank takes the same
    decisions and issues the same collectives in the same order."""

    def __init__(self, params: Iterable[torch.Tensor], bucket_mb: float = 64.0):
        self.params: List[torch.Tensor] = [p for p in params if p.requires_grad][::-1]
        cap = max(1, int(bucket_mb * (1 << 20) / 4))
        self.slots, self.buckets, o, start, members = {}, [], 0, 0, []
        for p in self.params:
            n = (p.numel() + 63) // 64 * 64
            self.slots[p] = (len(self.buckets), o, p.numel())
            members.append(p)
            o += n
            if o - start >= cap:
                self.buckets.append(dict(start=start, end=o, params=members, expect=len(members), pending=len(members),
                                         handle=None, arrived=[]))
                start, members = o, []
        if members:
            self.buckets.append(dict(start=start, end=o, params=members, expect=len(members), pending=len(members), handle=None,
                                     arrived=[]))
        self.total = o
        self.buf: Optional[torch.Tensor] = None
        self.seen, self.late, self.steps, self.arrived = set(), [], 0, set()
        self.hooks = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params]

    def _world(self) -> int:
        return dist.get_world_size() if dist.is_initialized() else 1

    def _pack(self, b) -> None:
        """Move the bucket's fresh gradients into their arena slots with ONE multi-tensor copy and make .grad views of
        the arena.  (The copy is 8 B per parameter on top of the 28 B/param optimizer step -- S: 123 MB in + out, about
        40 us per step; autograd owns the gradient buffers it hands to the hook, so writing them in place would need
        every backward kernel to target the arena.)"""
        ps = [p for p in b["arrived"] if p.grad is not None and p.grad.data_ptr() != self.buf[self.slots[p][1]:].data_ptr()]
        if ps:
            views = [self.buf[self.slots[p][1]:self.slots[p][1] + self.slots[p][2]].view_as(p) for p in ps]
            torch._foreach_copy_(views, [p.grad for p in ps])
            for p, v in zip(ps, views):
                p.grad = v
        b["arrived"] = []

    def _launch(self, b) -> None:
        if self._world() > 1:
            op = dist.ReduceOp.AVG if dist.get_backend() == "nccl" else dist.ReduceOp.SUM
            b["handle"] = dist.all_reduce(self.buf[b["start"]:b["end"]], op=op, async_op=True)
        else:
            b["handle"] = True

    @torch.no_grad()
    def _hook(self, p: torch.Tensor) -> None:
        if self.buf is None or self.buf.device != p.grad.device:
            self.buf = torch.zeros(self.total, device=p.grad.device, dtype=torch.float32)
        bi, o, n = self.slots[p]
        b = self.buckets[bi]
        self.seen.add(p)
        if b["handle"] is not None:            # its bucket is already on the wire
            if p in self.arrived:
                # a SECOND gradient for a parameter whose bucket is in flight (gradient accumulation, two backward
                # passes before finish()): autograd has just accumulated in place into the buffer NCCL is reducing
                raise RuntimeError("OverlappedGradArena: a parameter received a second gradient before finish(); "
                                   "the contract is one backward pass per finish() (train_temporal_parallel.py:243-245)")
            self.arrived.add(p)
            self.late.append(p)
            return
        self.arrived.add(p)
        b["arrived"].append(p)
        b["pending"] -= 1
        if b["pending"] == 0:
            self._pack(b)
            self._launch(b)

    @torch.no_grad()
    def finish(self) -> int:
        """Start the exchanges that are still pending, wait for all of them; returns the fp32 elements exchanged."""
        world, gloo = self._world(), dist.is_initialized() and dist.get_backend() != "nccl"
        for b in self.buckets:
            if b["handle"] is None and any(p.grad is not None for p in b["params"]):
                b["arrived"] = [p for p in b["params"] if p.grad is not None]
                self._pack(b)
                self._launch(b)
        n = 0
        for b in self.buckets:
            if b["handle"] is None:
                continue
            if b["handle"] is not True:
                b["handle"].wait()
            if world > 1 and gloo:
                self.buf[b["start"]:b["end"]].div_(world)
            n += b["end"] - b["start"]
        if self.late:                               # gradients that arrived after their bucket was on the wire: one exchange
            flat = torch.cat([p.grad.reshape(-1) for p in self.late])
            if world > 1:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM)
                flat.div_(world)
            o = 0
            for p in self.late:
                p.grad.copy_(flat[o:o + p.numel()].view_as(p))
                o += p.numel()
            n += flat.numel()
        # re-arm: from the second step on a bucket only waits for the parameters that really receive gradients
        self.steps += 1
        for b in self.buckets:
            b["expect"] = sum(1 for p in b["params"] if p in self.seen) or len(b["params"])
            b["pending"], b["handle"] = b["expect"], None
        self.late = []
        self.arrived = set()
        return n

    def close(self) -> None:
        for h in self.hooks:
            h.remove()
        self.hooks = []


class FusedGradExchange:
    """The DDP gradient average (train_temporal_parallel.py:185,244) for the one-call training step (train_engine.py):
    dpot_train_backward writes every parameter gradient straight into ONE flat arena laid out in the order groups of
    gradients become final (out_layer | block depth-1 | ... | block 0 | front) and records a CUDA event per group; as
    soon as the call has been enqueued, each group's all-reduce(AVG) is queued on a communication stream behind its
    event, so the exchange of the last layers travels over NVLink while backward still computes the first ones.  No
    packing copy: .grad of every parameter is a view of the arena.  finish() (after loss.backward(), before clip /
    optimizer.step()) makes the compute stream wait for the exchange.

    With several forward calls per optimizer step (T_ar > 1: autograd accumulates the calls' gradients in place into the
    arena views) the exchange cannot start before the last accumulation: finish() then runs one all-reduce of the arena.
    """

    def __init__(self, model, overlap: Optional[bool] = None, group_mb: float = 24.0):
        """overlap=False (default): one all-reduce of the arena after backward.  Measured on 8 x B200 (profiles/r02p_*):
        queuing the messages behind backward's completion events does NOT pay -- the contraction kernels are persistent,
        one CTA per SM with a static tile partition, so every SM NCCL occupies delays a whole share of tiles (DPOT-M,
        16 / GPU: 15.6 ms overlapped vs 15.1 ms sequential vs 13.6 ms without exchange); the exchange itself runs at
        ~620 GB/s bus bandwidth (1.37 ms for 489 MB).  For the two large models the overlap DOES pay (profiles/r02ah_*, 8 per
        GPU: DPOT-L 38.3 -> 36.8 ms, DPOT-H 59.8 -> 57.1 ms), so overlap=None (default) turns it on above 1 GB of gradients."""
        from .train_engine import _TrainEngine
        self.overlap = overlap
        self._auto = overlap is None
        if model._train_eng is None:
            model._train_eng = _TrainEngine(model)
        self.eng = model._train_eng
        self.eng.exchange = self
        if self._auto:
            self.overlap = self.eng.total * 4 > 1e9
        self.buf: Optional[torch.Tensor] = None
        self.events = None
        self.comm = None
        self.handles = []
        self.calls = 0
        self.overlapped = False
        # merge consecutive completion groups into messages of >= group_mb (launch latency vs overlap)
        self.msgs, lo, last = [], None, 0
        for gi, (a, b) in enumerate(self.eng.buckets):
            if lo is None:
                lo = a
            if (b - lo) * 4 >= group_mb * 2 ** 20 or gi == len(self.eng.buckets) - 1:
                self.msgs.append((gi, lo, b))          # (event index that completes the message, float range)
                lo = None

    def _world(self) -> int:
        return dist.get_world_size() if dist.is_initialized() else 1

    def begin(self, eng, dev):
        """Called by FusedTrainFn.backward: (flat gradient buffer, events) for this call, or (None, None)."""
        self.calls += 1
        if self.calls > 1:
            return None, None                          # a later call of the same step: plain buffer, autograd accumulates
        if self.buf is None or self.buf.device != dev:
            self.buf = torch.empty(eng.total, device=dev, dtype=torch.float32)
            self.events = [torch.cuda.Event() for _ in eng.buckets]
            for e in self.events:
                e.record()                             # torch creates the cudaEvent lazily: force it
            self.comm = torch.cuda.Stream(device=dev)
        self.overlapped = self.overlap and eng.n_forward == 1 and self._world() > 1
        return self.buf, (self.events if self.overlapped else None)

    def enqueued(self, eng, flat, events) -> None:
        """dpot_train_backward has been enqueued: queue every message's all-reduce behind its completion event."""
        op = dist.ReduceOp.AVG if dist.get_backend() == "nccl" else dist.ReduceOp.SUM
        with torch.cuda.stream(self.comm):
            for gi, lo, hi in self.msgs:
                self.comm.wait_event(events[gi])
                self.handles.append(dist.all_reduce(flat[lo:hi], op=op, async_op=True))

    @torch.no_grad()
    def finish(self) -> int:
        """Returns the number of fp32 elements exchanged."""
        world = self._world()
        n = 0
        if self.calls and world > 1:
            if self.overlapped:
                for h in self.handles:
                    h.wait()                           # the current stream waits for the message
            else:
                if dist.get_backend() == "nccl":
                    dist.all_reduce(self.buf, op=dist.ReduceOp.AVG)
                else:
                    dist.all_reduce(self.buf, op=dist.ReduceOp.SUM)
            if dist.get_backend() != "nccl":
                self.buf.div_(world)
            n = self.buf.numel()
        self.handles, self.calls, self.overlapped = [], 0, False
        self.eng.n_forward = 0
        return n

    def close(self) -> None:
        self.eng.exchange = None
