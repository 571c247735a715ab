"""Data-parallel training step on N GPUs of one box (launch with torchrun, one rank per GPU over NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/train_ddp_bench.py [S|M] [batch per GPU] [T_ar]

The reference's only exchange step is the DDP gradient average (train_temporal_parallel.py:185,244).  Measures, with
CUDA events and the max over ranks:
  * the step (forward x T_ar, loss, backward, exchange, clip + Adam) with NO exchange, with one flat all-reduce after
    backward (GradArena) and with bucketed all-reduces overlapped with backward (OverlappedGradArena);
  * the all-reduce of the whole gradient arena alone -> algorithm / bus bandwidth against NVLink;
  * gradient equality: the averaged N-rank gradient == the gradient of the concatenated global batch / N (rank 0
    recomputes it on one GPU).
Prints one JSON line on rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpot_b200 import zoo
from dpot_b200.models.dpot import DPOTNet
from dpot_b200.parallel import FusedGradExchange, GradArena, OverlappedGradArena, init_from_env
from dpot_b200.train import ar_train_step
from dpot_b200.utils.optimizer import Adam

from dpot_b200 import _lib
QUICK = os.environ.get("DDP_QUICK", "0") == "1"          # only the no-exchange / sequential / overlapped step times
BUDGET = int(os.environ.get("DPOT_SM_BUDGET", "0"))       # SM budget of the persistent kernels while overlapping
name = sys.argv[1] if len(sys.argv) > 1 else "S"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
T_ar = int(sys.argv[3]) if len(sys.argv) > 3 else 1
rank, world, dev = init_from_env()
SHAPES = {"L": dict(img_size=256, patch_size=16, modes=64)}          # BASELINE.json config 4
cfg = zoo.zoo_cfg(name, **SHAPES.get(name, {}))
R = cfg["img_size"]
model = zoo.synthetic_weights_(DPOTNet(**cfg), seed=0).to(dev).train()
opt = Adam(model.parameters(), lr=1e-4, betas=(0.9, 0.9), weight_decay=1e-6)
opt.grad_scale = 1.0
g = torch.Generator(device="cpu").manual_seed(100 + rank)
xx = torch.randn((B, R, R, 10, 4), generator=g).to(dev)
yy = torch.randn((B, R, R, T_ar, 4), generator=g).to(dev)
msk = torch.ones((B, R, R, 1, 4), device=dev)
nparam = sum(p.numel() for p in model.parameters())


def timed(fn, n=8, warm=3):
    for _ in range(warm):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


out = {"model": name, "world": world, "batch_per_gpu": B, "T_ar": T_ar, "params": nparam}
state = {"i": 0}


def step_with(arena):
    state["i"] += 1
    return ar_train_step(model, opt, xx, yy, msk, T_bundle=1, noise_scale=0.0, grad_clip=1e4, arena=arena, step=state["i"])


out["fused_training_step"] = bool(model._train_eng is not None and model._train_eng.supported) if hasattr(model, "_train_eng") else False
out["ms_step_no_exchange"] = timed(lambda: step_with(None))
out["fused_training_step"] = bool(model._train_eng is not None and model._train_eng.supported)
if not QUICK:
    flat = GradArena(model.parameters())
    out["ms_step_flat_allreduce"] = timed(lambda: step_with(flat))
if out["fused_training_step"]:
    seq = FusedGradExchange(model, overlap=False)
    out["ms_step_arena_sequential"] = timed(lambda: step_with(seq))
    seq.close()
    over = FusedGradExchange(model, overlap=True)
    out["sm_budget"], out["nccl_max_ctas"] = BUDGET, os.environ.get("NCCL_MAX_CTAS", "")
    _lib.load().dpot_set_sm_budget(BUDGET)
    out["messages"] = len(over.msgs)
else:
    over = OverlappedGradArena(model.parameters(), bucket_mb=32.0)
out["ms_step_overlapped"] = timed(lambda: step_with(over))
_lib.load().dpot_set_sm_budget(0)
# the exchange alone
buf = over.buf if over.buf is not None else torch.zeros(nparam, device=dev)
if world > 1 and not QUICK:
    ms_ar = timed(lambda: dist.all_reduce(buf, op=dist.ReduceOp.AVG), n=20, warm=5)
    nbytes = buf.numel() * 4
    out["allreduce_ms"] = ms_ar
    out["allreduce_bytes"] = nbytes
    out["allreduce_algbw_GBs"] = nbytes / (ms_ar * 1e-3) / 1e9
    out["allreduce_busbw_GBs"] = nbytes * 2 * (world - 1) / world / (ms_ar * 1e-3) / 1e9
over.close()
out["field_steps_per_s_overlapped"] = world * B * T_ar / (out["ms_step_overlapped"] * 1e-3)

# ---- gradient equality: N ranks x B samples averaged == one rank on the concatenated batch, divided by N
if world > 1 and not QUICK:
    Bc = 2
    model2 = zoo.synthetic_weights_(DPOTNet(**cfg), seed=0).to(dev).train()
    gx = torch.Generator(device="cpu").manual_seed(7)
    X = torch.randn((world * Bc, R, R, 10, 4), generator=gx).to(dev)
    Y = torch.randn((world * Bc, R, R, 1, 4), generator=gx).to(dev)
    M = torch.ones((world * Bc, R, R, 1, 4), device=dev)
    from dpot_b200.train import LpLossFn
    im0, _ = model2(X[:1].contiguous())        # builds the training engine
    del im0
    fusedx = model2._train_eng is not None and model2._train_eng.supported
    arena = FusedGradExchange(model2) if fusedx else OverlappedGradArena(model2.parameters(), bucket_mb=32.0)
    if fusedx:
        model2._train_eng.n_forward = 0
    sl = slice(rank * Bc, (rank + 1) * Bc)
    im, _ = model2(X[sl].contiguous())
    LpLossFn.apply(im, Y[sl].contiguous(), M[sl].contiguous()).backward()
    arena.finish()
    arena.close()
    if rank == 0:
        model3 = zoo.synthetic_weights_(DPOTNet(**cfg), seed=0).to(dev).train()
        im, _ = model3(X)
        LpLossFn.apply(im, Y, M).backward()
        worst = 0.0
        for (k, p2), (_, p3) in zip(model2.named_parameters(), model3.named_parameters()):
            if p3.grad is None:
                assert p2.grad is None or float(p2.grad.abs().max()) == 0.0, k
                continue
            ref = p3.grad / world
            e = float((p2.grad - ref).norm() / (ref.norm() + 1e-30))
            worst = max(worst, e)
        out["grad_equality_worst_rel_l2"] = worst
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
