"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: batch sharding, flat gradient arena
all-reduce == DDP average, max-over-ranks timing."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from dpot_b200 import parallel
    r, w, dev = parallel.init_from_env(backend="gloo")
    assert (r, w) == (rank, world) and dev.type == "cpu"
    torch.manual_seed(0)                                   # same parameters on every rank
    ps = [torch.nn.Parameter(torch.randn(s)) for s in [(3, 5), (7,), (2, 2, 2), (130,)]]
    frozen = torch.nn.Parameter(torch.randn(4), requires_grad=False)
    nograd = torch.nn.Parameter(torch.randn(6))            # receives no gradient (like cls_head in train_temporal.py)
    full = torch.arange(8 * 3, dtype=torch.float32).reshape(8, 3)
    mine = parallel.shard_batch(full, rank, world)
    assert mine.shape[0] == 4 and mine[0, 0].item() == rank * 12
    for i, p in enumerate(ps):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    arena = parallel.GradArena(ps + [frozen, nograd])
    n = arena.allreduce()
    assert n >= sum(p.numel() for p in ps)
    ok = True
    for i, p in enumerate(ps):
        want = (1 + 2) / 2 * (i + 1)                        # DDP average over 2 ranks
        ok &= bool(torch.allclose(p.grad, torch.full_like(p, want)))
        ok &= p.grad.data_ptr() >= arena.buf.data_ptr()      # grads are views of the arena
    ok &= nograd.grad is None
    tmax = parallel.max_over_ranks(10.0 + rank, dev)
    ok &= tmax == 11.0
    parallel.broadcast_parameters([ps[0]])
    q.put((rank, ok))
    dist.destroy_process_group()


def test_two_rank_gloo_gradient_average_and_timing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def test_shard_batch_rejects_ragged():
    from dpot_b200 import parallel
    with pytest.raises(ValueError):
        parallel.shard_batch(torch.zeros(5, 2), 0, 2)


def _overlap_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from dpot_b200 import parallel
    parallel.init_from_env(backend="gloo")
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 40), torch.nn.Tanh(), torch.nn.Linear(40, 40), torch.nn.Tanh(),
                              torch.nn.Linear(40, 3))
    unused = torch.nn.Parameter(torch.randn(50))            # never receives a gradient (cls_head in train_temporal.py)
    params = list(net.parameters()) + [unused]
    arena = parallel.OverlappedGradArena(params, bucket_mb=200 * 4 / (1 << 20))       # ~one Linear per bucket
    assert len(arena.buckets) >= 3
    ok = True
    for step in range(3):
        g = torch.Generator().manual_seed(100 + 10 * step + rank)       # different data on every rank
        x, y = torch.randn(16, 6, generator=g), torch.randn(16, 3, generator=g)
        for p in params:
            p.grad = None                                                # optimizer.zero_grad(set_to_none=True)
        loss = ((net(x) - y) ** 2).sum()
        loss.backward()
        arena.finish()
        got = [p.grad.clone() for p in net.parameters()]
        # reference: plain local gradients, averaged by hand
        ref_net = torch.nn.Sequential(torch.nn.Linear(6, 40), torch.nn.Tanh(), torch.nn.Linear(40, 40), torch.nn.Tanh(),
                                      torch.nn.Linear(40, 3))
        ref_net.load_state_dict(net.state_dict())
        ((ref_net(x) - y) ** 2).sum().backward()
        for gp, rp in zip(got, ref_net.parameters()):
            t = rp.grad.clone()
            dist.all_reduce(t)
            ok &= bool(torch.allclose(gp, t / world, rtol=1e-6, atol=1e-6))
        ok &= unused.grad is None
        ok &= all(p.grad.data_ptr() >= arena.buf.data_ptr() for p in net.parameters())
        with torch.no_grad():
            for p in net.parameters():
                p -= 0.01 * p.grad
    arena.close()
    q.put((rank, ok))
    dist.destroy_process_group()


def test_two_rank_gloo_overlapped_bucketed_allreduce():
    """OverlappedGradArena: bucketed asynchronous all-reduce launched from autograd hooks == DDP average, over several
    steps, with a parameter that never receives a gradient."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_overlap_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]
