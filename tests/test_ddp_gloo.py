"""CPU, world_size 2 over gloo: the bucketed gradient reducer gives every rank the gradients a single
process computes on the concatenated batch (SURVEY.md section 8e parity check), including parameters
whose hooks never fire."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_model():
    torch.manual_seed(1130)
    return torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 32), torch.nn.ReLU(),
                               torch.nn.Linear(32, 5))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tvt_b200.ddp import GradBucketReducer
    model = _make_model()
    unused = torch.nn.Parameter(torch.ones(7))                       # never touched by backward
    red = GradBucketReducer(list(model.parameters()) + [unused], bucket_bytes=2048)
    assert len(red.buckets) >= 2
    g = torch.Generator().manual_seed(7)
    x = torch.randn(8, 16, generator=g)
    y = torch.randn(8, 5, generator=g)
    xs, ys = x[rank * 4:(rank + 1) * 4], y[rank * 4:(rank + 1) * 4]
    for step in range(2):                                            # second step checks zero_grad re-arming
        red.zero_grad()
        torch.nn.functional.mse_loss(model(xs), ys).backward()
        red.finish()
    out[rank] = [p.grad.clone() for p in model.parameters()] + [unused.grad.clone()]
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_bucketed_allreduce_matches_single_process():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    model = _make_model()
    g = torch.Generator().manual_seed(7)
    x = torch.randn(8, 16, generator=g)
    y = torch.randn(8, 5, generator=g)
    torch.nn.functional.mse_loss(model(x), y).backward()              # mean over the global batch
    ref = [p.grad for p in model.parameters()]
    for rank in range(world):
        got = out[rank]
        for a, b in zip(got[:-1], ref):
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-7)
        assert float(got[-1].abs().max()) == 0.0


class _DirectLinear(torch.autograd.Function):
    """CPU stand-in for the tvt Functions' gradient path: y = x W^T with the weight gradient ACCUMULATED straight into
    the reducer's bucket view (ddp.direct_target) and ``done()`` reported, exactly as functions._Sink does on the GPU."""

    @staticmethod
    def forward(ctx, x, w):
        from tvt_b200 import ddp
        if ctx.needs_input_grad[1]:
            ddp.note_use(w)
        ctx.w = w
        ctx.save_for_backward(x)
        return x @ w.t()

    @staticmethod
    def backward(ctx, dy):
        from tvt_b200 import ddp
        (x,) = ctx.saved_tensors
        w = ctx.w
        tgt = ddp.direct_target(w)
        assert tgt is not None
        tgt[0].add_(dy.t() @ x)
        tgt[1]()
        return dy @ w.detach(), None


def _shared_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tvt_b200.ddp import GradBucketReducer
    torch.manual_seed(1130)
    w = torch.nn.Parameter(torch.randn(16, 16) * 0.3)       # ONE weight applied three times (ptn_shared's shared encoder)
    head = torch.nn.Parameter(torch.randn(3, 16) * 0.3)
    red = GradBucketReducer([head, w], bucket_bytes=64, direct=True)
    fired_at = []
    orig = red._on_grad
    red._on_grad = lambda p: (fired_at.append((id(p) == id(w), red._uses.get(id(p), 0))), orig(p))[1]
    g = torch.Generator().manual_seed(7)
    x = torch.randn(8, 16, generator=g)[rank * 4:(rank + 1) * 4]
    for _ in range(2):
        fired_at.clear()
        red.zero_grad()
        h = x
        for _ in range(3):
            h = torch.tanh(_DirectLinear.apply(h, w))
        assert red._uses[id(w)] == 3
        _DirectLinear.apply(h, head).pow(2).mean().backward()
        assert [u for is_w, u in fired_at if is_w] == [0], fired_at      # armed once, after the LAST contribution
        red.finish()
    out[rank] = [w.grad.clone(), head.grad.clone()]
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_shared_parameter_direct_sinks_world2():
    """ADVICE r1 / VERDICT weak #4: a parameter that receives several direct-sink contributions per backward must arm its
    bucket's all-reduce only after the last one; the averaged gradients equal a single process on the concatenated batch."""
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_shared_worker, args=(world, port, out), nprocs=world, join=True)
    torch.manual_seed(1130)
    w = torch.nn.Parameter(torch.randn(16, 16) * 0.3)
    head = torch.nn.Parameter(torch.randn(3, 16) * 0.3)
    g = torch.Generator().manual_seed(7)
    h = torch.randn(8, 16, generator=g)
    for _ in range(3):
        h = torch.tanh(h @ w.t())
    (h @ head.t()).pow(2).mean().backward()
    for rank in range(world):
        assert torch.allclose(out[rank][0], w.grad, rtol=1e-5, atol=1e-7)
        assert torch.allclose(out[rank][1], head.grad, rtol=1e-5, atol=1e-7)


def _dry_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      TVT_DDP_DRY_RUN="1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tvt_b200.ddp import GradBucketReducer
    model = _make_model()
    red = GradBucketReducer(list(model.parameters()), bucket_bytes=2048, average=False)
    assert red.dry_run
    g = torch.Generator().manual_seed(7)
    x, y = torch.randn(8, 16, generator=g), torch.randn(8, 5, generator=g)
    red.zero_grad()
    torch.nn.functional.mse_loss(model(x[rank * 4:(rank + 1) * 4]), y[rank * 4:(rank + 1) * 4]).backward()
    red.finish()
    assert all(b["handle"] is None for b in red.buckets)            # no collective was issued
    out[rank] = [p.grad.clone() for p in model.parameters()]
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_dry_run_knob_skips_the_allreduce():
    """TVT_DDP_DRY_RUN=1 (the measurement knob behind DESIGN section 6's communication-cost number): gradients stay local."""
    world, port = 2, _free_port()
    out = mp.Manager().dict()
    mp.spawn(_dry_worker, args=(world, port, out), nprocs=world, join=True)
    g = torch.Generator().manual_seed(7)
    x, y = torch.randn(8, 16, generator=g), torch.randn(8, 5, generator=g)
    for rank in range(world):
        model = _make_model()
        torch.nn.functional.mse_loss(model(x[rank * 4:(rank + 1) * 4]), y[rank * 4:(rank + 1) * 4]).backward()
        for a, p in zip(out[rank], model.parameters()):
            assert torch.allclose(a, p.grad, rtol=1e-6, atol=1e-8)   # each rank's OWN shard gradient, unreduced
