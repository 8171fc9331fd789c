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
