"""Worker of tests/test_ddp_nccl_gpu.py (launched with torch.distributed.run, one rank per GPU, NCCL).

Each rank runs the tvt kernels (direct gradient sinks into the flat buckets, bucket all-reduces launched from inside
backward) on its shard of a global batch, for two models: the C2-style cross-attention fusion transformer with the
pyramid head, and SimpleTransformer.ptn_shared (ONE encoder applied to every expert: several contributions per parameter).
Rank 0 then recomputes the gradients in a single process by accumulating the same shards one after the other (clips are
independent and every loss is a mean, so this is the gradient of the concatenated batch) and writes the comparison."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def targets(B, C, gen):
    y = (torch.rand(B, C, generator=gen) < 0.15).float()
    y[torch.arange(B), torch.randint(0, C, (B,), generator=gen)] = 1.0
    return y


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def main():
    out_path, precision = sys.argv[1], sys.argv[2]
    import tvt_b200  # noqa: F401
    from tvt_b200 import ddp, hostapi
    from tvt_b200.functions import DistillLossFn
    rank, local, world = ddp.init_from_env("nccl")
    dev = torch.device("cuda", local)
    Bs = 16                                    # clips per rank
    gen = torch.Generator().manual_seed(1130)
    result = {"world": world, "precision": precision, "cases": {}}

    def zero_dropout(m):
        for s in m.modules():
            if isinstance(s, torch.nn.Dropout):
                s.p = 0.0
        return m

    # ---- case 1: cross-attention fusion + pyramid
    kw = dict(in_dims=(256, 128), d=128, nhead=2, nhid=256, nlayers=2, dropout=0.0, batch_size=Bs, frames=12, n_classes=15,
              fusion="cross", pyramid=True, precision=precision)
    xs_all = [torch.randn(world * Bs, 12, D, generator=gen) for D in kw["in_dims"]]
    y_all = targets(world * Bs, 15, gen)

    def fusion_loss(mod, xs, y):
        logits, _, ploss = mod(xs, y)
        return DistillLossFn.apply(logits, None, y, 1.0, 0.0, 0.0, 1.0)[0] + ploss[0]

    # ---- case 2: ptn_shared
    cfg = dict(batch_size=Bs, seq_len=12, cls=1, dropout=0.0, input_dimension=128, nhead=2, nhid=256, nlayers=2,
               model="ptn_shared", learning_rate=1e-3, momentum=0.0, weight_decay=0.0, n_classes=15, precision=precision)
    x2_all = torch.randn(world * Bs, 12, 3, 128, generator=gen)
    y2_all = targets(world * Bs, 15, gen)

    cases = (("fusion_cross_pyramid", lambda: zero_dropout(hostapi.FusionTransformer(**kw)), fusion_loss,
              lambda r: ([x[r * Bs:(r + 1) * Bs].to(dev) for x in xs_all], y_all[r * Bs:(r + 1) * Bs].to(dev)), ()),
             ("ptn_shared", lambda: hostapi.SimpleTransformer(**cfg), lambda m, x, y: m._loss(m.ptn_shared(x), y),
              lambda r: (x2_all[r * Bs:(r + 1) * Bs].to(dev), y2_all[r * Bs:(r + 1) * Bs].to(dev)), ("mlp_encoder", "encoder_layers")))
    for name, build, loss_fn, shard, skip in cases:
        torch.manual_seed(1130)
        mod = build().to(dev).train()
        named = [(n, p) for n, p in mod.named_parameters() if not n.startswith(skip)]
        red = ddp.GradBucketReducer([p for _, p in named], bucket_bytes=1 << 16, direct=True)
        launched_in_backward = 0
        for _ in range(2):                          # the second step checks zero_grad() re-arming
            red.zero_grad()
            loss = loss_fn(mod, *shard(rank))
            loss.backward()
            launched_in_backward = sum(b["handle"] is not None for b in red.buckets)
            red.finish()
        got = {n: p.grad.detach().clone() for n, p in named}
        red.remove()
        torch.cuda.synchronize()
        # every rank must now hold bit-identical averaged gradients
        cs = torch.stack([g.double().sum() for g in got.values()])
        lo, hi = cs.clone(), cs.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        identical = bool(torch.equal(lo, hi))
        if rank == 0:
            torch.manual_seed(1130)
            ref = build().to(dev).train()
            ref.load_state_dict(mod.state_dict())
            for r in range(world):                  # accumulate the shards: gradient of the concatenated batch
                (loss_fn(ref, *shard(r)) / world).backward()
            errs = {n: rel(got[n], p.grad) for n, p in ref.named_parameters() if n in got and p.grad is not None}
            worst = max(errs, key=errs.get)
            result["cases"][name] = {"params": len(errs), "worst": errs[worst], "worst_param": worst, "buckets": len(red.buckets),
                                     "allreduces_launched_inside_backward": launched_in_backward, "ranks_identical": identical}
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump(result, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
