"""Event trace of the gradient all-reduces against backward: run under torchrun with >= 2 ranks; rank 0 profiles one
eager C5 step with torch.profiler (CUPTI kernel records) and prints, for every NCCL all-reduce kernel, when it ran relative to the
step and how much of it was covered by compute kernels on the other stream, plus the time the step spent waiting for the LAST
all-reduce after the last backward kernel had finished (the exposed part).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/ddp_trace.py > profiles/r02_ddp_overlap.txt
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench, tvt_b200
from tvt_b200 import ddp, optim

rank, local, world = ddp.init_from_env()
dev = torch.device("cuda", local)
w = dict(bench.WORKLOADS["c5"]); B = w["batch"]
model = bench.build_model(w, B, "bf16", 0.5, dev)
reducer = ddp.GradBucketReducer([p for p in model.student.parameters()], bucket_bytes=32 << 20, average=False)
opt = optim.FlatOptimizer(reducer, modes=[model.student.mode], kind="adamw", lr=1e-4, weight_decay=0.01)
xs, y = bench.synth_batch(w, B, 1130 + rank)
xs, y = [x.bfloat16().to(dev) for x in xs], y.to(dev)
for _ in range(3):
    bench.gpu_step(w, model, reducer, opt, xs, y)
torch.cuda.synchronize(); dist.barrier()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    bench.gpu_step(w, model, reducer, opt, xs, y)
    torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
    ev.sort(key=lambda e: e.time_range.start)
    t0 = ev[0].time_range.start
    nccl = [e for e in ev if "nccl" in e.name.lower()]
    comp = [e for e in ev if "nccl" not in e.name.lower() and "memcpy" not in e.name.lower() and "memset" not in e.name.lower()]
    step_end = max(e.time_range.end for e in ev)
    optim_ev = [e for e in comp if "optim_kernel" in e.name]
    last_bwd_end = max(e.time_range.end for e in comp if e.time_range.end <= (optim_ev[0].time_range.start if optim_ev else step_end))
    print(f"# C5, {world} x B200, one eager step on rank 0 (torch.profiler CUDA kernel records); times in ms from the first kernel of the step")
    print(f"# step: {(step_end - t0) / 1e3:.3f} ms; {len(comp)} compute kernels, {len(nccl)} NCCL kernels; last backward kernel ends at {(last_bwd_end - t0) / 1e3:.3f} ms")
    tot = cov_tot = 0.0
    for i, e in enumerate(nccl):
        s, t = e.time_range.start, e.time_range.end
        covered = sum(max(0.0, min(t, c.time_range.end) - max(s, c.time_range.start)) for c in comp if c.time_range.end > s and c.time_range.start < t)
        covered = min(covered, t - s)
        tot += t - s; cov_tot += covered
        print(f"allreduce {i:2d}: {(s - t0) / 1e3:8.3f} -> {(t - t0) / 1e3:8.3f} ms  ({(t - s) / 1e3:6.3f} ms, {covered / (t - s) * 100:5.1f} % under compute kernels)  {e.name[:60]}")
    # does a compute kernel run longer when an all-reduce kernel shares the GPU with it?  Kernels are keyed by (name, position in
    # the repeating per-layer pattern is unknown here, so:) name + duration cluster: the reference is the median of the same-name
    # kernels that did NOT overlap an all-reduce and lie within +-35 % of it.
    import json, statistics
    recs = []
    for c in comp:
        s, t = c.time_range.start, c.time_range.end
        ov = sum(max(0.0, min(t, e.time_range.end) - max(s, e.time_range.start)) for e in nccl)
        recs.append({"name": c.name, "start_us": s - t0, "dur_us": t - s, "nccl_overlap_us": ov})
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump({"compute": recs, "nccl": [{"start_us": e.time_range.start - t0, "dur_us": e.time_range.end - e.time_range.start} for e in nccl]},
              open("gpurun_out/ddp_trace_kernels.json", "w"))
    excess = 0.0; n_ov = 0
    for r in recs:
        if r["nccl_overlap_us"] <= 0: continue
        ref = [q["dur_us"] for q in recs if q["name"] == r["name"] and q["nccl_overlap_us"] <= 0 and 0.65 * r["dur_us"] <= q["dur_us"] <= 1.0 * r["dur_us"] + 1e-9]
        if len(ref) >= 2:
            excess += r["dur_us"] - statistics.median(ref); n_ov += 1
    print(f"# compute kernels overlapping an all-reduce: {sum(1 for r in recs if r['nccl_overlap_us'] > 0)}; against the median of same-name non-overlapped kernels "
          f"({n_ov} with a reference) they ran {excess / 1e3:.3f} ms longer in total")
    last_nccl_end = max(e.time_range.end for e in nccl) if nccl else last_bwd_end
    print(f"# all-reduce kernel time {tot / 1e3:.3f} ms, {cov_tot / tot * 100 if tot else 0:.1f} % of it concurrent with compute kernels")
    print(f"# exposed: the last all-reduce ends {(last_nccl_end - last_bwd_end) / 1e3:.3f} ms after the last backward kernel (the optimizer waits for it)")
dist.barrier()
dist.destroy_process_group()
