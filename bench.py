#!/usr/bin/env python
"""Headline benchmark: train clips/sec (forward + backward + optimizer step) of the temporal video
transformer hot path on N B200s, clip-batch data parallel (weak scaling).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5|c4|c3|c2|c1] [--impl reference]

Prints ONE JSON line (rank 0).  `value` = whole-job clips/s with the step's inputs already resident in
HBM; `e2e` = the same step driven from pinned HOST batches (host->device copy of every step's inputs and
a device->host read of the loss inside the timed region, copies prefetched on a side stream).
`roofline` describes the dominant kernel (the tcgen05 GEMM) from a CUDA-event-instrumented step run after
the timed region; `cpu_baseline` times the CPU oracle (oracle/param.py, the reference's torch.nn
composition) on a bounded sample of the same workload on this box's host cores.
`--impl reference` times that CPU oracle alone and prints the same line shape with "impl": "reference".
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

# BASELINE.json configs: per-GPU clips, frames, width, layers, heads, ff
WORKLOADS = {
    "c1": dict(desc="C1 single-modality 4-layer d=512 encoder, 16 frames x 2048-d RGB, batch 8", batch=8, frames=16, d=512,
               layers=4, heads=8, ff=2048, teacher=None, student_dims=(2048,), pyramid=False, distill=False),
    "c2": dict(desc="C2 cross-attention fusion of 3 expert streams (2048/1024/128-d), 32 frames, batch 64", batch=64, frames=32,
               d=512, layers=4, heads=8, ff=2048, teacher=None, student_dims=(2048, 1024, 128), pyramid=False, distill=False,
               fusion="cross"),
    "c3": dict(desc="C3 pyramid network (groups 2,3,4) over 64 frames, batch 128", batch=128, frames=64, d=512, layers=4,
               heads=8, ff=2048, teacher=None, student_dims=(512,), pyramid=True, distill=False),
    "c4": dict(desc="C4 frozen 3-expert teacher -> RGB student distillation (KL+CE), 32 frames, 256 clips/GPU", batch=256,
               frames=32, d=512, layers=4, heads=8, ff=2048, teacher=(2048, 1024, 128), student_dims=(2048,), pyramid=False,
               distill=True),
    "c5": dict(desc="C5 pyramid + cross-attention teacher + distillation, 128 frames, d=768, 12 layers, 256 clips/GPU "
                    "(global batch 2048 on 8 GPUs)", batch=256, frames=128, d=768, layers=12, heads=12, ff=3072,
               teacher=(2048, 1024, 128), student_dims=(2048,), pyramid=True, distill=True),
}
N_CLASSES = 15
# whole-step CUDA-graph replay of the training step (hostapi.GraphedTrainStep) unless --graph 0: 4x at C1, 2.5x at C2, ~1 % at C5
GRAPH_DEFAULT = True


def flops_per_clip(w):
    """Algorithmic FLOPs per clip per training step: 3x forward for trained nets,
    1x for the frozen teacher; attention counted un-padded."""
    T, d, L, ff = w["frames"], w["d"], w["layers"], w["ff"]
    S = T + 1

    def net(dims, fusion, pyramid):
        f = 0.0
        for D in dims:
            f += 2 * T * D * d                                           # input projection
            f += L * (2 * S * d * 3 * d + 2 * S * d * d + 4 * S * d * ff)  # encoder GEMMs
            f += L * 4 * S * S * d                                       # attention
        if fusion == "cross" and len(dims) > 1:
            E = len(dims)
            f += 4 * S * d * d + 4 * (E - 1) * S * d * d + 4 * S * (E - 1) * S * d + 4 * S * d * ff
        if pyramid:
            for g in (2, 3, 4):
                f += 2 * d * (T // g) * 512 + 2 * 512 * 512 + 2 * 512 * N_CLASSES
        return f + 2 * d * N_CLASSES

    student = net(w["student_dims"], w.get("fusion", "sum"), w["pyramid"])
    teacher = net(w["teacher"], "cross", False) if w["teacher"] else 0.0
    return 3 * student + teacher


def synth_batch(w, batch, seed, device="cpu", pin=False):
    """Synthetic expert features: post-ReLU-like non-negative RGB / motion features,
    raw Gaussian audio, 30 % of (clip, frame, expert) vectors zeroed like the loader's augmentation;
    multi-hot labels with at least one positive."""
    g = torch.Generator().manual_seed(seed)
    dims = w["teacher"] if w["teacher"] else w["student_dims"]
    xs = []
    for D in dims:
        x = torch.randn(batch, w["frames"], D, generator=g)
        if D > 128:
            x = torch.relu(x * 0.5)
        keep = (torch.rand(batch, w["frames"], 1, generator=g) >= 0.3).float()
        xs.append(x * keep)
    y = (torch.rand(batch, N_CLASSES, generator=g) < 0.15).float()
    y[torch.arange(batch), torch.randint(0, N_CLASSES, (batch,), generator=g)] = 1.0
    if pin:
        xs = [x.pin_memory() for x in xs]
        y = y.pin_memory()
    if device != "cpu":
        xs = [x.to(device) for x in xs]
        y = y.to(device)
    return xs, y


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU oracle arm
def build_oracle(w, batch):
    from oracle import param
    common = dict(d=w["d"], nhead=w["heads"], nhid=w["ff"], nlayers=w["layers"], dropout=0.5, batch_size=batch, frames=w["frames"],
                  n_classes=N_CLASSES)
    torch.manual_seed(1130)
    teacher = param.FusionTransformer(in_dims=w["teacher"], fusion="cross", **common).eval() if w["teacher"] else None
    student = param.FusionTransformer(in_dims=w["student_dims"], fusion=w.get("fusion", "sum"), pyramid=w["pyramid"], **common).train()
    return teacher, student


def oracle_step(w, teacher, student, opt, xs, y):
    from oracle import param
    opt.zero_grad(set_to_none=False)
    if teacher is not None:
        with torch.no_grad():
            t_logits, _ = teacher(xs)
        s_logits, pyr = student(xs[:len(w["student_dims"])])
        loss, _ = param.distill_loss(s_logits, t_logits, y, temperature=2.0, alpha=1.0, pyramid=pyr)
    else:
        s_logits, pyr = student(xs)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(s_logits, y)
        if pyr is not None:
            loss = loss + torch.nn.functional.binary_cross_entropy(pyr, y)
    loss.backward()
    opt.step()
    return float(loss.detach())


def time_oracle(w, sample_clips, steps, warmup):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    teacher, student = build_oracle(w, sample_clips)
    opt = torch.optim.AdamW(student.parameters(), lr=1e-4, weight_decay=0.01)
    xs, y = synth_batch(w, sample_clips, 1130)
    for _ in range(warmup):
        oracle_step(w, teacher, student, opt, xs, y)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle_step(w, teacher, student, opt, xs, y)
    dt = (time.perf_counter() - t0) / steps
    return sample_clips / dt, dt, cores


def run_reference(args, w, rank):
    if rank != 0:
        return
    sample = args.cpu_clips or max(1, min(w["batch"], int(1.2e12 / flops_per_clip(w)) or 1))   # ~2 s of CPU work per step at C5
    cps, dt, cores = time_oracle(w, sample, args.steps, args.warmup)
    line = {"impl": "reference", "metric": "train clips/sec fwd+bwd", "value": round(cps, 3), "unit": "clips/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 2), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": w["desc"], "sample_clips_per_step": sample, "optimizer": "AdamW", "device": "host CPU"},
            "cpu_baseline": {"value": round(cps, 3), "unit": "clips/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} clips/step of the same workload (oracle/param.py, torch {torch.__version__} CPU, fp32)"},
            "e2e": {"value": round(cps, 3), "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def build_model(w, batch, precision, dropout, device):
    from tvt_b200 import hostapi
    common = dict(d=w["d"], nhead=w["heads"], nhid=w["ff"], nlayers=w["layers"], dropout=dropout, batch_size=batch,
                  frames=w["frames"], n_classes=N_CLASSES, precision=precision)
    torch.manual_seed(1130)
    student = hostapi.FusionTransformer(in_dims=w["student_dims"], fusion=w.get("fusion", "sum"), pyramid=w["pyramid"], **common).to(device)
    if w["teacher"]:
        teacher = hostapi.FusionTransformer(in_dims=w["teacher"], fusion="cross", **common).to(device)
        return hostapi.DistillationTrainer(teacher, student, temperature=2.0, alpha=1.0).train()
    return student.train()


def gpu_step(w, model, reducer, opt, xs, y):
    from tvt_b200.functions import DistillLossFn
    reducer.zero_grad()
    if w["teacher"]:
        loss = model.training_step({"experts": xs, "label": y})
    else:
        logits, _, ploss = model(xs, y if w["pyramid"] else None)
        loss = DistillLossFn.apply(logits, None, y, 1.0, 0.0, 0.0, 1.0)[0]
        if ploss is not None:
            loss = loss + ploss[0]
    loss.backward()
    reducer.finish()
    opt.step()
    return loss


def time_torch_gpu(w, B, dev, steps=5, warmup=3):
    """The library path on the SAME GPU: the oracle's torch.nn composition (what the reference's modules dispatch to)
    under stock bf16 autocast — cuBLAS GEMMs, ATen SDPA, ATen LayerNorm, fused AdamW.  SURVEY fact 1: this, not the CPU,
    is the kernel path to beat.  Returns clips/s, ms/step."""
    from oracle import param
    teacher, student = build_oracle(w, B)
    student.to(dev)
    if teacher is not None:
        teacher.to(dev)
    opt = torch.optim.AdamW(student.parameters(), lr=1e-4, weight_decay=0.01, fused=True)
    xs, y = synth_batch(w, B, 1130, device=dev)
    ns = len(w["student_dims"])

    def step():
        opt.zero_grad(set_to_none=False)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            if teacher is not None:
                with torch.no_grad():
                    t_logits, _ = teacher(xs)
                s_logits, pyr = student(xs[:ns])
            else:
                s_logits, pyr = student(xs)
                t_logits = None
        pyr = None if pyr is None else pyr.float().clamp(1e-6, 1 - 1e-6)
        if t_logits is not None:
            loss, _ = param.distill_loss(s_logits.float(), t_logits.float(), y, temperature=2.0, alpha=1.0, pyramid=pyr)
        else:
            loss = torch.nn.functional.binary_cross_entropy_with_logits(s_logits.float(), y)
            if pyr is not None:
                loss = loss + torch.nn.functional.binary_cross_entropy(pyr, y)
        loss.backward()
        opt.step()

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del teacher, student, opt, xs, y
    torch.cuda.empty_cache()
    return B / ms * 1e3, ms


def run_tvt(w, args, ctx, steps, warmup, sample_clocks=False, graph=None):
    """One workload on this rank's GPU: W warm-up + K timed steps with resident inputs, K timed steps end to end from
    pinned host batches, one CUDA-event-instrumented step.  Returns a dict of raw measurements (max over ranks)."""
    import torch.distributed as dist
    from tvt_b200 import capi, ddp, optim
    rank, local, world, dev = ctx
    B = w["batch"]
    graph = args.graph if graph is None else graph
    model = build_model(w, B, args.precision, args.dropout, dev)
    trainable = [p for p in model.parameters() if p.requires_grad]
    reducer = ddp.GradBucketReducer(trainable, bucket_bytes=int(os.environ.get("TVT_BUCKET_MB", "32")) << 20, average=False)
    student = model.student if w["teacher"] else model
    # flat-bucket AdamW: one launch per gradient bucket, 1 / world averaging and bf16 weight planes fused
    opt = optim.FlatOptimizer(reducer, modes=[student.mode], kind="adamw", lr=1e-4, weight_decay=0.01)

    # two resident batches (alternated) + the same two as pinned host batches for the e2e leg.  Host features are stored
    # in the compute dtype (bf16 mode: bf16 — what a feature store for this path holds; the model's first op is that cast)
    hdt = torch.bfloat16 if (args.precision == "bf16" and args.host_dtype == "bf16") else torch.float32
    host = []
    for i in range(2):
        xs, y = synth_batch(w, B, 1130 + rank + 1000 * i)
        host.append(([x.to(hdt).pin_memory() for x in xs], y.pin_memory()))
    resident = [([x.to(dev) for x in xs], y.to(dev)) for xs, y in host]
    h2d_bytes = sum(x.numel() * x.element_size() for x in host[0][0]) + host[0][1].numel() * 4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(run_one, nsteps, finish=None):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(nsteps):
            run_one(i)
        if finish is not None:
            finish()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms / nsteps

    stepper = None
    step_fn = lambda xs, y: gpu_step(w, model, reducer, opt, xs, y)   # noqa: E731
    if graph:
        from tvt_b200.hostapi import GraphedTrainStep
        try:
            stepper = GraphedTrainStep(lambda xs, y: gpu_step(w, model, reducer, opt, xs, y), opt, resident[0], warmup=3)
            stepper.prepare(*resident[1])
            step_fn = lambda xs, y: stepper(xs, y)            # noqa: E731
        except Exception as e:      # every rank runs the same deterministic program, so all ranks fall back together
            print(f"bench.py: CUDA-graph capture of the training step failed ({e!r:.200}); falling back to eager launches", file=sys.stderr)
            if stepper is not None:
                stepper.close()
            stepper, graph = None, False
            torch.cuda.synchronize()

    # ---- leg 1: inputs resident in HBM
    for i in range(warmup):
        step_fn(*resident[i % 2])
    sampler = ClockSampler(local) if sample_clocks else None
    if sampler is not None and rank == 0:
        sampler.start()
    l0 = capi.launches
    ms = timed(lambda i: step_fn(*resident[i % 2]), steps)
    launches = (capi.launches - l0) // steps
    if stepper is not None:
        launches = stepper.kernels_per_replay
    clocks = sampler.stop() if (sampler is not None and rank == 0) else None

    # ---- leg 2: end to end from pinned host batches, copies prefetched on a side stream
    copy_stream = torch.cuda.Stream(device=dev)
    # two preallocated device slots (no allocator traffic in the loop); slot i % 2 is refilled for step i + 2 only after
    # step i, its last reader, has finished on the compute stream
    slots = [([torch.empty_like(x, device=dev) for x in xs], torch.empty_like(y, device=dev)) for xs, y in host]
    slot_free = [None, None]
    if stepper is not None:
        for sl in slots:
            stepper.prepare(*sl)                            # one graph per input buffer set, captured outside the timed regions

    def prefetch(i):
        if slot_free[i % 2] is not None:
            copy_stream.wait_event(slot_free[i % 2])
        with torch.cuda.stream(copy_stream):
            xs, y = host[i % 2]
            dxs, dy_ = slots[i % 2]
            for dst, src in zip(dxs, xs):
                dst.copy_(src, non_blocking=True)
            dy_.copy_(y, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ev

    losses = []
    state = {"ev": None, "pending": None}
    loss_host = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]

    def read_pending():
        # device -> host read of a step's loss: the copy into pinned memory was enqueued right behind that step; it is
        # collected one step later (as a training loop logs), after the NEXT step has been enqueued, so the read never
        # drains the GPU.  Every step's loss is read inside the timed region (finish() collects the last one).
        if state["pending"] is not None:
            ev, buf = state["pending"]
            ev.synchronize()
            losses.append(float(buf[0]))
            state["pending"] = None

    def e2e_one(i):
        if state["ev"] is None:
            state["ev"] = prefetch(i)
        torch.cuda.current_stream().wait_event(state["ev"])
        xs, y = slots[i % 2]
        state["ev"] = prefetch(i + 1)                       # next step's inputs copy while this step computes
        loss = step_fn(xs, y)
        buf = loss_host[i % 2]
        buf.copy_(loss.detach().reshape(1).float(), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        slot_free[i % 2] = ev                               # this step was the slot's last reader
        read_pending()                                      # step i - 1's loss, now that step i is in the queue
        state["pending"] = (ev, buf)

    e2e_one(0)
    read_pending()
    state["ev"] = None
    ms_e2e = timed(e2e_one, steps, finish=read_pending)
    assert len(losses) == steps + 1 and all(math.isfinite(v) for v in losses), "e2e leg: a step's loss was not read back"

    # ---- data-parallel replicas must hold identical parameters after the same number of identical updates
    checksum_equal = None
    if world > 1:
        cs = torch.zeros(2, dtype=torch.float64, device=dev)
        for b in opt.buckets:
            cs[0] += b["p"].double().sum()
            cs[1] += b["p"].double().abs().sum()
        lo, hi = cs.clone(), cs.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        checksum_equal = bool(torch.equal(lo, hi))
        dry = os.environ.get("TVT_DDP_DRY_RUN", "0") == "1"   # measurement knob (ddp.GradBucketReducer): no all-reduces, replicas diverge by design
        assert checksum_equal or dry, f"rank {rank}: parameter checksums differ across ranks after training: {lo.tolist()} vs {hi.tolist()}"

    # ---- instrumented step: CUDA events around every kernel-launching C-ABI call (not part of the timings)
    breakdown = capi.profile_step(lambda: gpu_step(w, model, reducer, opt, *resident[0]))
    detail = dict(capi.last_profile_detail)
    out = {"ms": ms, "ms_e2e": ms_e2e, "launches": int(launches), "h2d_bytes": h2d_bytes, "clocks": clocks, "breakdown": breakdown,
           "detail": detail, "checksum_equal": checksum_equal, "host_dtype": str(hdt).replace("torch.", ""), "graph": bool(graph),
           "final_loss": losses[-1]}
    reducer.remove()
    if stepper is not None:
        stepper.close()
    del model, reducer, opt, stepper, step_fn, resident, slots, host, student, trainable
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return out


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def summarize_kernels(w, r, peaks):
    """Dominant kernel (GEMM) roofline + the attention metric from the instrumented step of run ``r``."""
    B = w["batch"]
    breakdown = r["breakdown"]
    gemm = breakdown.get("tvt_gemm", {"ms": 0.0, "flops": 0.0, "calls": 0})
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    achieved = gemm["flops"] / (gemm["ms"] * 1e-3) / 1e12 if gemm["ms"] > 0 else 0.0
    total_ms = sum(v["ms"] for v in breakdown.values()) or 1.0
    n_tok, burst = B * (w["frames"] + 1), peaks.get("bf16_tflops", peak_tf)
    att = {}
    for name, tensors in (("fwd", 4), ("bwd", 8)):
        v = breakdown.get(f"tvt_attention_{name}")
        if v and v["ms"] > 0:
            tf = v["flops"] / (v["ms"] * 1e-3) / 1e12
            gbs = v["calls"] * tensors * n_tok * w["d"] * 2 / (v["ms"] * 1e-3) / 1e9
            att[name] = {"tflops": round(tf, 1), "frac_of_bf16_peak": round(tf / burst, 4), "hbm_gbs": round(gbs, 1),
                         "frac_of_hbm_peak": round(gbs / peaks.get("hbm_gbs", 6500.0), 4), "launches_per_step": v["calls"],
                         "us_per_launch": round(v["ms"] * 1e3 / v["calls"], 1)}
    shares = {k: round(v["ms"] / total_ms, 4) for k, v in sorted(breakdown.items(), key=lambda kv: -kv[1]["ms"])[:6]}
    return gemm, achieved, peak_tf, total_ms, att, shares


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="tvt", choices=["tvt", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--dropout", type=float, default=0.5, help="config.yaml dropout (0.5 in the reference)")
    ap.add_argument("--batch", type=int, default=0, help="clips per GPU (default: the workload's)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: the workload's batch is the GLOBAL batch, split evenly over the ranks (SURVEY 8e: C4 256 -> 128/64/32)")
    ap.add_argument("--host-dtype", default="bf16", choices=["bf16", "fp32"], help="dtype of the pinned host feature batches (e2e leg)")
    ap.add_argument("--graph", type=int, default=-1, help="1 (default): replay the whole training step as a CUDA graph; 0: eager launches")
    ap.add_argument("--cpu-clips", type=int, default=0, help="clips per step of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the C1-C4 lines and the stock-PyTorch-on-GPU yardstick")
    ap.add_argument("--breakdown", default="", help="write the per-kernel CUDA-event breakdown to this file")
    args = ap.parse_args()
    args.graph_explicit = args.graph >= 0
    w = dict(WORKLOADS[args.workload])
    if args.batch:
        w["batch"] = args.batch
    args.warmup = max(args.warmup, 3) if args.impl == "tvt" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        run_reference(args, w, rank)
        return

    import torch.distributed as dist
    import tvt_b200  # noqa: F401
    from tvt_b200 import capi, ddp
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the tvt arm has no CPU path (use --impl reference for the CPU oracle)")
    # NCCL prints its version banner on stdout when the communicator is created (NCCL_DEBUG=VERSION/WARN ignore
    # NCCL_DEBUG_FILE): create it now, with file descriptor 1 pointed at stderr, so stdout carries the JSON line only
    sys.stdout.flush()
    saved_fd = os.dup(1)
    os.dup2(2, 1)
    try:
        rank, local, world = ddp.init_from_env()
        dev = torch.device("cuda", local)
        if world > 1:
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved_fd, 1)
        os.close(saved_fd)
    capi.load()
    if capi.load().tvt_device_check() != 0:
        raise SystemExit("bench.py: " + capi.last_error())
    ctx = (rank, local, world, dev)
    global_batch = w["batch"] * world
    if args.scaling == "strong":
        if w["batch"] % world:
            raise SystemExit(f"bench.py: --scaling strong needs the global batch {w['batch']} divisible by {world} ranks")
        global_batch = w["batch"]
        w["batch"] //= world
    B = w["batch"]
    use_graph = GRAPH_DEFAULT if args.graph < 0 else bool(args.graph)
    args.graph = use_graph
    r = run_tvt(w, args, ctx, args.steps, args.warmup, sample_clocks=True)

    # ---- the other BASELINE configs (C1-C4), few steps each, same legs; C4 also with its strong-scaling split
    extras = {}
    if not args.no_extras and args.workload == "c5":
        for name in ("c1", "c2", "c3", "c4"):
            for scaling in (("weak", "strong") if (name == "c4" and world > 1) else ("weak",)):
                wx = dict(WORKLOADS[name])
                gb = wx["batch"] * world
                if scaling == "strong":
                    gb = wx["batch"]
                    wx["batch"] //= world
                rx = run_tvt(wx, args, ctx, min(args.steps, 20), 3, graph=(args.graph if args.graph_explicit else GRAPH_DEFAULT))
                extras[name if scaling == "weak" else name + "_strong"] = (wx, gb, scaling, rx)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s sustained (of fallback)"
    gemm, achieved, peak_tf, total_ms, att, shares = summarize_kernels(w, r, peaks)
    traffic = None
    try:   # dram__bytes_read + write per launch from the committed ncu --set full capture of this kernel
        traffic = json.load(open(os.path.join(ROOT, "profiles", "gemm_traffic.json")))["mean_dram_bytes_per_launch"]
    except Exception:
        pass
    fl = flops_per_clip(w)
    ms, ms_e2e = r["ms"], r["ms_e2e"]
    value = world * B / (ms * 1e-3)
    line = {
        "metric": "train clips/sec fwd+bwd", "value": round(value, 1), "unit": "clips/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": args.precision, "data": "synthetic",
        "config": {"workload": w["desc"], "clips_per_gpu": B, "global_batch": global_batch, "frames": w["frames"], "d_model": w["d"],
                   "layers": w["layers"], "heads": w["heads"], "ff": w["ff"], "dropout": args.dropout, "optimizer": "AdamW (tvt flat-bucket kernel)",
                   "parallelism": f"dp{world}", "l2": "inputs larger than L2 (%.0f MB per step, two alternating batches)" % (r["h2d_bytes"] / 1e6),
                   "gflop_per_clip_step": round(fl / 1e9, 2), "host_feature_dtype": r["host_dtype"],
                   "launch_mode": "whole-step CUDA graph replay" if r["graph"] else "eager launches chained with programmatic dependent launch"},
        "e2e": {"value": round(world * B / (ms_e2e * 1e-3), 1), "unit": "clips/s", "h2d_bytes_per_step": r["h2d_bytes"], "d2h_bytes_per_step": 4,
                "ms_per_step": round(ms_e2e, 3),
                "loss_read": "every step, via pinned memory, collected one step later (after the next step is enqueued); the last inside the timed region"},
        "gpu_launches": int(r["launches"]),
        "clocks": r["clocks"],
        "model_tflops": round(value / world * fl / 1e12, 1),
        "model_frac_of_peak": round(value / world * fl / 1e12 / peak_tf, 4),
        "roofline": {"kernel": "tvt::gemm::gemm_kernel (tcgen05)", "bound": "tensor", "achieved": round(achieved, 1), "peak": peak_tf,
                     "unit": "TFLOP/s", "frac": round(achieved / peak_tf, 4), "traffic": traffic,
                     "traffic_source": "constant: mean dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full "
                                       "capture (profiles/gemm_traffic.json), not measured in this run",
                     "peak_source": peak_src, "launches_per_step": gemm["calls"], "share_of_kernel_time": round(gemm["ms"] / total_ms, 4)},
        "kernel_time_shares": shares,
    }
    if r["checksum_equal"] is not None:
        line["ranks_hold_identical_parameters"] = r["checksum_equal"]
        if os.environ.get("TVT_DDP_DRY_RUN", "0") == "1":
            line["ddp_dry_run"] = "TVT_DDP_DRY_RUN=1: gradient all-reduces skipped (measurement of the communication cost only; not a valid training run)"
    # BASELINE.json's second metric: attention TFLOP/s against the bf16 peak (un-padded algorithmic FLOPs of every attention
    # launch of the instrumented step).  At S = frames + 1 <= 129 and head_dim 64 the kernels have S / 2 FLOP per HBM byte, below
    # the ridge, so the bandwidth figure (q, k, v, o / their gradients, bf16) is the roofline that actually bounds them.
    line["attention"] = att
    if extras:
        cfgs = {}
        for key, (wx, gb, scaling, rx) in extras.items():
            gx, ach, _, tot, attx, sh = summarize_kernels(wx, rx, peaks)
            flx = flops_per_clip(wx)
            vx = gb / (rx["ms"] * 1e-3)
            dom = next(iter(sh)) if sh else None
            cfgs[key] = {"workload": wx["desc"], "scaling": scaling, "clips_per_gpu": wx["batch"], "global_batch": gb,
                         "value": round(vx, 1), "unit": "clips/s", "ms_per_step": round(rx["ms"], 3),
                         "e2e": {"value": round(gb / (rx["ms_e2e"] * 1e-3), 1), "ms_per_step": round(rx["ms_e2e"], 3),
                                 "h2d_bytes_per_step": rx["h2d_bytes"], "d2h_bytes_per_step": 4},
                         "gpu_launches": rx["launches"], "launch_mode": "cuda graph" if rx["graph"] else "eager",
                         "model_tflops": round(vx / world * flx / 1e12, 1), "model_frac_of_peak": round(vx / world * flx / 1e12 / peak_tf, 4),
                         "dominant_kernel": dom, "kernel_time_shares": sh,
                         "roofline": {"kernel": "tvt::gemm::gemm_kernel (tcgen05)", "bound": "tensor", "achieved": round(ach, 1), "peak": peak_tf,
                                      "unit": "TFLOP/s", "frac": round(ach / peak_tf, 4)},
                         "attention": attx}
        line["configs"] = cfgs
    if world == 1 and not args.no_extras:
        # the library path on the same GPU (stock PyTorch: cuBLAS + ATen SDPA / LayerNorm + fused AdamW under bf16 autocast)
        try:
            cps, ms_t = time_torch_gpu(w, B, dev)
            line["gpu_library_baseline"] = {"value": round(cps, 1), "unit": "clips/s", "ms_per_step": round(ms_t, 3),
                                            "what": f"oracle/param.py modules on the same GPU under torch.autocast(bf16), fused AdamW, torch {torch.__version__}; "
                                                    "5 timed steps after 3 warm-up, same workload and batch",
                                            "tvt_over_library": round(value / cps, 3)}
        except Exception as e:      # informational: never fail the bench line for it
            line["gpu_library_baseline"] = {"unavailable": repr(e)[:200]}
    if not args.no_cpu_baseline and world == 1:
        # rank 0 at N = 1 only (at N > 1 the other ranks would sit in a barrier while the host runs the oracle)
        sample = args.cpu_clips or max(1, min(B, int(1.2e12 / fl) or 1))
        cps, dt, cores = time_oracle(w, sample, 3, 1)
        line["cpu_baseline"] = {"value": round(cps, 3), "unit": "clips/s", "cores": cores, "kind": "port",
                                "sample": f"{sample} clips/step x 3 steps (+1 warm-up) of the same workload (oracle/param.py on torch CPU, fp32), "
                                          f"{dt * 3:.1f} s of CPU work"}
    if args.breakdown:
        breakdown = r["breakdown"]
        with open(args.breakdown, "w") as f:
            f.write(f"# per-kernel CUDA-event breakdown of one instrumented step ({w['desc']}); ms_per_step timed = {ms:.3f}\n")
            for k, v in sorted(breakdown.items(), key=lambda kv: -kv[1]["ms"]):
                tf = v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 and v["flops"] else 0.0
                f.write(f"{k:28s} calls={v['calls']:5d}  ms={v['ms']:9.3f}  share={v['ms'] / total_ms:6.3f}  TFLOP/s={tf:8.1f}\n")
            f.write("# per launch shape (GEMM stage letters: b bias, R relu, m relu mask, d dropout, r residual, F fp32 out, L hi/lo planes)\n")
            for (k, det), v in sorted(r["detail"].items(), key=lambda kv: -kv[1]["ms"]):
                if not det:
                    continue
                tf = v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 and v["flops"] else 0.0
                f.write(f"  {k:20s} {det:62s} calls={v['calls']:4d}  ms={v['ms']:8.3f}  us/call={v['ms'] * 1e3 / v['calls']:8.1f}  TFLOP/s={tf:7.1f}\n")
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
