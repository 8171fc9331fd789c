"""Evaluation-path throughput on the BASELINE configs: inference-only forward + the
evaluation read-out into an EvalBuffer, eager launches versus one CUDA-graph replay per batch.  Inputs resident in
HBM; CUDA events; prints clips/s and launches per batch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import tvt_b200
from tvt_b200 import capi, hostapi

dev = "cuda:0"
for name in sys.argv[1:] or ["c1", "c2", "c3", "c4", "c5"]:
    w = bench.WORKLOADS[name]
    B = w["batch"]
    common = dict(d=w["d"], nhead=w["heads"], nhid=w["ff"], nlayers=w["layers"], dropout=0.5, batch_size=B, frames=w["frames"],
                  n_classes=bench.N_CLASSES, precision="bf16")
    torch.manual_seed(1130)
    # what is evaluated is the deployed network: the student (the teacher only exists during distillation)
    model = hostapi.FusionTransformer(in_dims=w["student_dims"], fusion=w.get("fusion", "sum"), pyramid=w["pyramid"], **common).to(dev).eval()
    wl = dict(w, teacher=None)
    batches = [bench.synth_batch(wl, B, 1130 + i, device=dev) for i in range(4)]
    fwd = lambda *xs: model(list(xs))[0]
    graphed = hostapi.GraphedForward(fwd, batches[0][0])
    buf = hostapi.EvalBuffer(capacity=B * 64, n_classes=bench.N_CLASSES)

    def run(call, iters):
        buf.reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        with torch.no_grad():
            for i in range(iters):
                xs, y = batches[i % 4]
                buf.append(call(*xs), y)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    iters = 60 if name in ("c1", "c2", "c3") else 20
    for label, call in (("eager", fwd), ("graph", graphed)):
        run(call, 4)
        l0 = capi.launches
        ms = run(call, iters)
        print(f"{name} eval {label:5s}: {ms:8.3f} ms/batch  {B / ms * 1e3:10.0f} clips/s   ({(capi.launches - l0) // iters} C-ABI launches per batch on the host side)", flush=True)
    a = fwd(*batches[1][0]); b = graphed(*batches[1][0])
    print(f"   graph == eager: {torch.equal(a, b)}")
