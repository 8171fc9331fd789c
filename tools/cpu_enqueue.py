"""How long does the host take to enqueue one C5 training step, against the GPU's time for it?  (CPU-bound check.)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench, tvt_b200
from tvt_b200 import capi, ddp, optim
w = dict(bench.WORKLOADS["c5"])
dev = torch.device("cuda", 0)
capi.load()
B = w["batch"]
model = bench.build_model(w, B, "bf16", 0.5, dev)
trainable = [p for p in model.parameters() if p.requires_grad]
reducer = ddp.GradBucketReducer(trainable, bucket_bytes=32 << 20, average=False)
opt = optim.FlatOptimizer(reducer, modes=[model.student.mode], kind="adamw", lr=1e-4, weight_decay=0.01)
xs, y = bench.synth_batch(w, B, 1130)
xs, y = [x.to(dev) for x in xs], y.to(dev)
for _ in range(3):
    bench.gpu_step(w, model, reducer, opt, xs, y)
torch.cuda.synchronize()
K = 8
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(K):
    bench.gpu_step(w, model, reducer, opt, xs, y)
e1.record()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3 * (t1 - t0) / K:.2f} ms/step, GPU {e0.elapsed_time(e1) / K:.2f} ms/step, wall incl. drain {1e3 * (t2 - t0) / K:.2f} ms/step")
# one step from an idle GPU (what the e2e leg sees after loss.item())
for _ in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    loss = bench.gpu_step(w, model, reducer, opt, xs, y)
    t1 = time.perf_counter()
    loss.item()
    t2 = time.perf_counter()
    print(f"from idle: enqueue {1e3 * (t1 - t0):.2f} ms, step complete after {1e3 * (t2 - t0):.2f} ms")
