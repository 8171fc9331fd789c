"""Informational yardstick (not a bench arm): the CPU oracle's torch.nn composition run by stock PyTorch on the
same B200 — cuBLAS / ATen SDPA / ATen LayerNorm under bf16 autocast — at the C5 per-GPU shard.  This is the
library path the hand-written kernels are measured against in DESIGN.md."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from oracle import param

w = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c5"])
B = int(sys.argv[2]) if len(sys.argv) > 2 else w["batch"]
dev = "cuda:0"
teacher, student = bench.build_oracle(w, B)
student.to(dev)
if teacher is not None:
    teacher.to(dev)
opt = torch.optim.AdamW(student.parameters(), lr=1e-4, weight_decay=0.01, fused=True)
xs, y = bench.synth_batch(w, B, 1130, device=dev)


def step():
    opt.zero_grad(set_to_none=False)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        if teacher is not None:
            with torch.no_grad():
                t_logits, _ = teacher(xs)
            s_logits, pyr = student(xs[:len(w["student_dims"])])
        else:
            s_logits, pyr = student(xs)
            t_logits = None
    if t_logits is not None:
        loss, _ = param.distill_loss(s_logits.float(), t_logits.float(), y, temperature=2.0, alpha=1.0,
                                     pyramid=None if pyr is None else pyr.float().clamp(1e-6, 1 - 1e-6))
    else:
        loss = torch.nn.functional.binary_cross_entropy_with_logits(s_logits.float(), y)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
K = 5
for _ in range(K):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(f"stock torch eager bf16-autocast on B200, {w['desc']}: {ms:.2f} ms/step = {B / ms * 1e3:.1f} clips/s "
      f"(dropout 0.5 train mode, fused AdamW, torch {torch.__version__})")
