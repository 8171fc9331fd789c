"""Diagnostic: where the fp32-accumulate mode's gradient error at full BASELINE sizes comes from.  Compares, per parameter,
(a) tvt fp32 mode vs the fp32 oracle, (b) tvt vs a float64 run of the oracle, (c) the fp32 oracle itself vs float64."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, torch.nn.functional as F
import bench, tvt_b200
from tvt_b200 import hostapi
from oracle import param
from util import rel_err, copy_state
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
DEV = "cuda:0"
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
w = bench.WORKLOADS[name]; B = int(sys.argv[2]) if len(sys.argv) > 2 else w["batch"]; T = w["frames"]
common = dict(d=w["d"], nhead=w["heads"], nhid=w["ff"], nlayers=w["layers"], dropout=0.0, batch_size=B, frames=T, n_classes=15)
torch.manual_seed(1130)
ref = param.FusionTransformer(in_dims=w["student_dims"], fusion=w.get("fusion", "sum"), pyramid=w["pyramid"], **common).to(DEV)
mod = copy_state(hostapi.FusionTransformer(in_dims=w["student_dims"], fusion=w.get("fusion", "sum"), pyramid=w["pyramid"], precision="fp32", **common), ref).to(DEV)
for m in (ref, mod):
    for s in m.modules():
        if isinstance(s, torch.nn.Dropout): s.p = 0.0
gen = torch.Generator().manual_seed(1130)
xs = [(torch.relu(torch.randn(B, T, D, generator=gen) * 0.5) if D > 128 else torch.randn(B, T, D, generator=gen)).to(DEV) for D in w["student_dims"]]
y = (torch.rand(B, 15, generator=gen) < 0.15).float().to(DEV)
def loss_of(m, xs, y):
    lg, pr = m(xs)
    l = F.binary_cross_entropy_with_logits(lg, y)
    return l + F.binary_cross_entropy(pr, y) if pr is not None else l
loss_of(ref, xs, y).backward()
r64 = copy.deepcopy(ref).double(); r64.zero_grad(set_to_none=True)
loss_of(r64, [x.double() for x in xs], y.double()).backward()
from tvt_b200.functions import DistillLossFn
mod.train()
lg, _, pl = mod(xs, y if w["pyramid"] else None)
l = DistillLossFn.apply(lg, None, y, 1.0, 0.0, 0.0, 1.0)[0]
(l + pl[0] if pl is not None else l).backward()
g32 = dict(ref.named_parameters()); g64 = dict(r64.named_parameters())
tot = [0.0, 0.0, 0.0, 0.0]
rows = []
for n, p in mod.named_parameters():
    if g32[n].grad is None: continue
    a, b, c = rel_err(p.grad, g32[n].grad), rel_err(p.grad, g64[n].grad), rel_err(g32[n].grad, g64[n].grad)
    rows.append((a, n, b, c))
    d64 = g64[n].grad.double()
    tot[0] += float((p.grad.double() - g32[n].grad.double()).pow(2).sum()); tot[1] += float((p.grad.double() - d64).pow(2).sum())
    tot[2] += float((g32[n].grad.double() - d64).pow(2).sum()); tot[3] += float(d64.pow(2).sum())
rows.sort(reverse=True)
print(f"{name} B={B}: global normwise grad error  tvt-vs-fp32 {(tot[0]/tot[3])**.5:.2e}  tvt-vs-f64 {(tot[1]/tot[3])**.5:.2e}  fp32-vs-f64 {(tot[2]/tot[3])**.5:.2e}")
for a, n, b, c in (rows if os.environ.get("ALL") else rows[:14]):
    print(f"  {n:60s} tvt-vs-fp32 {a:.2e}  tvt-vs-f64 {b:.2e}  torchfp32-vs-f64 {c:.2e}")
