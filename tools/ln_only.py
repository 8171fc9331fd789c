"""Runs LayerNorm forward / backward a few times at the C5 shape (for ncu captures)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tvt_b200
from tvt_b200 import ops
n, d = 33024, 768
xs = [torch.randn(n, d, device="cuda").bfloat16() for _ in range(4)]
g, b = torch.ones(d, device="cuda"), torch.zeros(d, device="cuda")
for i in range(3):
    y, mean, rstd = ops.layernorm_fwd(xs[i], g, b)
dg, db, dbias = (torch.zeros(d, device="cuda") for _ in range(3))
for i in range(3):
    ops.layernorm_bwd(xs[(i + 1) % 4], xs[i], mean, rstd, g, dgamma=dg, dbeta=db, dbias=dbias)
torch.cuda.synchronize()
print("ok")
