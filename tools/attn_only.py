import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, tvt_b200
from tvt_b200 import ops
B, S, d, H = 256, 129, 768, 12
n = B * S
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(n, 3 * d, device="cuda", generator=g).to(torch.bfloat16)
do = torch.randn(n, d, device="cuda", generator=g).to(torch.bfloat16)
dqkv = torch.empty_like(qkv)
for _ in range(3):
    o, lse = ops.attention_fwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], B, H, S, S, 64, 0.125)
    ops.attention_bwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], o, do, lse, dqkv[:, :d], dqkv[:, d:2 * d], dqkv[:, 2 * d:], B, H, S, S, 64, 0.125)
torch.cuda.synchronize()
print("done")
