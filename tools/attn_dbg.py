import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, tvt_b200
from tvt_b200 import ops
B, S, d, H = 256, 129, 768, 12
n = B * S
g = torch.Generator(device="cuda").manual_seed(0)
qkv = [torch.randn(n, 3 * d, device="cuda", generator=g).to(torch.bfloat16) for _ in range(4)]
do = torch.randn(n, d, device="cuda", generator=g).to(torch.bfloat16)
dqkv = torch.empty_like(qkv[0])
os.environ.pop("TVT_ATTN_DBG", None) if sys.argv[1] == "fwd" else None
o, lse = ops.attention_fwd(qkv[0][:, :d], qkv[0][:, d:2 * d], qkv[0][:, 2 * d:], B, H, S, S, 64, 0.125)
def fwd(i): ops.attention_fwd(qkv[i % 4][:, :d], qkv[i % 4][:, d:2 * d], qkv[i % 4][:, 2 * d:], B, H, S, S, 64, 0.125)
def bwd(i): ops.attention_bwd(qkv[i % 4][:, :d], qkv[i % 4][:, d:2 * d], qkv[i % 4][:, 2 * d:], o, do, lse, dqkv[:, :d], dqkv[:, d:2 * d], dqkv[:, 2 * d:], B, H, S, S, 64, 0.125)
run = fwd if sys.argv[1] == "fwd" else bwd
for _ in range(3): run(0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for i in range(20): run(i)
e1.record(); torch.cuda.synchronize()
print(sys.argv[1], "dbg", os.environ.get("TVT_ATTN_DBG", "0"), "%.1f us" % (e0.elapsed_time(e1) / 20 * 1e3))
