"""Phase breakdown of the tcgen05 attention backward (SM-clock stamps of three CTAs); needs tools/attn_stamps.sh."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, tvt_b200
from tvt_b200 import capi, ops
capi.LIB_PATH = os.path.join(ROOT, "tools", "bin", "libtvt_stamps.so")
B, S, d, H = 256, 129, 768, 12
n = B * S
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(n, 3 * d, device="cuda", generator=g).to(torch.bfloat16)
do = torch.randn(n, d, device="cuda", generator=g).to(torch.bfloat16)
dqkv = torch.empty_like(qkv)
LABELS = {0: "entry", 1: "sync #1 (TMEM, tail scratch)", 2: "sync #2 (D partials)", 3: "S,dP ready", 4: "P/dS written", 5: "sync #3",
          6: "dV,dK ready", 7: "dK/dV drained", 8: "dQ + tail keys ready", 9: "tail keys + dQ drained", 10: "final sync",
          16: "issuer: operands landed", 17: "issuer: S,dP issued", 18: "issuer: dV,dK issued", 19: "issuer: dQ + tail keys issued",
          24: "tail warp 0 start", 25: "tail warp 1 start", 28: "tail warp 0 done", 29: "tail warp 1 done", 30: "tail warp 2 done",
          31: "tail warp 3 done"}
for p in (0.0, 0.5):
    o, lse = ops.attention_fwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], B, H, S, S, 64, 0.125, dropout_p=p, seed=1)
    for _ in range(3):
        ops.attention_bwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], o, do, lse, dqkv[:, :d], dqkv[:, d:2 * d], dqkv[:, 2 * d:], B, H, S, S, 64,
                          0.125, dropout_p=p, seed=1)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 96)()
    assert capi.load().tvt_debug_attn_stamps(buf) == 0
    for c in range(3):
        st = list(buf[32 * c:32 * c + 32])
        print(f"--- dropout {p}  CTA {['first', 'middle', 'last'][c]}  total {st[10] - st[0]} clk")
        ev = [(st[k] - st[0], v) for k, v in LABELS.items()]
        for t, name in sorted(ev):
            print(f"   {t:8d}  {name}")
