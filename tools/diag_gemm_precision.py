"""Diagnostic: error of the fp32-mode (hi/lo bf16 planes, 3 tcgen05 passes) GEMM against float64 as a function of the length
of the accumulation chain kept in TMEM (split-K with fp32 atomics shortens it) and of the exact three-plane layout."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, tvt_b200
from tvt_b200 import ops
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda:0"
g = torch.Generator().manual_seed(0)
for (M, N, K) in ((256, 512, 896), (2112, 2048, 512), (2112, 512, 2048)):
    x = torch.relu(torch.randn(M, K, generator=g)).to(dev)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
    ref = x.double() @ w.double().t()
    def err(y): return float((y.double() - ref).norm() / ref.norm())
    print(f"M={M} N={N} K={K}: torch fp32 {err(x @ w.t()):.2e}", end="")
    m = ops.Mode("fp32")
    xp = m.split(x); wh, wl = m.weight(w)
    kb = (K + 63) // 64
    for s in (1, 2, 4, 8, 16, 32):
        if s > kb: continue
        out = torch.zeros(M, N, device=dev)
        ops.gemm(xp[0], wh, M, N, K, a_lo=xp[1], b_lo=wl, out_f32=out, splits=s, atomic=s > 1)
        print(f"  splits={s}: {err(out):.2e}", end="")
    a4 = ops.split_f32x3(x, 0); b4 = ops.split_f32x3(w, 1)
    for s in (1, 8, 32):
        out = torch.zeros(M, N, device=dev)
        ops.gemm(a4[0], b4[0], M, N, 4 * K, a_lo=a4[1], b_lo=b4[1], out_f32=out, splits=s, atomic=s > 1)
        print(f"  x3 splits={s}: {err(out):.2e}", end="")
    # single pass bf16 for scale
    out = torch.zeros(M, N, device=dev)
    ops.gemm(xp[0], wh, M, N, K, out_f32=out)
    print(f"  1-pass bf16: {err(out):.2e}")
