"""Weight-gradient GEMM with and without the fused bias gradient (tvt_gemm_args.a_rowsum) against the separate column-sum kernel,
C5 shapes, CUDA events.  python tools/wgrad_time.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, tvt_b200
from tvt_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)
def timed(fn, reps=30):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps
for M, N, K in [(33024, 3072, 768), (33024, 2304, 768), (33024, 768, 3072), (33024, 768, 768), (8448, 2048, 512), (2112, 512, 512)]:
    dy = torch.randn(M, N, device=dev, generator=g).to(torch.bfloat16)
    x = torch.randn(M, K, device=dev, generator=g).to(torch.bfloat16)
    splits = ops.pick_splits(((N + 127) // 128) * ((K + 255) // 256), (M + 63) // 64)
    dw, db = torch.zeros(N, K, device=dev), torch.zeros(N, device=dev)
    t_plain = timed(lambda: ops.gemm(dy, x, N, K, M, a_mn=True, b_mn=True, out_f32=dw, splits=splits, atomic=True))
    t_col = timed(lambda: ops.colsum(dy, db))
    ok = ops.rowsum_supported(N, K, M, splits)
    t_fused = timed(lambda: ops.gemm(dy, x, N, K, M, a_mn=True, b_mn=True, out_f32=dw, splits=splits, atomic=True, a_rowsum=db)) if ok else float("nan")
    print(f"wgrad dW[{N},{K}] over {M} tokens, splits={splits}: plain {t_plain:7.1f} us + colsum {t_col:6.1f} us = {t_plain + t_col:7.1f} us   fused {t_fused:7.1f} us")
