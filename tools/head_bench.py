import os, sys
sys.path.insert(0, "/root/repo")
import torch, tvt_b200
from tvt_b200 import ops
dev = "cuda"
for M, K, C in ((256, 768, 15), (256, 512, 15)):
    x = torch.randn(M, K, device=dev).bfloat16(); w = torch.randn(C, K, device=dev); b = torch.randn(C, device=dev)
    dy = torch.randn(M, C, device=dev)
    for name, fn in (("fwd", lambda: ops.head_linear_fwd(x, w, b)), ("bwd", lambda: ops.head_linear_bwd(x, w, dy))):
        for _ in range(5): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(50): fn()
        e1.record(); torch.cuda.synchronize()
        print(M, K, C, name, f"{e0.elapsed_time(e1) / 50 * 1e3:.1f} us per call")
