"""Per-kernel microbenchmarks at the C5 shapes (CUDA events, back-to-back launches, L2-exceeding working sets
rotated between iterations).  Prints achieved TFLOP/s or GB/s next to the measured peaks."""
import json, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import tvt_b200
from tvt_b200 import ops

dev = "cuda:0"
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def report(name, ms, flops=0.0, bytes_=0.0):
    s = f"{name:44s} {ms * 1e3:9.1f} us"
    if flops:
        s += f"  {flops / ms / 1e9:8.1f} TFLOP/s ({flops / ms / 1e9 / peaks['bf16_tflops'] * 100:5.1f}% of measured burst)"
    if bytes_:
        s += f"  {bytes_ / ms / 1e6:8.1f} GB/s ({bytes_ / ms / 1e6 / peaks['hbm_gbs'] * 100:5.1f}% of measured copy)"
    print(s, flush=True)


B, S, d, H, ff = 256, 129, 768, 12, 3072
n = B * S
R = 4  # rotate buffers so consecutive launches do not hit L2-resident data
g = torch.Generator(device=dev).manual_seed(0)
bf = lambda *sh: (torch.randn(*sh, device=dev, generator=g)).to(torch.bfloat16)

# ---- attention
qkv = [bf(n, 3 * d) for _ in range(R)]
for p in (0.0, 0.5):
    ms = timeit(lambda i: ops.attention_fwd(qkv[i % R][:, :d], qkv[i % R][:, d:2 * d], qkv[i % R][:, 2 * d:], B, H, S, S, 64, 0.125, dropout_p=p, seed=1))
    report(f"attention_fwd tcgen05 S=129 p={p}", ms, flops=4.0 * B * H * S * S * 64)
o, lse = ops.attention_fwd(qkv[0][:, :d], qkv[0][:, d:2 * d], qkv[0][:, 2 * d:], B, H, S, S, 64, 0.125)
do = bf(n, d)
dqkv = [torch.empty_like(qkv[0]) for _ in range(R)]
for p in (0.0, 0.5):
    ms = timeit(lambda i: ops.attention_bwd(qkv[i % R][:, :d], qkv[i % R][:, d:2 * d], qkv[i % R][:, 2 * d:], o, do, lse, dqkv[i % R][:, :d],
                                            dqkv[i % R][:, d:2 * d], dqkv[i % R][:, 2 * d:], B, H, S, S, 64, 0.125, dropout_p=p, seed=1))
    report(f"attention_bwd tcgen05 S=129 p={p}", ms, flops=10.0 * B * H * S * S * 64)
q2, kv2 = bf(n, d), bf(2 * n, 2 * d)
ms = timeit(lambda i: ops.attention_fwd(q2, kv2[:, :d], kv2[:, d:], B, H, S, 2 * S, 64, 0.125))
report("attention_fwd cross Sq=129 Sk=258", ms, flops=4.0 * B * H * S * 2 * S * 64)

# ---- layernorm
xs = [bf(n, d) for _ in range(R)]
gamma, beta = torch.ones(d, device=dev), torch.zeros(d, device=dev)
ms = timeit(lambda i: ops.layernorm_fwd(xs[i % R], gamma, beta))
report("layernorm_fwd [33024,768] bf16", ms, bytes_=2.0 * n * d * 2 + 8 * n)
y, mean, rstd = ops.layernorm_fwd(xs[0], gamma, beta)
dg, db, dbias = (torch.zeros(d, device=dev) for _ in range(3))
ms = timeit(lambda i: ops.layernorm_bwd(xs[(i + 1) % R], xs[i % R], mean, rstd, gamma, dgamma=dg, dbeta=db, dbias=dbias))
report("layernorm_bwd [33024,768] bf16", ms, bytes_=3.0 * n * d * 2 + 8 * n)
ms = timeit(lambda i: ops.layernorm_bwd(xs[(i + 1) % R], xs[i % R], mean, rstd, gamma, dgamma=dg, dbeta=db, dbias=dbias, dropout_p=0.5, seed=7))
report("layernorm_bwd + dropout-masked dz (LN2 of a layer)", ms, bytes_=4.0 * n * d * 2 + 8 * n)
ms = timeit(lambda i: ops.layernorm_bwd(xs[(i + 1) % R], xs[i % R], mean, rstd, gamma, dgamma=dg, dbeta=db, dres=xs[(i + 2) % R]))
report("layernorm_bwd + residual-path gradient", ms, bytes_=4.0 * n * d * 2 + 8 * n)
hs = [bf(n, ff) for _ in range(2)]
out = torch.zeros(ff, device=dev)
ms = timeit(lambda i: ops.colsum(hs[i % 2], out))
report("colsum [33024,3072] bf16", ms, bytes_=n * ff * 2.0)

# ---- loader-side pad / drop / noise / cast (C5 teacher inputs: 256 clips x 128 frames per expert)
for D, Dout in ((2048, 2048), (1024, 1024), (128, 2048)):
    xr = [torch.randn(B * 128, D, device=dev, generator=g) for _ in range(R)]
    ms = timeit(lambda i: ops.feature_augment(xr[i % R], Dout, p_drop=0.3, p_noise=0.3, seed=5 + i))
    report(f"feature_augment [{B * 128},{D}]->{Dout} f32->bf16", ms, bytes_=B * 128 * (D * 4.0 + Dout * 2.0))
del xr

# ---- spatial pyramid pooling (TPN.py:2-40): C3's 128 clips x 64 frames = 8192 frames of (128,28,28) + (256,14,14) + (512,7,7)
frames = 8192 if "--small" not in sys.argv else 2048
for Cc, s_ in ((128, 28), (256, 14), (512, 7)):
    for dt in (torch.bfloat16, torch.float32):
        maps = [torch.randn(frames, Cc, s_, s_, device=dev, generator=g).to(dt) for _ in range(2)]
        pooled = torch.empty(frames, Cc, device=dev)
        ms = timeit(lambda i: ops.spatial_pool(maps[i % 2], pooled, 0))
        nm = "bf16" if dt == torch.bfloat16 else "f32"
        report(f"spatial_pool_fwd [{frames},{Cc},{s_},{s_}] {nm}", ms, bytes_=frames * Cc * (s_ * s_ * maps[0].element_size() + 4.0))
        ms = timeit(lambda i: ops.spatial_pool_bwd(pooled, maps[i % 2], 0))
        report(f"spatial_pool_bwd [{frames},{Cc},{s_},{s_}] {nm}", ms, bytes_=frames * Cc * (s_ * s_ * maps[0].element_size() + 4.0))
        del maps
if "--only-bandwidth" in sys.argv:
    sys.exit(0)

# ---- GEMMs of one encoder layer (forward, dgrad, wgrad)
mode = ops.Mode("bf16")
def gemm_case(name, M, N, K, **kw):
    a = [bf(*( (K, M) if kw.get("a_mn") else (M, K))) for _ in range(R)]
    b = [bf(*((K, N) if kw.get("b_mn") else (N, K))) for _ in range(2)]
    splits = kw.pop("splits", 1)
    if splits > 1:
        outf = torch.zeros(M, N, device=dev)
        fn = lambda i: ops.gemm(a[i % R], b[i % 2], M, N, K, out_f32=outf, splits=splits, atomic=True, **kw)
    else:
        outb = [torch.empty(M, N, dtype=torch.bfloat16, device=dev) for _ in range(2)]
        extra = {}
        if kw.pop("fused", False):
            extra = dict(bias=torch.zeros(N, device=dev), residual=bf(M, N))
        if kw.pop("relu", False):
            extra = dict(bias=torch.zeros(N, device=dev), act=ops.ACT_RELU, dropout_p=0.5, seed=3)
        fn = lambda i: ops.gemm(a[i % R], b[i % 2], M, N, K, out_bf16=outb[i % 2], **extra, **kw)
    report(name, timeit(fn), flops=2.0 * M * N * K)

gemm_case("gemm fwd qkv      33024x2304x768", n, 3 * d, d)
gemm_case("gemm fwd out+res  33024x768x768", n, d, d, fused=True)
gemm_case("gemm fwd ffn1+relu+drop 33024x3072x768", n, ff, d, relu=True)
gemm_case("gemm fwd ffn2+res 33024x768x3072", n, d, ff, fused=True)
gemm_case("gemm dgrad ffn2   33024x3072x768", n, ff, d, b_mn=True)
gemm_case("gemm dgrad ffn1   33024x768x3072", n, d, ff, b_mn=True)
gemm_case("gemm wgrad ffn1   3072x768x33024", ff, d, n, a_mn=True, b_mn=True, splits=ops.pick_splits(24 * 3, 516))
gemm_case("gemm wgrad qkv    2304x768x33024", 3 * d, d, n, a_mn=True, b_mn=True, splits=ops.pick_splits(18 * 3, 516))
gemm_case("gemm wgrad ffn2   768x3072x33024", d, ff, n, a_mn=True, b_mn=True, splits=ops.pick_splits(6 * 12, 516))
gemm_case("gemm wgrad out    768x768x33024", d, d, n, a_mn=True, b_mn=True, splits=ops.pick_splits(6 * 3, 516))
gemm_case("gemm inproj rgb   32768x768x2048", B * 128, d, 2048)
