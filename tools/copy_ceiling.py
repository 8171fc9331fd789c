"""What a plain device copy / torch LayerNorm achieve at the LayerNorm shape (the practical bandwidth ceiling for a 100 MB pass)."""
import torch
n, d, R = 33024, 768, 4
xs = [torch.randn(n, d, device="cuda").bfloat16() for _ in range(R)]
ys = [torch.empty_like(xs[0]) for _ in range(R)]
def timeit(fn, iters=40):
    for i in range(5): fn(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for i in range(iters): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
t = timeit(lambda i: ys[i % R].copy_(xs[i % R]))
print(f"copy [33024,768] bf16: {t:.1f} us  {2 * n * d * 2 / t / 1e3:.0f} GB/s")
ln = torch.nn.LayerNorm(d, device="cuda", dtype=torch.bfloat16)
with torch.no_grad():
    t = timeit(lambda i: ln(xs[i % R]))
print(f"torch LayerNorm fwd: {t:.1f} us  {2 * n * d * 2 / t / 1e3:.0f} GB/s")
big = [torch.empty(256 * 1024 * 1024, device="cuda", dtype=torch.uint8) for _ in range(2)]
t = timeit(lambda i: big[1].copy_(big[0]), iters=10)
print(f"copy 256 MiB: {t:.1f} us  {2 * 256 * 1.048576 / t * 1e3:.0f} GB/s")
