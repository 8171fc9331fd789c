"""Library yardstick for the GEMM shapes of one C5 encoder layer: torch.matmul (cuBLAS) and F.linear with bias
(cuBLASLt epilogue) in bf16 on the same B200, inputs rotated out of L2.  Informational."""
import torch, torch.nn.functional as F
dev = "cuda:0"
n, d, ff = 33024, 768, 3072
g = torch.Generator(device=dev).manual_seed(0)
def bf(*s): return torch.randn(*s, device=dev, generator=g).to(torch.bfloat16)
def timeit(fn, iters=20):
    for _ in range(3): fn(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for i in range(iters): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for name, M, N, K in (("qkv", n, 3 * d, d), ("out", n, d, d), ("ffn1", n, ff, d), ("ffn2", n, d, ff)):
    xs = [bf(M, K) for _ in range(4)]; w = bf(N, K); b = bf(N)
    ms = timeit(lambda i: torch.matmul(xs[i % 4], w.t()))
    ms2 = timeit(lambda i: F.linear(xs[i % 4], w, b))
    ms3 = timeit(lambda i: F.relu(F.linear(xs[i % 4], w, b)))
    fl = 2.0 * M * N * K
    print(f"{name:5s} {M}x{N}x{K}: matmul {ms*1e3:7.1f} us {fl/ms/1e9:7.1f} TF/s | linear+bias {ms2*1e3:7.1f} us {fl/ms2/1e9:7.1f} TF/s | +relu (2 kernels) {ms3*1e3:7.1f} us")
for name, M, N, K in (("wgrad ffn1", ff, d, n), ("wgrad qkv", 3 * d, d, n), ("wgrad out", d, d, n)):
    dy = [bf(K, M) for _ in range(2)]; x = [bf(K, N) for _ in range(2)]
    ms = timeit(lambda i: torch.matmul(dy[i % 2].t(), x[i % 2]))
    print(f"{name:10s} {M}x{N}x{K}: matmul {ms*1e3:7.1f} us {2.0*M*N*K/ms/1e9:7.1f} TF/s")
