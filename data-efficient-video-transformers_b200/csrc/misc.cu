// Small bandwidth / latency kernels on the hot path (see include/tvt.h): column sums for bias
// gradients, fp32 -> bf16 (hi, lo) splitting, bias+activation after split-K GEMMs, the skinny class-head
// linear layer and the CLS gather + expert sum of SimpleTransformer.ptn (src/models/transformer.py:123-130).
#include "tvt_common.cuh"

namespace tvt {
namespace misc {

// ---------------------------------------------------------------- colsum
// grid (col tiles of 32*V, row slabs); each thread owns V columns and strides over the rows of its slab.
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* x, float* out, long long rows, long long cols, long long ld, long long rows_per_slab) {
  constexpr int V = Vec16<T>::kN;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long col = (blockIdx.x * 32ll + lane) * V;
  const long long r0 = blockIdx.y * rows_per_slab, r1 = r0 + rows_per_slab < rows ? r0 + rows_per_slab : rows;
  float acc[V];
#pragma unroll
  for (int i = 0; i < V; ++i) acc[i] = 0.0f;
  if (col < cols) {
    for (long long r = r0 + warp; r < r1; r += 8) {
      float v[V];
      Vec16<T>::load(x + r * ld + col, v);
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] += v[i];
    }
  }
  __shared__ float red[8][32 * V + 1];
#pragma unroll
  for (int i = 0; i < V; ++i) red[warp][lane * V + i] = acc[i];
  __syncthreads();
  for (int t = threadIdx.x; t < 32 * V; t += 256) {
    float s = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][t];
    const long long c = blockIdx.x * 32ll * V + t;
    if (c < cols) atomicAdd(out + c, s);
  }
}

// ---------------------------------------------------------------- split
// Three-plane split for the EXACT forward GEMMs of the fp32 parity mode: x = x0 + x1 + x2 with bf16 planes captures all 24
// mantissa bits, and the six products of order <= 2 (x0w0 + x0w1 + x1w0 + x0w2 + x2w0 + x1w1) reproduce an fp32 GEMM to
// ~2^-24.  They are obtained from the ordinary 3-pass kernel (a*b + a*b_lo + a_lo*b) by concatenating along K:
//   A  = [x0 | x0 | x2 | x1],  A_lo = [x1 | 0 | 0 | 0];   B = [w0 | w2 | w0 | w1],  B_lo = [w1 | 0 | 0 | 0]   (K' = 4K).
// Why it matters: with two planes (2^-17) a pre-activation within ~1e-5 of zero can land on the other side of a ReLU, and
// ONE flipped gate in a sparsely driven layer (only the CLS rows carry gradient behind the read-out) is a > 1e-3
// relative error of that layer's weight gradient.
__global__ void __launch_bounds__(256) split3_kernel(const float* x, __nv_bfloat16* hi4, __nv_bfloat16* lo4, long long rows, long long cols,
                                                      long long ld, int operand) {
  const long long vpr = cols / 8, total = rows * vpr;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = idx / vpr, c = (idx - r * vpr) * 8;
    const float* src = x + r * ld + c;
    const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    float p0[8], p1[8], p2[8], z[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      p0[j] = __bfloat162float(__float2bfloat16_rn(v[j]));
      const float r1 = v[j] - p0[j];                           // exact
      p1[j] = __bfloat162float(__float2bfloat16_rn(r1));
      p2[j] = r1 - p1[j];                                      // exact; rounded to bf16 by the store
      z[j] = 0.0f;
    }
    __nv_bfloat16* h = hi4 + r * 4 * cols + c;
    __nv_bfloat16* l = lo4 + r * 4 * cols + c;
    Vec16<__nv_bfloat16>::store(h, p0);
    Vec16<__nv_bfloat16>::store(h + cols, operand == 0 ? p0 : p2);
    Vec16<__nv_bfloat16>::store(h + 2 * cols, operand == 0 ? p2 : p0);
    Vec16<__nv_bfloat16>::store(h + 3 * cols, p1);
    Vec16<__nv_bfloat16>::store(l, p1);
    Vec16<__nv_bfloat16>::store(l + cols, z);
    Vec16<__nv_bfloat16>::store(l + 2 * cols, z);
    Vec16<__nv_bfloat16>::store(l + 3 * cols, z);
  }
}

__global__ void __launch_bounds__(256) split_kernel(const float* x, __nv_bfloat16* hi, __nv_bfloat16* lo, long long n) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 8;
  for (long long i = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) * 8; i < n; i += stride) {
    if (i + 8 <= n) {
      float v[8], h[8], l[8];
      const float4 a = *reinterpret_cast<const float4*>(x + i), b = *reinterpret_cast<const float4*>(x + i + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
#pragma unroll
      for (int j = 0; j < 8; ++j) { h[j] = __bfloat162float(__float2bfloat16_rn(v[j])); l[j] = v[j] - h[j]; }
      Vec16<__nv_bfloat16>::store(hi + i, h);
      if (lo) Vec16<__nv_bfloat16>::store(lo + i, l);
    } else {
      for (long long j = i; j < n; ++j) {
        const __nv_bfloat16 h = __float2bfloat16_rn(x[j]);
        hi[j] = h;
        if (lo) lo[j] = __float2bfloat16_rn(x[j] - __bfloat162float(h));
      }
    }
  }
}

// ---------------------------------------------------------------- bias + act (+dropout) (+residual)
// y = dropout(act(x + bias)) + residual, optional pre-activation store: the epilogue of a split-K GEMM (whose partial sums
// met in fp32 atomics and could not apply it) — Reasoning's skinny first layer, and every forward GEMM of the fp32 parity mode
// (short tensor-core accumulation chains, see ops.Mode.linear_fwd).
template <typename T>
__global__ void __launch_bounds__(256) bias_act_kernel(const float* x, const float* bias, T* y, const T* residual, T* preact, long long rows,
                                                       long long cols, int act, float dscale, unsigned thr16, unsigned long long seed,
                                                       const unsigned long long* seed_src) {
  seed = mix_seed(seed, seed_src);
  const long long n = rows * cols;
  for (long long i = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) * 4; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x * 4) {
    const float4 a = *reinterpret_cast<const float4*>(x + i);
    float v[4] = {a.x, a.y, a.z, a.w};
    const long long c = i % cols;
    const uint64_t bits = thr16 ? dropout_bits4(seed, static_cast<unsigned long long>(i) >> 2) : 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (bias) v[j] += bias[c + j];
      if (preact) preact[i + j] = Elem<T>::from_f(v[j]);
      if (act == TVT_ACT_RELU) v[j] = fmaxf(v[j], 0.0f);
      else if (act == TVT_ACT_GELU) v[j] = gelu_f(v[j]);
      if (thr16) v[j] = dropout_keep_lane(bits, j, thr16) ? v[j] * dscale : 0.0f;
      if (residual) v[j] += Elem<T>::to_f(residual[i + j]);
      y[i + j] = Elem<T>::from_f(v[j]);
    }
  }
}

// ---------------------------------------------------------------- positional encoding
template <typename T>
__global__ void __launch_bounds__(256) posenc_kernel(const T* x, const float* pe, T* y, long long rows, int d, int S, float dscale,
                                                     unsigned thr16, unsigned long long seed, const unsigned long long* seed_src) {
  seed = mix_seed(seed, seed_src);
  constexpr int V = Vec16<T>::kN;
  const long long n = rows * d;
  for (long long i = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) * V; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x * V) {
    const long long row = i / d;
    const int col = static_cast<int>(i - row * d);
    const float* pr = pe + (row % S) * d + col;
    float v[V];
    Vec16<T>::load(x + i, v);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      v[j] += pr[j];
      if (thr16) v[j] = dropout_keep(seed, static_cast<unsigned long long>(i + j), thr16) ? v[j] * dscale : 0.0f;
    }
    Vec16<T>::store(y + i, v);
  }
}

// ---------------------------------------------------------------- activation backward
template <typename T>
__global__ void __launch_bounds__(256) act_bwd_kernel(const T* dy, const T* yz, T* dx, long long n, int act, float dscale, unsigned thr16,
                                                      unsigned long long seed, const unsigned long long* seed_src) {
  seed = mix_seed(seed, seed_src);
  constexpr int V = Vec16<T>::kN;
  for (long long i = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) * V; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x * V) {
    float g[V], r[V];
    Vec16<T>::load(dy + i, g);
    Vec16<T>::load(yz + i, r);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      if (act == TVT_ACT_RELU) g[j] = r[j] > 0.0f ? g[j] : 0.0f;
      else if (act == TVT_ACT_GELU) g[j] *= gelu_grad_f(r[j]);
      if (thr16) g[j] = dropout_keep(seed, static_cast<unsigned long long>(i + j), thr16) ? g[j] * dscale : 0.0f;
    }
    Vec16<T>::store(dx + i, g);
  }
}

// ---------------------------------------------------------------- optimizer step on a flat bucket
struct OptParams {
  float* p; const float* g; float* m; float* v; __nv_bfloat16* hi; __nv_bfloat16* lo; long long n; int kind; int first;
  float lr, b1, b2, eps, wd, mom, gs, bc1_inv, bc2_rsqrt;
  const long long* step_dev;   // optional device-resident step count (whole-step CUDA graphs): overrides first / bc1_inv / bc2_rsqrt
};
__global__ void __launch_bounds__(256) optim_kernel(OptParams a) {
  if (a.step_dev != nullptr) {
    const float t = static_cast<float>(__ldg(a.step_dev));
    a.first = t <= 1.0f;
    a.bc1_inv = 1.0f / (1.0f - powf(a.b1, t));
    a.bc2_rsqrt = 1.0f / sqrtf(1.0f - powf(a.b2, t));
  }
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 4;
  for (long long i = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) * 4; i < a.n; i += stride) {
    float p[4], g[4], m[4], v[4] = {0.f, 0.f, 0.f, 0.f};
    const bool full = i + 4 <= a.n;
    if (full) {
      const float4 p4 = *reinterpret_cast<const float4*>(a.p + i), g4 = *reinterpret_cast<const float4*>(a.g + i);
      const float4 m4 = *reinterpret_cast<const float4*>(a.m + i);
      p[0] = p4.x; p[1] = p4.y; p[2] = p4.z; p[3] = p4.w; g[0] = g4.x; g[1] = g4.y; g[2] = g4.z; g[3] = g4.w;
      m[0] = m4.x; m[1] = m4.y; m[2] = m4.z; m[3] = m4.w;
      if (a.kind != 1) { const float4 v4 = *reinterpret_cast<const float4*>(a.v + i); v[0] = v4.x; v[1] = v4.y; v[2] = v4.z; v[3] = v4.w; }
    } else {
      for (int j = 0; j < 4; ++j) {
        const bool ok = i + j < a.n;
        p[j] = ok ? a.p[i + j] : 0.f; g[j] = ok ? a.g[i + j] : 0.f; m[j] = ok ? a.m[i + j] : 0.f;
        v[j] = ok && a.kind != 1 ? a.v[i + j] : 0.f;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float gj = g[j] * a.gs;
      if (a.kind == 0) {
        p[j] *= 1.0f - a.lr * a.wd;
        m[j] = a.b1 * m[j] + (1.0f - a.b1) * gj;
        v[j] = a.b2 * v[j] + (1.0f - a.b2) * gj * gj;
        const float denom = sqrtf(v[j]) * a.bc2_rsqrt + a.eps;
        p[j] -= a.lr * a.bc1_inv * (m[j] / denom);
      } else if (a.kind == 1) {
        gj += a.wd * p[j];
        m[j] = a.first ? gj : a.mom * m[j] + gj;
        p[j] -= a.lr * m[j];
      } else {   // Adagrad (torch.optim.Adagrad, lr_decay = 0, initial accumulator 0): v is the running sum of g^2
        gj += a.wd * p[j];
        v[j] += gj * gj;
        p[j] -= a.lr * gj / (sqrtf(v[j]) + a.eps);
      }
    }
    if (full) {
      *reinterpret_cast<float4*>(a.p + i) = make_float4(p[0], p[1], p[2], p[3]);
      *reinterpret_cast<float4*>(a.m + i) = make_float4(m[0], m[1], m[2], m[3]);
      if (a.kind != 1) *reinterpret_cast<float4*>(a.v + i) = make_float4(v[0], v[1], v[2], v[3]);
      if (a.hi) {
        float h[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = __bfloat162float(__float2bfloat16_rn(p[j]));
        *reinterpret_cast<uint2*>(a.hi + i) = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
        if (a.lo) *reinterpret_cast<uint2*>(a.lo + i) = make_uint2(pack_bf16x2(p[0] - h[0], p[1] - h[1]), pack_bf16x2(p[2] - h[2], p[3] - h[3]));
      }
    } else {
      for (int j = 0; j < 4 && i + j < a.n; ++j) {
        a.p[i + j] = p[j]; a.m[i + j] = m[j];
        if (a.kind != 1) a.v[i + j] = v[j];
        if (a.hi) {
          const __nv_bfloat16 h = __float2bfloat16_rn(p[j]);
          a.hi[i + j] = h;
          if (a.lo) a.lo[i + j] = __float2bfloat16_rn(p[j] - __bfloat162float(h));
        }
      }
    }
  }
}

// ---------------------------------------------------------------- class-head linear
// one warp per row; eight classes per pass share each x element, so the row is read ceil(C / 8) times from L1 and the eight
// dot products give the FMA pipe and the final shuffle trees independent work (the one-class-at-a-time version was a
// 15 x (24 loads + 5 shuffles) dependent chain: 64 us for 256 x 768 -> 15)
template <typename T>
__global__ void __launch_bounds__(128) head_fwd_kernel(const T* x, const float* w, const float* b, float* y, long long M, long long K, int C) {
  const int lane = threadIdx.x & 31;
  const long long m = blockIdx.x * 4ll + (threadIdx.x >> 5);
  if (m >= M) return;
  const T* xr = x + m * K;
  for (int c0 = 0; c0 < C; c0 += 8) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
    // class indices beyond C are clamped (their sums are discarded): unconditional loads, so the compiler batches the
    // nine loads of an iteration and overlaps iterations instead of serialising load -> FMA -> load
    const float* wr[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) wr[j] = w + static_cast<long long>(c0 + j < C ? c0 + j : C - 1) * K;
#pragma unroll 4
    for (long long k = lane; k < K; k += 32) {
      const float xv = Elem<T>::to_f(xr[k]);
      float wv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) wv[j] = __ldg(wr[j] + k);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(xv, wv[j], acc[j]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
    }
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (c0 + j < C) y[m * C + c0 + j] = acc[j] + (b ? b[c0 + j] : 0.0f);
    }
  }
}
// dx[m,k] = sum_c dy[m,c] w[c,k]
template <typename T>
__global__ void __launch_bounds__(256) head_dx_kernel(const float* dy, const float* w, T* dx, long long M, long long K, int C) {
  const long long n = M * K;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m = i / K, k = i - m * K;
    float acc = 0.0f;
#pragma unroll 5
    for (int c = 0; c < C; ++c) acc = fmaf(__ldg(dy + m * C + c), __ldg(w + c * K + k), acc);
    dx[i] = Elem<T>::from_f(acc);
  }
}
// dw[c,k] += sum_m dy[m,c] x[m,k];  db[c] += sum_m dy[m,c].  Thread == input column k, eight classes per pass in registers,
// blockIdx.y splits the rows into segments of kHeadRows whose partial sums meet in fp32 atomics (the one-owner-per-(c, k)
// version walked all M rows serially with one FMA in flight: 55 us for 256 x 768 -> 15)
constexpr int kHeadRows = 32;
template <typename T>
__global__ void __launch_bounds__(256) head_dw_kernel(const float* dy, const T* x, float* dw, float* db, long long M, long long K, int C) {
  const long long k = blockIdx.x * 256ll + threadIdx.x;
  const long long m0 = static_cast<long long>(blockIdx.y) * kHeadRows;
  const long long m1 = m0 + kHeadRows < M ? m0 + kHeadRows : M;
  if (k >= K) return;
  for (int c0 = 0; c0 < C; c0 += 8) {
    float acc[8], accb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = accb[j] = 0.0f;
    int cj[8];   // clamped class indices: unconditional (batched) loads, surplus sums are discarded
#pragma unroll
    for (int j = 0; j < 8; ++j) cj[j] = c0 + j < C ? c0 + j : C - 1;
#pragma unroll 4
    for (long long m = m0; m < m1; ++m) {
      const float xv = Elem<T>::to_f(x[m * K + k]);
      float g[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = __ldg(dy + m * C + cj[j]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[j] = fmaf(g[j], xv, acc[j]);
        accb[j] += g[j];
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (c0 + j < C) {
        atomicAdd(dw + (c0 + j) * K + k, acc[j]);
        if (k == 0 && db) atomicAdd(db + c0 + j, accb[j]);
      }
    }
  }
}

// ---------------------------------------------------------------- CLS gather + expert sum
struct ClsParams { const void* tok[TVT_MAX_EXPERTS]; void* out; long long B, S, d; int E; };
template <typename T>
__global__ void __launch_bounds__(256) cls_sum_kernel(const ClsParams p) {
  constexpr int V = Vec16<T>::kN;
  const long long vecs = p.d / V, n = p.B * vecs;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long b = i / vecs;
    const int col = static_cast<int>(i - b * vecs) * V;
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = 0.0f;
    for (int e = 0; e < p.E; ++e) {
      float v[V];
      Vec16<T>::load(reinterpret_cast<const T*>(p.tok[e]) + b * p.S * p.d + col, v);
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] += v[j];
    }
    Vec16<T>::store(reinterpret_cast<T*>(p.out) + b * p.d + col, acc);
  }
}

static int grid1d(long long items, int threads) {
  const long long want = (items + threads - 1) / threads;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  return static_cast<int>(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace misc
}  // namespace tvt

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int tvt_colsum(const tvt_colsum_args* a, void* stream) {
  using namespace tvt;
  TVT_REQUIRE(a != nullptr && a->x && a->out, "tvt_colsum: null pointer");
  TVT_REQUIRE(a->rows >= 0 && a->cols > 0 && a->cols % 8 == 0 && a->ld >= a->cols && a->ld % 8 == 0, "tvt_colsum: cols and ld must be multiples of 8");
  TVT_REQUIRE(a->dtype == TVT_BF16 || a->dtype == TVT_F32, "tvt_colsum: bad dtype");
  TVT_REQUIRE(al16(a->x), "tvt_colsum: x must be 16-byte aligned");
  if (a->rows == 0) return TVT_OK;
  int rc = require_sm100();
  if (rc != TVT_OK) return rc;
  const int V = a->dtype == TVT_F32 ? 4 : 8;
  const long long col_tiles = (a->cols + 32 * V - 1) / (32 * V);
  long long slabs = (static_cast<long long>(num_sms()) * 4 + col_tiles - 1) / col_tiles;
  const long long max_slabs = (a->rows + 63) / 64;
  if (slabs > max_slabs) slabs = max_slabs;
  if (slabs < 1) slabs = 1;
  const long long rps = (a->rows + slabs - 1) / slabs;
  dim3 grid(static_cast<unsigned>(col_tiles), static_cast<unsigned>((a->rows + rps - 1) / rps));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (a->dtype == TVT_F32) misc::colsum_kernel<float><<<grid, 256, 0, s>>>((const float*)a->x, a->out, a->rows, a->cols, a->ld, rps);
  else misc::colsum_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)a->x, a->out, a->rows, a->cols, a->ld, rps);
  return check_launch("tvt_colsum");
}

extern "C" int tvt_split_f32(const tvt_split_args* a, void* stream) {
  using namespace tvt;
  TVT_REQUIRE(a != nullptr && a->x && a->hi, "tvt_split_f32: null pointer");
  TVT_REQUIRE(a->n >= 0, "tvt_split_f32: negative size");
  TVT_REQUIRE(al16(a->x) && al16(a->hi) && al16(a->lo), "tvt_split_f32: pointers must be 16-byte aligned");
  if (a->n == 0) return TVT_OK;
  int rc = require_sm100();
  if (rc != TVT_OK) return rc;
  misc::split_kernel<<<misc::grid1d((a->n + 7) / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      a->x, (__nv_bfloat16*)a->hi, (__nv_bfloat16*)a->lo, a->n);
  return check_launch("tvt_split_f32");
}

extern "C" int tvt_split_f32x3(const tvt_split3_args* a, void* stream) {
  using namespace tvt;
  TVT_REQUIRE(a != nullptr && a->x && a->hi4 && a->lo4, "tvt_split_f32x3: null pointer");
  TVT_REQUIRE(a->rows > 0 && a->cols > 0 && a->cols % 8 == 0, "tvt_split_f32x3: cols must be a positive multiple of 8");
  TVT_REQUIRE(a->ld >= a->cols, "tvt_split_f32x3: ld smaller than cols");
  TVT_REQUIRE(a->operand == 0 || a->operand == 1, "tvt_split_f32x3: operand must be 0 (A) or 1 (B)");
  TVT_REQUIRE(al16(a->x) && al16(a->hi4) && al16(a->lo4) && a->ld % 4 == 0, "tvt_split_f32x3: pointers / pitch must keep 16-byte alignment");
  int rc = require_sm100();
  if (rc != TVT_OK) return rc;
  const long long vecs = a->rows * (a->cols / 8);
  misc::split3_kernel<<<misc::grid1d(vecs, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      a->x, (__nv_bfloat16*)a->hi4, (__nv_bfloat16*)a->lo4, a->rows, a->cols, a->ld, a->operand);
  return check_launch("tvt_split_f32x3");
}

extern "C" int tvt_bias_act_fwd(const tvt_bias_act_args* a, void* stream) {
  using namespace tvt;
  TVT_REQUIRE(a != nullptr && a->x && a->y, "tvt_bias_act_fwd: null pointer");
  TVT_REQUIRE(a->rows >= 0 && a->cols > 0 && a->cols % 4 == 0, "tvt_bias_act_fwd: cols must be a multiple of 4");
  TVT_REQUIRE(a->out_dtype == TVT_BF16 || a->out_dtype == TVT_F32, "tvt_bias_act_fwd: bad dtype");
  TVT_REQUIRE(a->act >= TVT_ACT_NONE && a->act <= TVT_ACT_GELU, "tvt_bias_act_fwd: bad act");
  TVT_REQUIRE(a->dropout_p >= 0.0f && a->dropout_p < 1.0f, "tvt_bias_act_fwd: dropout_p must be in [0,1)");
  TVT_REQUIRE(al16(a->x), "tvt_bias_act_fwd: x must be 16-byte aligned");
  if (a->rows == 0) return TVT_OK;
  int rc = require_sm100();
  if (rc != TVT_OK) return rc;
  unsigned thr = 0; float sc = 1.0f;
  if (a->dropout_p > 0.0f) { thr = (unsigned)(a->dropout_p * 65536.0f + 0.5f); sc = 65536.0f / (65536.0f - (float)thr); }
  const long long n4 = a->rows * a->cols / 4;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (a->out_dtype == TVT_F32) misc::bias_act_kernel<float><<<misc::grid1d(n4, 256), 256, 0, s>>>(a->x, a->bias, (float*)a->y, (const float*)a->residual, (float*)a->preact, a->rows, a->cols, a->act, sc, thr, a->dropout_seed, seed_source());
  else misc::bias_act_kernel<__nv_bfloat16><<<misc::grid1d(n4, 256), 256, 0, s>>>(a->x, a->bias, (__nv_bfloat16*)a->y, (const __nv_bfloat16*)a->residual, (__nv_bfloat16*)a->preact, a->rows, a->cols, a->act, sc, thr, a->dropout_seed, seed_source());
  return check_launch("tvt_bias_act_fwd");
}

extern "C" int tvt_posenc_fwd(const tvt_posenc_args* a, void* stream) {
  using namespace tvt;
  TVT_REQUIRE(a != nullptr && a->x && a->pe && a->y, "tvt_posenc_fwd: null pointer");
  TVT_REQUIRE(a->rows >= 0 && a->d > 0 && a->d % 8 == 0 && a->seq_len > 0, "tvt_posenc_fwd: bad shape (d must be a multiple of 8)");
  TVT_REQUIRE(a->rows % a->seq_len == 0, "tvt_posenc_fwd: rows must be a multiple of seq_len");
  TVT_REQUIRE(a->dtype == TVT_BF16 || a->dtype == TVT_F32, "tvt_posenc_fwd: bad dtype");
  TVT_REQUIRE(a->dropout_p >= 0.0f && a->dropout_p < 1.0f, "tvt_posenc_fwd: dropout_p must be in [0,1)");
  TVT_REQUIRE(al16(a->x) && al16(a->y), "tvt_posenc_fwd: pointers must be 16-byte aligned");
  if (a->rows == 0) return TVT_OK;
  int rc = require_sm100();
  if (rc != TVT_OK) return rc;
  unsigned thr = 0; float sc = 1.0f;
  if (a->dropout_p > 0.0f) { thr = (unsigned)(a->dropout_p * 65536.0f + 0.5f); sc = 65536.0f / (65536.0f - (float)thr); }
  const long long n = a->rows * a->d;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (a->dtype == TVT_F32) misc::posenc_kernel<float><<<misc::grid1d(n / 4, 256), 256, 0, s>>>((const float*)a->x, a->pe, (float*)a->y, a->rows, (int)a->d, (int)a->seq_len, sc, thr, a->dropout_seed, seed_source());
  else misc::posenc_kernel<__nv_bfloat16><<<misc::grid1d(n / 8, 256), 256, 0, s>>>((const __nv_bfloat16*)a->x, a->pe, (__nv_bfloat16*)a->y, a->rows, (int)a->d, (int)a->seq_len, sc, thr, a->dropout_seed, seed_source());
  return check_launch("tvt_posenc_fwd");
}

extern "C" int tvt_act_bwd(const tvt_act_bwd_args* a, void* stream) {
  using namespace tvt;
  TVT_REQUIRE(a != nullptr && a->dy && a->y_or_z && a->dx, "tvt_act_bwd: null pointer");
  TVT_REQUIRE(a->rows >= 0 && a->cols > 0 && (a->rows * a->cols) % 8 == 0, "tvt_act_bwd: element count must be a multiple of 8");
  TVT_REQUIRE(a->dtype == TVT_BF16 || a->dtype == TVT_F32, "tvt_act_bwd: bad dtype");
  TVT_REQUIRE(a->act >= TVT_ACT_NONE && a->act <= TVT_ACT_GELU, "tvt_act_bwd: bad act");
  TVT_REQUIRE(a->dropout_p >= 0.0f && a->dropout_p < 1.0f, "tvt_act_bwd: dropout_p must be in [0,1)");
  TVT_REQUIRE(al16(a->dy) && al16(a->y_or_z) && al16(a->dx), "tvt_act_bwd: pointers must be 16-byte aligned");
  if (a->rows == 0) return TVT_OK;
  int rc = require_sm100();
  if (rc != TVT_OK) return rc;
  unsigned thr = 0; float sc = 1.0f;
  if (a->dropout_p > 0.0f) { thr = (unsigned)(a->dropout_p * 65536.0f + 0.5f); sc = 65536.0f / (65536.0f - (float)thr); }
  const long long n = a->rows * a->cols;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (a->dtype == TVT_F32) misc::act_bwd_kernel<float><<<misc::grid1d(n / 4, 256), 256, 0, s>>>((const float*)a->dy, (const float*)a->y_or_z, (float*)a->dx, n, a->act, sc, thr, a->dropout_seed, seed_source());
  else misc::act_bwd_kernel<__nv_bfloat16><<<misc::grid1d(n / 8, 256), 256, 0, s>>>((const __nv_bfloat16*)a->dy, (const __nv_bfloat16*)a->y_or_z, (__nv_bfloat16*)a->dx, n, a->act, sc, thr, a->dropout_seed, seed_source());
  return check_launch("tvt_act_bwd");
}

extern "C" int tvt_optim_step(const tvt_optim_step_args* a, void* stream) {
  using namespace tvt;
  TVT_REQUIRE(a != nullptr && a->p && a->g && a->m, "tvt_optim_step: null pointer");
  TVT_REQUIRE(a->kind >= 0 && a->kind <= 2, "tvt_optim_step: kind must be 0 (AdamW), 1 (SGD) or 2 (Adagrad)");
  TVT_REQUIRE(a->kind == 1 || a->v, "tvt_optim_step: AdamW / Adagrad need the second-moment buffer");
  TVT_REQUIRE(a->n >= 0 && (a->step >= 1 || a->step_dev), "tvt_optim_step: bad n / step");
  TVT_REQUIRE(al16(a->p) && al16(a->g) && al16(a->m) && al16(a->v) && (reinterpret_cast<uintptr_t>(a->p_hi) & 7) == 0 &&
                  (reinterpret_cast<uintptr_t>(a->p_lo) & 7) == 0,
              "tvt_optim_step: buffers must be 16-byte aligned (bf16 planes 8-byte)");
  TVT_REQUIRE(!a->p_lo || a->p_hi, "tvt_optim_step: p_lo without p_hi");
  if (a->n == 0) return TVT_OK;
  int rc = require_sm100();
  if (rc != TVT_OK) return rc;
  misc::OptParams q{};
  q.p = a->p; q.g = a->g; q.m = a->m; q.v = a->v; q.hi = (__nv_bfloat16*)a->p_hi; q.lo = (__nv_bfloat16*)a->p_lo; q.n = a->n;
  q.kind = a->kind; q.first = a->step == 1;
  q.lr = a->lr; q.b1 = a->beta1; q.b2 = a->beta2; q.eps = a->eps; q.wd = a->weight_decay; q.mom = a->momentum;
  q.gs = a->grad_scale == 0.0f ? 1.0f : a->grad_scale;
  const float t = (float)(a->step >= 1 ? a->step : 1);
  q.bc1_inv = 1.0f / (1.0f - powf(a->beta1, t));
  q.bc2_rsqrt = 1.0f / sqrtf(1.0f - powf(a->beta2, t));
  q.step_dev = reinterpret_cast<const long long*>(a->step_dev);
  misc::optim_kernel<<<misc::grid1d((a->n + 3) / 4, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(q);
  return check_launch("tvt_optim_step");
}

extern "C" int tvt_head_linear_fwd(const tvt_head_linear_fwd_args* a, void* stream) {
  using namespace tvt;
  TVT_REQUIRE(a != nullptr && a->x && a->w && a->y, "tvt_head_linear_fwd: null pointer");
  TVT_REQUIRE(a->m > 0 && a->k > 0 && a->classes > 0 && a->classes <= 1024, "tvt_head_linear_fwd: bad shape");
  TVT_REQUIRE(a->dtype == TVT_BF16 || a->dtype == TVT_F32, "tvt_head_linear_fwd: bad dtype");
  int rc = require_sm100();
  if (rc != TVT_OK) return rc;
  const int grid = static_cast<int>((a->m + 3) / 4);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (a->dtype == TVT_F32) misc::head_fwd_kernel<float><<<grid, 128, 0, s>>>((const float*)a->x, a->w, a->b, a->y, a->m, a->k, (int)a->classes);
  else misc::head_fwd_kernel<__nv_bfloat16><<<grid, 128, 0, s>>>((const __nv_bfloat16*)a->x, a->w, a->b, a->y, a->m, a->k, (int)a->classes);
  return check_launch("tvt_head_linear_fwd");
}

extern "C" int tvt_head_linear_bwd(const tvt_head_linear_bwd_args* a, void* stream) {
  using namespace tvt;
  TVT_REQUIRE(a != nullptr && a->dy && a->w && a->x, "tvt_head_linear_bwd: null pointer");
  TVT_REQUIRE(a->m > 0 && a->k > 0 && a->classes > 0 && a->classes <= 1024, "tvt_head_linear_bwd: bad shape");
  TVT_REQUIRE(a->dtype == TVT_BF16 || a->dtype == TVT_F32, "tvt_head_linear_bwd: bad dtype");
  int rc = require_sm100();
  if (rc != TVT_OK) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int C = (int)a->classes;
  if (a->dx) {
    const int g = misc::grid1d(a->m * a->k, 256);
    if (a->dtype == TVT_F32) misc::head_dx_kernel<float><<<g, 256, 0, s>>>(a->dy, a->w, (float*)a->dx, a->m, a->k, C);
    else misc::head_dx_kernel<__nv_bfloat16><<<g, 256, 0, s>>>(a->dy, a->w, (__nv_bfloat16*)a->dx, a->m, a->k, C);
    rc = check_launch("tvt_head_linear_bwd(dx)");
    if (rc != TVT_OK) return rc;
  }
  if (a->dw) {
    const dim3 g(static_cast<unsigned>((a->k + 255) / 256), static_cast<unsigned>((a->m + misc::kHeadRows - 1) / misc::kHeadRows));
    if (a->dtype == TVT_F32) misc::head_dw_kernel<float><<<g, 256, 0, s>>>(a->dy, (const float*)a->x, a->dw, a->db, a->m, a->k, C);
    else misc::head_dw_kernel<__nv_bfloat16><<<g, 256, 0, s>>>(a->dy, (const __nv_bfloat16*)a->x, a->dw, a->db, a->m, a->k, C);
    rc = check_launch("tvt_head_linear_bwd(dw)");
  }
  return rc;
}

extern "C" int tvt_cls_sum_fwd(const tvt_cls_sum_args* a, void* stream) {
  using namespace tvt;
  TVT_REQUIRE(a != nullptr && a->out, "tvt_cls_sum_fwd: null pointer");
  TVT_REQUIRE(a->num_experts >= 1 && a->num_experts <= TVT_MAX_EXPERTS, "tvt_cls_sum_fwd: num_experts out of range");
  TVT_REQUIRE(a->batch > 0 && a->seq_len > 0 && a->d > 0 && a->d % 8 == 0, "tvt_cls_sum_fwd: bad shape");
  TVT_REQUIRE(a->dtype == TVT_BF16 || a->dtype == TVT_F32, "tvt_cls_sum_fwd: bad dtype");
  misc::ClsParams p{};
  for (int e = 0; e < a->num_experts; ++e) {
    TVT_REQUIRE(a->tokens[e] && al16(a->tokens[e]), "tvt_cls_sum_fwd: bad token pointer");
    p.tok[e] = a->tokens[e];
  }
  int rc = require_sm100();
  if (rc != TVT_OK) return rc;
  p.out = a->out; p.B = a->batch; p.S = a->seq_len; p.d = a->d; p.E = a->num_experts;
  const long long items = a->batch * (a->d / (a->dtype == TVT_F32 ? 4 : 8));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (a->dtype == TVT_F32) misc::cls_sum_kernel<float><<<misc::grid1d(items, 256), 256, 0, s>>>(p);
  else misc::cls_sum_kernel<__nv_bfloat16><<<misc::grid1d(items, 256), 256, 0, s>>>(p);
  return check_launch("tvt_cls_sum_fwd");
}

// ---------------------------------------------------------------------------------------- evaluation read-out
namespace tvt {
namespace misc {
struct EvalParams {
  const float* logits; const void* target; float* probs; int32_t* labels; uint16_t* bits; int32_t* top1;
  long long B, C, row0; int nthr, tgt_f64; float thr[TVT_MAX_THRESHOLDS];
};
// one warp per clip: lanes stride over the classes, then a warp argmax (lowest index among equal maxima)
__global__ void __launch_bounds__(256) eval_readout_kernel(const EvalParams p) {
  const long long b = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (b >= p.B) return;
  const int lane = threadIdx.x & 31;
  const long long orow = (p.row0 + b) * p.C;
  float best = -INFINITY;
  int besti = 0x7fffffff;
  for (int c = lane; c < p.C; c += 32) {
    const float z = p.logits[b * p.C + c];
    const float pr = 1.0f / (1.0f + expf(-z));
    p.probs[orow + c] = pr;
    if (p.labels && p.target)
      p.labels[orow + c] = p.tgt_f64 ? static_cast<int32_t>(reinterpret_cast<const double*>(p.target)[b * p.C + c])
                                     : static_cast<int32_t>(reinterpret_cast<const float*>(p.target)[b * p.C + c]);
    if (p.bits) {
      unsigned m = 0;
      for (int k = 0; k < p.nthr; ++k) m |= (pr > p.thr[k] ? 1u : 0u) << k;
      p.bits[orow + c] = static_cast<uint16_t>(m);
    }
    if (z > best || (z != z && best == best)) { best = z; besti = c; }   // NaN counts as the maximum, like torch.argmax
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    const bool take = (ob > best) || (ob != ob && best == best) || (ob == best && oi < besti) || (ob != ob && best != best && oi < besti);
    if (take) { best = ob; besti = oi; }
  }
  if (lane == 0 && p.top1) p.top1[p.row0 + b] = besti;
}
}  // namespace misc
}  // namespace tvt

extern "C" int tvt_eval_readout(const tvt_eval_readout_args* a, void* stream) {
  using namespace tvt;
  TVT_REQUIRE(a != nullptr && a->logits && a->probs, "tvt_eval_readout: null pointer");
  TVT_REQUIRE(a->batch >= 0 && a->classes > 0, "tvt_eval_readout: bad shape");
  TVT_REQUIRE(a->row_offset >= 0 && a->row_offset + a->batch <= a->capacity,
              "tvt_eval_readout: rows [%lld, %lld) exceed the running buffer's capacity %lld", (long long)a->row_offset,
              (long long)(a->row_offset + a->batch), (long long)a->capacity);
  TVT_REQUIRE(a->num_thresholds >= 0 && a->num_thresholds <= TVT_MAX_THRESHOLDS, "tvt_eval_readout: too many thresholds");
  TVT_REQUIRE(!a->target || a->target_dtype == TVT_F32 || a->target_dtype == TVT_F64, "tvt_eval_readout: target must be f32 or f64");
  if (a->batch == 0) return TVT_OK;
  int rc = require_sm100();
  if (rc != TVT_OK) return rc;
  misc::EvalParams p{};
  p.logits = a->logits; p.target = a->target; p.probs = a->probs; p.labels = a->labels; p.bits = a->pred_bits; p.top1 = a->top1;
  p.B = a->batch; p.C = a->classes; p.row0 = a->row_offset; p.nthr = a->num_thresholds; p.tgt_f64 = a->target_dtype == TVT_F64;
  for (int k = 0; k < a->num_thresholds; ++k) p.thr[k] = a->thresholds[k];
  misc::eval_readout_kernel<<<static_cast<unsigned>((a->batch + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("tvt_eval_readout");
}

// ---------------------------------------------------------------------------------------- loader augmentation
namespace tvt {
namespace misc {
constexpr uint32_t kAugNoiseStream = 0x6E6F6973u;   // xor'ed into the seed's high word for the per-element stream
__device__ __forceinline__ float aug_u01(uint32_t w) { return (static_cast<float>(w >> 8) + 0.5f) * (1.0f / 16777216.0f); }   // (0, 1)

template <typename T>
__global__ void __launch_bounds__(256) feature_augment_kernel(const float* x, T* y, long long rows, int d_in, int d_out, float p_drop,
                                                              float p_noise, float noise_std, unsigned long long seed, const unsigned long long* seed_src) {
  seed = mix_seed(seed, seed_src);
  uint32_t rk[kDropoutRounds], rkn[kDropoutRounds];
#pragma unroll
  for (int r = 0; r < kDropoutRounds; ++r) rk[r] = rkn[r] = static_cast<uint32_t>(seed) + r * kDropoutWeyl;
  const uint32_t hi_row = static_cast<uint32_t>(seed >> 32), hi_noise = hi_row ^ kAugNoiseStream;
  const int q_per_row = d_out / 4;
  const long long total = rows * q_per_row;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / q_per_row;
    const int c = static_cast<int>(i - row * q_per_row) * 4;
    uint32_t w0, w1;
    dropout_words(rk, static_cast<uint32_t>(row), static_cast<uint32_t>(row >> 32) ^ hi_row, w0, w1);
    const bool drop = aug_u01(w0) < p_drop, noise = aug_u01(w1) < p_noise;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (!drop && c < d_in) {
      const float4 a = *reinterpret_cast<const float4*>(x + row * d_in + c);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    }
    if (noise) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {   // one hash -> one Box-Muller pair -> columns c + 2h, c + 2h + 1
        const unsigned long long ctr = static_cast<unsigned long long>(row) * (d_out / 2) + (c >> 1) + h;
        uint32_t a, b;
        dropout_words(rkn, static_cast<uint32_t>(ctr), static_cast<uint32_t>(ctr >> 32) ^ hi_noise, a, b);
        const float rad = sqrtf(-2.0f * logf(aug_u01(a))) * noise_std;
        float sn, cs;
        sincosf(6.283185307179586f * aug_u01(b), &sn, &cs);
        v[2 * h] += rad * cs;
        v[2 * h + 1] += rad * sn;
      }
    }
    if constexpr (sizeof(T) == 4) {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + row * d_out + c) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(y) + row * d_out + c) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
    }
  }
}
// Same function of (seed, row, column) as feature_augment_kernel, mapped WARP PER ROW for widths that are multiples of 8: the
// row's drop / noise decision is hashed once per row instead of once per four outputs, the 64-bit row division disappears, and
// every lane moves 16 bytes of bf16 output (32 of fp32) per iteration.  The thread-per-quad kernel spent ~80 instructions per
// 8 output bytes, which made the write-dominated pad case ([*, 128] -> 2048: 22 % of the copy peak) instruction-bound.
template <typename T>
__global__ void __launch_bounds__(256) feature_augment_rows_kernel(const float* x, T* y, long long rows, int d_in, int d_out, float p_drop,
                                                                   float p_noise, float noise_std, unsigned long long seed,
                                                                   const unsigned long long* seed_src) {
  seed = mix_seed(seed, seed_src);
  uint32_t rk[kDropoutRounds];
#pragma unroll
  for (int r = 0; r < kDropoutRounds; ++r) rk[r] = static_cast<uint32_t>(seed) + r * kDropoutWeyl;
  const uint32_t hi_row = static_cast<uint32_t>(seed >> 32), hi_noise = hi_row ^ kAugNoiseStream;
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long row = warp0; row < rows; row += nwarps) {
    uint32_t w0, w1;
    dropout_words(rk, static_cast<uint32_t>(row), static_cast<uint32_t>(row >> 32) ^ hi_row, w0, w1);
    const bool drop = aug_u01(w0) < p_drop, noise = aug_u01(w1) < p_noise;
    const float* xr = x + row * d_in;
    const unsigned long long ctr0 = static_cast<unsigned long long>(row) * (d_out / 2);
    for (int c = lane * 8; c < d_out; c += 256) {
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (!drop && c < d_in) {
        const float4 a = *reinterpret_cast<const float4*>(xr + c), b = *reinterpret_cast<const float4*>(xr + c + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      }
      if (noise) {
#pragma unroll
        for (int h = 0; h < 4; ++h) {   // one hash -> one Box-Muller pair -> columns c + 2h, c + 2h + 1
          const unsigned long long ctr = ctr0 + (c >> 1) + h;
          uint32_t a, b;
          dropout_words(rk, static_cast<uint32_t>(ctr), static_cast<uint32_t>(ctr >> 32) ^ hi_noise, a, b);
          const float rad = sqrtf(-2.0f * logf(aug_u01(a))) * noise_std;
          float sn, cs;
          sincosf(6.283185307179586f * aug_u01(b), &sn, &cs);
          v[2 * h] += rad * cs;
          v[2 * h + 1] += rad * sn;
        }
      }
      if constexpr (sizeof(T) == 4) {
        float* dst = reinterpret_cast<float*>(y) + row * d_out + c;
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
      } else {
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(y) + row * d_out + c) =
            make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
      }
    }
  }
}
}  // namespace misc
}  // namespace tvt

extern "C" int tvt_feature_augment(const tvt_feature_augment_args* a, void* stream) {
  using namespace tvt;
  TVT_REQUIRE(a != nullptr && a->x && a->y, "tvt_feature_augment: null pointer");
  TVT_REQUIRE(a->rows >= 0 && a->d_in > 0 && a->d_out >= a->d_in && a->d_in % 4 == 0 && a->d_out % 4 == 0,
              "tvt_feature_augment: need 0 < d_in <= d_out, both multiples of 4 (got %lld, %lld)", (long long)a->d_in, (long long)a->d_out);
  TVT_REQUIRE(a->d_out < (1ll << 31), "tvt_feature_augment: d_out exceeds int32");
  TVT_REQUIRE(a->p_drop >= 0.0f && a->p_drop <= 1.0f && a->p_noise >= 0.0f && a->p_noise <= 1.0f && a->noise_std >= 0.0f,
              "tvt_feature_augment: probabilities must be in [0, 1] and noise_std >= 0");
  TVT_REQUIRE(a->out_dtype == TVT_BF16 || a->out_dtype == TVT_F32, "tvt_feature_augment: bad out_dtype");
  TVT_REQUIRE(al16(a->x) && al16(a->y), "tvt_feature_augment: pointers must be 16-byte aligned");
  if (a->rows == 0) return TVT_OK;
  int rc = require_sm100();
  if (rc != TVT_OK) return rc;
  const long long items = a->rows * (a->d_out / 4);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (a->d_in % 8 == 0 && a->d_out % 8 == 0) {          // warp per row (see feature_augment_rows_kernel)
    const long long blocks = (a->rows + 7) / 8, cap = static_cast<long long>(num_sms()) * 8;
    const int grid = static_cast<int>(blocks < cap ? blocks : cap);
    if (a->out_dtype == TVT_F32)
      misc::feature_augment_rows_kernel<float><<<grid, 256, 0, s>>>(a->x, (float*)a->y, a->rows, (int)a->d_in, (int)a->d_out, a->p_drop, a->p_noise, a->noise_std, a->seed, seed_source());
    else
      misc::feature_augment_rows_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(a->x, (__nv_bfloat16*)a->y, a->rows, (int)a->d_in, (int)a->d_out, a->p_drop, a->p_noise, a->noise_std, a->seed, seed_source());
    return check_launch("tvt_feature_augment");
  }
  if (a->out_dtype == TVT_F32)
    misc::feature_augment_kernel<float><<<misc::grid1d(items, 256), 256, 0, s>>>(a->x, (float*)a->y, a->rows, (int)a->d_in, (int)a->d_out, a->p_drop, a->p_noise, a->noise_std, a->seed, seed_source());
  else
    misc::feature_augment_kernel<__nv_bfloat16><<<misc::grid1d(items, 256), 256, 0, s>>>(a->x, (__nv_bfloat16*)a->y, a->rows, (int)a->d_in, (int)a->d_out, a->p_drop, a->p_noise, a->noise_std, a->seed, seed_source());
  return check_launch("tvt_feature_augment");
}
