// Collaborative gating fusion glue (reference: src/models/collabgating.py:17-56,59-87).
// The reference's O(E^2) per-vector Linear calls collapse into three stacked tensor-core GEMMs (hostapi/collabgating.py);
// what sits BETWEEN those GEMMs - the nearest-neighbour stretch of narrow experts to 2048, the pairwise-sum bookkeeping
// T_i = (E-1) C_i + sum_{j>i} C_j + sum_{j<i} P(C_j), the GLU gate sum_i C_i * sigmoid(C_i + A_i) and the L2 normalisation of the
// embedding - is elementwise / row-wise bandwidth work: one kernel each, forward and backward, 16-byte vector IO, every
// element read once.
#include "tvt_common.cuh"

namespace tvt {
namespace collab {

constexpr int kMaxE = TVT_MAX_EXPERTS;

template <typename T>
__global__ void __launch_bounds__(256) stretch_kernel(const float* x, T* y, long long rows, int d_in, int d_out, long long ld_out) {
  const long long total = rows * (d_out / 4);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / (d_out / 4);
    const int c = static_cast<int>(i - r * (d_out / 4)) * 4;
    const float* src = x + r * d_in;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = src[static_cast<long long>(c + j) * d_in / d_out];   // F.interpolate(mode="nearest"): floor(i * in / out)
    T* dst = y + r * ld_out + c;
    if constexpr (sizeof(T) == 4) *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    else *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
  }
}

struct MixParams { const void* c; const void* pc; void* t; const void* dt; void* dc; void* dpc; long long n, d; int E; };

// T_i = (E - 1) C_i + sum_{j > i} C_j + sum_{j < i} PC_j   (thread = one 16-byte column vector of one row, all experts).
// Two sweeps over the experts (total, then running sums) keep the register footprint independent of E; the second sweep's
// loads hit L1 / L2.
template <typename T>
__global__ void __launch_bounds__(256) mix_fwd_kernel(const MixParams p) {
  constexpr int V = Vec16<T>::kN;
  const long long vecs = p.n * (p.d / V), plane = p.n * p.d;
  const float w = static_cast<float>(p.E - 1);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < vecs; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long off = i * V;
    float total[V], cum[V], pcrun[V];
#pragma unroll
    for (int k = 0; k < V; ++k) total[k] = cum[k] = pcrun[k] = 0.0f;
    for (int e = 0; e < p.E; ++e) {
      float c[V];
      Vec16<T>::load(reinterpret_cast<const T*>(p.c) + e * plane + off, c);
#pragma unroll
      for (int k = 0; k < V; ++k) total[k] += c[k];
    }
    for (int e = 0; e < p.E; ++e) {
      float c[V], o[V];
      Vec16<T>::load(reinterpret_cast<const T*>(p.c) + e * plane + off, c);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        cum[k] += c[k];
        o[k] = w * c[k] + (total[k] - cum[k]) + pcrun[k];
      }
      Vec16<T>::store(reinterpret_cast<T*>(p.t) + e * plane + off, o);
      if (e + 1 < p.E) {
        float pc[V];
        Vec16<T>::load(reinterpret_cast<const T*>(p.pc) + e * plane + off, pc);
#pragma unroll
        for (int k = 0; k < V; ++k) pcrun[k] += pc[k];
      }
    }
  }
}

// dC_j = (E - 1) dT_j + sum_{i < j} dT_i ;  dPC_j = sum_{i > j} dT_i
template <typename T>
__global__ void __launch_bounds__(256) mix_bwd_kernel(const MixParams p) {
  constexpr int V = Vec16<T>::kN;
  const long long vecs = p.n * (p.d / V), plane = p.n * p.d;
  const float w = static_cast<float>(p.E - 1);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < vecs; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long off = i * V;
    float total[V], cum[V];
#pragma unroll
    for (int k = 0; k < V; ++k) total[k] = cum[k] = 0.0f;
    for (int e = 0; e < p.E; ++e) {
      float g[V];
      Vec16<T>::load(reinterpret_cast<const T*>(p.dt) + e * plane + off, g);
#pragma unroll
      for (int k = 0; k < V; ++k) total[k] += g[k];
    }
    for (int e = 0; e < p.E; ++e) {
      float g[V], o[V];
      Vec16<T>::load(reinterpret_cast<const T*>(p.dt) + e * plane + off, g);
#pragma unroll
      for (int k = 0; k < V; ++k) { o[k] = w * g[k] + cum[k]; cum[k] += g[k]; }
      Vec16<T>::store(reinterpret_cast<T*>(p.dc) + e * plane + off, o);
      if (e + 1 < p.E) {
#pragma unroll
        for (int k = 0; k < V; ++k) o[k] = total[k] - cum[k];
        Vec16<T>::store(reinterpret_cast<T*>(p.dpc) + e * plane + off, o);
      }
    }
  }
}

struct GateParams { const void* c; const void* a; void* g; const void* dg; void* dc; void* da; long long n, d; int E; };

// g = sum_i C_i * sigmoid(C_i + A_i)   (ContextGating's GLU on cat(C_i, C_i + A_i), collabgating.py:83-85, summed over experts :50)
template <typename T>
__global__ void __launch_bounds__(256) gate_fwd_kernel(const GateParams p) {
  constexpr int V = Vec16<T>::kN;
  const long long vecs = p.n * (p.d / V), plane = p.n * p.d;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < vecs; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long off = i * V;
    float acc[V];
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] = 0.0f;
    for (int e = 0; e < p.E; ++e) {
      float c[V], a[V];
      Vec16<T>::load(reinterpret_cast<const T*>(p.c) + e * plane + off, c);
      Vec16<T>::load(reinterpret_cast<const T*>(p.a) + e * plane + off, a);
#pragma unroll
      for (int k = 0; k < V; ++k) acc[k] += c[k] / (1.0f + expf(-(c[k] + a[k])));
    }
    Vec16<T>::store(reinterpret_cast<T*>(p.g) + off, acc);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) gate_bwd_kernel(const GateParams p) {
  constexpr int V = Vec16<T>::kN;
  const long long vecs = p.n * (p.d / V), plane = p.n * p.d;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < vecs; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long off = i * V;
    float dg[V];
    Vec16<T>::load(reinterpret_cast<const T*>(p.dg) + off, dg);
    for (int e = 0; e < p.E; ++e) {
      float c[V], a[V], dc[V], da[V];
      Vec16<T>::load(reinterpret_cast<const T*>(p.c) + e * plane + off, c);
      Vec16<T>::load(reinterpret_cast<const T*>(p.a) + e * plane + off, a);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const float s = 1.0f / (1.0f + expf(-(c[k] + a[k])));
        const float ds = c[k] * s * (1.0f - s);
        dc[k] = dg[k] * (s + ds);
        da[k] = dg[k] * ds;
      }
      Vec16<T>::store(reinterpret_cast<T*>(p.dc) + e * plane + off, dc);
      Vec16<T>::store(reinterpret_cast<T*>(p.da) + e * plane + off, da);
    }
  }
}

// y = x / max(||x||_2, eps) per row (F.normalize, collabgating.py:66-69); one warp per row; y fp32.
template <typename T>
__global__ void __launch_bounds__(256) l2norm_fwd_kernel(const T* x, float* y, float* inv_norm, long long rows, int d, float eps) {
  constexpr int V = Vec16<T>::kN;
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5, nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long r = warp0; r < rows; r += nwarps) {
    float ss = 0.0f;
    for (int c = lane * V; c < d; c += 32 * V) {
      float v[V];
      Vec16<T>::load(x + r * d + c, v);
#pragma unroll
      for (int k = 0; k < V; ++k) ss = fmaf(v[k], v[k], ss);
    }
    ss = warp_sum(ss);
    const float inv = 1.0f / fmaxf(sqrtf(ss), eps);
    if (lane == 0 && inv_norm) inv_norm[r] = inv;
    for (int c = lane * V; c < d; c += 32 * V) {
      float v[V];
      Vec16<T>::load(x + r * d + c, v);
#pragma unroll
      for (int k = 0; k < V; k += 4) *reinterpret_cast<float4*>(y + r * d + c + k) = make_float4(v[k] * inv, v[k + 1] * inv, v[k + 2] * inv, v[k + 3] * inv);
    }
  }
}

// dx = (dy - y (y . dy)) * inv_norm   (rows whose norm was clamped by eps: dx = dy / eps, torch's clamp_min gradient)
template <typename T>
__global__ void __launch_bounds__(256) l2norm_bwd_kernel(const float* dy, const float* y, const float* inv_norm, T* dx, long long rows, int d, float eps) {
  constexpr int V = Vec16<T>::kN;
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5, nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long r = warp0; r < rows; r += nwarps) {
    float dot = 0.0f;
    for (int c = lane * 4; c < d; c += 128) {
      const float4 a = *reinterpret_cast<const float4*>(dy + r * d + c), b = *reinterpret_cast<const float4*>(y + r * d + c);
      dot += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
    }
    dot = warp_sum(dot);
    const float inv = inv_norm[r];
    if (inv >= 1.0f / eps) dot = 0.0f;            // the clamp was active: the norm does not depend on x
    for (int c = lane * V; c < d; c += 32 * V) {
      float o[V];
#pragma unroll
      for (int k = 0; k < V; ++k) o[k] = (dy[r * d + c + k] - y[r * d + c + k] * dot) * inv;
      Vec16<T>::store(dx + r * d + c, o);
    }
  }
}

static int grid_for(long long items) {
  const long long want = (items + 255) / 256, cap = static_cast<long long>(num_sms()) * 8;
  return static_cast<int>(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace collab
}  // namespace tvt

static bool al16c(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int tvt_stretch_cast(const tvt_stretch_cast_args* a, void* stream) {
  using namespace tvt;
  TVT_REQUIRE(a != nullptr && a->x && a->y, "tvt_stretch_cast: null pointer");
  TVT_REQUIRE(a->rows >= 0 && a->d_in > 0 && a->d_out > 0 && a->d_out % 4 == 0 && a->ld_out >= a->d_out && a->ld_out % 8 == 0,
              "tvt_stretch_cast: need d_out % 4 == 0 and ld_out >= d_out, a multiple of 8");
  TVT_REQUIRE(a->out_dtype == TVT_BF16 || a->out_dtype == TVT_F32, "tvt_stretch_cast: bad out_dtype");
  TVT_REQUIRE(al16c(a->y), "tvt_stretch_cast: y must be 16-byte aligned");
  if (a->rows == 0) return TVT_OK;
  int rc = require_sm100();
  if (rc != TVT_OK) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = collab::grid_for(a->rows * (a->d_out / 4));
  if (a->out_dtype == TVT_F32) collab::stretch_kernel<float><<<grid, 256, 0, s>>>(a->x, (float*)a->y, a->rows, (int)a->d_in, (int)a->d_out, a->ld_out);
  else collab::stretch_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(a->x, (__nv_bfloat16*)a->y, a->rows, (int)a->d_in, (int)a->d_out, a->ld_out);
  return check_launch("tvt_stretch_cast");
}

static int collab_check(const tvt_collab_args* a, const char* who) {
  using namespace tvt;
  TVT_REQUIRE(a != nullptr, "%s: null args", who);
  TVT_REQUIRE(a->experts >= 2 && a->experts <= TVT_MAX_EXPERTS, "%s: experts must be in [2, %d]", who, TVT_MAX_EXPERTS);
  TVT_REQUIRE(a->rows > 0 && a->d > 0 && a->d % 8 == 0, "%s: bad shape (d must be a multiple of 8)", who);
  TVT_REQUIRE(a->dtype == TVT_BF16 || a->dtype == TVT_F32, "%s: bad dtype", who);
  return TVT_OK;
}

extern "C" int tvt_collab_mix_fwd(const tvt_collab_args* a, void* stream) {
  using namespace tvt;
  int rc = collab_check(a, "tvt_collab_mix_fwd");
  if (rc != TVT_OK) return rc;
  TVT_REQUIRE(a->c && a->pc && a->out && al16c(a->c) && al16c(a->pc) && al16c(a->out), "tvt_collab_mix_fwd: null / unaligned pointer");
  if ((rc = require_sm100()) != TVT_OK) return rc;
  collab::MixParams p{a->c, a->pc, a->out, nullptr, nullptr, nullptr, a->rows, a->d, a->experts};
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = collab::grid_for(a->rows * (a->d / (a->dtype == TVT_F32 ? 4 : 8)));
  if (a->dtype == TVT_F32) collab::mix_fwd_kernel<float><<<grid, 256, 0, s>>>(p);
  else collab::mix_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(p);
  return check_launch("tvt_collab_mix_fwd");
}

extern "C" int tvt_collab_mix_bwd(const tvt_collab_args* a, void* stream) {
  using namespace tvt;
  int rc = collab_check(a, "tvt_collab_mix_bwd");
  if (rc != TVT_OK) return rc;
  TVT_REQUIRE(a->dout && a->dc && a->dpc && al16c(a->dout) && al16c(a->dc) && al16c(a->dpc), "tvt_collab_mix_bwd: null / unaligned pointer");
  if ((rc = require_sm100()) != TVT_OK) return rc;
  collab::MixParams p{nullptr, nullptr, nullptr, a->dout, a->dc, a->dpc, a->rows, a->d, a->experts};
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = collab::grid_for(a->rows * (a->d / (a->dtype == TVT_F32 ? 4 : 8)));
  if (a->dtype == TVT_F32) collab::mix_bwd_kernel<float><<<grid, 256, 0, s>>>(p);
  else collab::mix_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(p);
  return check_launch("tvt_collab_mix_bwd");
}

extern "C" int tvt_collab_gate_fwd(const tvt_collab_args* a, void* stream) {
  using namespace tvt;
  int rc = collab_check(a, "tvt_collab_gate_fwd");
  if (rc != TVT_OK) return rc;
  TVT_REQUIRE(a->c && a->a && a->out && al16c(a->c) && al16c(a->a) && al16c(a->out), "tvt_collab_gate_fwd: null / unaligned pointer");
  if ((rc = require_sm100()) != TVT_OK) return rc;
  collab::GateParams p{a->c, a->a, a->out, nullptr, nullptr, nullptr, a->rows, a->d, a->experts};
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = collab::grid_for(a->rows * (a->d / (a->dtype == TVT_F32 ? 4 : 8)));
  if (a->dtype == TVT_F32) collab::gate_fwd_kernel<float><<<grid, 256, 0, s>>>(p);
  else collab::gate_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(p);
  return check_launch("tvt_collab_gate_fwd");
}

extern "C" int tvt_collab_gate_bwd(const tvt_collab_args* a, void* stream) {
  using namespace tvt;
  int rc = collab_check(a, "tvt_collab_gate_bwd");
  if (rc != TVT_OK) return rc;
  TVT_REQUIRE(a->c && a->a && a->dout && a->dc && a->da && al16c(a->c) && al16c(a->a) && al16c(a->dout) && al16c(a->dc) && al16c(a->da),
              "tvt_collab_gate_bwd: null / unaligned pointer");
  if ((rc = require_sm100()) != TVT_OK) return rc;
  collab::GateParams p{a->c, a->a, nullptr, a->dout, a->dc, a->da, a->rows, a->d, a->experts};
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = collab::grid_for(a->rows * (a->d / (a->dtype == TVT_F32 ? 4 : 8)));
  if (a->dtype == TVT_F32) collab::gate_bwd_kernel<float><<<grid, 256, 0, s>>>(p);
  else collab::gate_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(p);
  return check_launch("tvt_collab_gate_bwd");
}

extern "C" int tvt_l2norm_fwd(const tvt_l2norm_args* a, void* stream) {
  using namespace tvt;
  TVT_REQUIRE(a != nullptr && a->x && a->y && a->inv_norm, "tvt_l2norm_fwd: null pointer");
  TVT_REQUIRE(a->rows >= 0 && a->d > 0 && a->d % 8 == 0, "tvt_l2norm_fwd: d must be a multiple of 8");
  TVT_REQUIRE(a->dtype == TVT_BF16 || a->dtype == TVT_F32, "tvt_l2norm_fwd: bad dtype");
  TVT_REQUIRE(al16c(a->x) && al16c(a->y), "tvt_l2norm_fwd: pointers must be 16-byte aligned");
  if (a->rows == 0) return TVT_OK;
  int rc = require_sm100();
  if (rc != TVT_OK) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = collab::grid_for(a->rows * 32);
  if (a->dtype == TVT_F32) collab::l2norm_fwd_kernel<float><<<grid, 256, 0, s>>>((const float*)a->x, a->y, a->inv_norm, a->rows, (int)a->d, a->eps);
  else collab::l2norm_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)a->x, a->y, a->inv_norm, a->rows, (int)a->d, a->eps);
  return check_launch("tvt_l2norm_fwd");
}

extern "C" int tvt_l2norm_bwd(const tvt_l2norm_args* a, void* stream) {
  using namespace tvt;
  TVT_REQUIRE(a != nullptr && a->dy && a->y && a->inv_norm && a->dx, "tvt_l2norm_bwd: null pointer");
  TVT_REQUIRE(a->rows >= 0 && a->d > 0 && a->d % 8 == 0, "tvt_l2norm_bwd: d must be a multiple of 8");
  TVT_REQUIRE(a->dtype == TVT_BF16 || a->dtype == TVT_F32, "tvt_l2norm_bwd: bad dtype");
  TVT_REQUIRE(al16c(a->dy) && al16c(a->y) && al16c(a->dx), "tvt_l2norm_bwd: pointers must be 16-byte aligned");
  if (a->rows == 0) return TVT_OK;
  int rc = require_sm100();
  if (rc != TVT_OK) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = collab::grid_for(a->rows * 32);
  if (a->dtype == TVT_F32) collab::l2norm_bwd_kernel<float><<<grid, 256, 0, s>>>(a->dy, a->y, a->inv_norm, (float*)a->dx, a->rows, (int)a->d, a->eps);
  else collab::l2norm_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(a->dy, a->y, a->inv_norm, (__nv_bfloat16*)a->dx, a->rows, (int)a->d, a->eps);
  return check_launch("tvt_l2norm_bwd");
}
