// Spatial-temporal pyramid pooling kernels (see include/tvt.h).
//   tvt_pyramid_pool_fwd / _bwd : every sum_group scale of Reasoning.forward (reference
//                                 src/models/TPN.py:64-72,106-110) in ONE pass over the frame tokens,
//                                 with the leading ReLU of each relation MLP (TPN.py:89) fused.
//   tvt_spatial_pool_fwd        : AvgPool2d to 1x1 of Feature_Pyramid_{low,Mid,High} (TPN.py:5,19,32).
// Pure bandwidth kernels: 16-byte vector IO, each input element read exactly once.
#include "tvt_common.cuh"

namespace tvt {
namespace pool {

struct FwdParams {
  const void* x; long long B, T, d, xbs, xfs; int ns; int g[TVT_MAX_POOL_SCALES]; void* out[TVT_MAX_POOL_SCALES]; int relu;
  long long L, nseg;  // frames per thread segment (a common multiple of every group size) and segments per clip
};

// Thread -> (clip b, segment of L frames, 16-byte column vector).  L is a common multiple of every
// group size, so each group lies inside one segment and every scale's partial sums stay in registers;
// each input element is read exactly once.
template <typename T>
__global__ void __launch_bounds__(256) fwd_kernel(const FwdParams p) {
  constexpr int V = Vec16<T>::kN;
  const long long vecs = p.d / V;
  const long long total = p.B * p.nseg * vecs;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long bs = idx / vecs;
    const int col = static_cast<int>(idx - bs * vecs) * V;
    const long long b = bs / p.nseg, seg = bs - b * p.nseg;
    const long long t0 = seg * p.L, t1 = t0 + p.L < p.T ? t0 + p.L : p.T;
    const T* xb = reinterpret_cast<const T*>(p.x) + b * p.xbs + col;
    float acc[TVT_MAX_POOL_SCALES][V];
#pragma unroll
    for (int s = 0; s < TVT_MAX_POOL_SCALES; ++s)
#pragma unroll
      for (int i = 0; i < V; ++i) acc[s][i] = 0.0f;
    for (long long t = t0; t < t1; ++t) {
      float v[V];
      Vec16<T>::load(xb + t * p.xfs, v);
#pragma unroll
      for (int s = 0; s < TVT_MAX_POOL_SCALES; ++s) {
        if (s < p.ns) {
          const int g = p.g[s];
#pragma unroll
          for (int i = 0; i < V; ++i) acc[s][i] += v[i];
          if ((t + 1) % g == 0) {
            const long long k = t / g;
            float o[V];
#pragma unroll
            for (int i = 0; i < V; ++i) { o[i] = p.relu ? fmaxf(acc[s][i], 0.0f) : acc[s][i]; acc[s][i] = 0.0f; }
            Vec16<T>::store(reinterpret_cast<T*>(p.out[s]) + (b * (p.T / g) + k) * p.d + col, o);
          }
        }
      }
    }
  }
}

struct BwdParams {
  const void* dout[TVT_MAX_POOL_SCALES]; const void* out[TVT_MAX_POOL_SCALES]; void* dx;
  long long B, T, d, dbs, dfs; int ns; int g[TVT_MAX_POOL_SCALES]; int relu; int accumulate;
};

template <typename T>
__global__ void __launch_bounds__(256) bwd_kernel(const BwdParams p) {
  constexpr int V = Vec16<T>::kN;
  const long long vecs = p.d / V;
  const long long total = p.B * p.T * vecs;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long bt = idx / vecs;
    const int col = static_cast<int>(idx - bt * vecs) * V;
    const long long b = bt / p.T, t = bt - b * p.T;
    T* dst = reinterpret_cast<T*>(p.dx) + b * p.dbs + t * p.dfs + col;
    float acc[V];
    if (p.accumulate) {
      Vec16<T>::load(dst, acc);
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] = 0.0f;
    }
    for (int s = 0; s < p.ns; ++s) {
      const int g = p.g[s];
      const long long kmax = p.T / g;
      const long long k = t / g;
      if (k < kmax) {
        const long long off = (b * kmax + k) * p.d + col;
        float dv[V];
        Vec16<T>::load(reinterpret_cast<const T*>(p.dout[s]) + off, dv);
        if (p.relu) {
          float ov[V];
          Vec16<T>::load(reinterpret_cast<const T*>(p.out[s]) + off, ov);
#pragma unroll
          for (int i = 0; i < V; ++i) acc[i] += ov[i] > 0.0f ? dv[i] : 0.0f;
        } else {
#pragma unroll
          for (int i = 0; i < V; ++i) acc[i] += dv[i];
        }
      }
    }
    Vec16<T>::store(dst, acc);
  }
}

struct SpatialParams { const void* x; void* out; long long frames, C, hw, ld_out, col_off; int out_f32; };

// One warp per (frame, channel): coalesced read of the HW contiguous elements, shuffle reduce.
template <typename T>
__global__ void __launch_bounds__(256) spatial_kernel(const SpatialParams p) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const long long total = p.frames * p.C;
  for (long long fc = warp0; fc < total; fc += nwarps) {
    const T* src = reinterpret_cast<const T*>(p.x) + fc * p.hw;
    float acc = 0.0f;
    for (long long i = lane; i < p.hw; i += 32) acc += Elem<T>::to_f(src[i]);
    acc = warp_sum(acc) / static_cast<float>(p.hw);
    if (lane == 0) {
      const long long f = fc / p.C, c = fc - f * p.C;
      const long long o = f * p.ld_out + p.col_off + c;
      if (p.out_f32) reinterpret_cast<float*>(p.out)[o] = acc;
      else reinterpret_cast<__nv_bfloat16*>(p.out)[o] = __float2bfloat16_rn(acc);
    }
  }
}

static int grid_for(long long work_items, int threads) {
  const long long want = (work_items + threads - 1) / threads;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  return static_cast<int>(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace pool
}  // namespace tvt

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int tvt_pyramid_pool_fwd(const tvt_pyramid_pool_fwd_args* a, void* stream) {
  using namespace tvt;
  TVT_REQUIRE(a != nullptr && a->x, "tvt_pyramid_pool_fwd: null pointer");
  TVT_REQUIRE(a->batch > 0 && a->frames > 0 && a->d > 0 && a->d % 8 == 0, "tvt_pyramid_pool_fwd: bad shape (d must be a multiple of 8)");
  TVT_REQUIRE(a->num_scales >= 1 && a->num_scales <= TVT_MAX_POOL_SCALES, "tvt_pyramid_pool_fwd: num_scales out of range");
  TVT_REQUIRE(a->dtype == TVT_BF16 || a->dtype == TVT_F32, "tvt_pyramid_pool_fwd: bad dtype");
  TVT_REQUIRE(a->x_batch_stride % 8 == 0 && a->x_frame_stride % 8 == 0 && al16(a->x), "tvt_pyramid_pool_fwd: strides / base must keep 16-byte alignment");
  pool::FwdParams p{};
  for (int i = 0; i < a->num_scales; ++i) {
    TVT_REQUIRE(a->groups[i] >= 1 && a->groups[i] <= a->frames, "tvt_pyramid_pool_fwd: group size %d out of range", a->groups[i]);
    TVT_REQUIRE(a->out[i] && al16(a->out[i]), "tvt_pyramid_pool_fwd: bad output pointer");
    p.g[i] = a->groups[i]; p.out[i] = a->out[i];
  }
  int rc = require_sm100();
  if (rc != TVT_OK) return rc;
  p.x = a->x; p.B = a->batch; p.T = a->frames; p.d = a->d; p.xbs = a->x_batch_stride; p.xfs = a->x_frame_stride; p.ns = a->num_scales; p.relu = a->relu;
  // segment length: least common multiple of the group sizes (whole clip if that gets long)
  long long L = 1;
  for (int i = 0; i < a->num_scales; ++i) {
    long long x = L, y = a->groups[i];
    while (y) { const long long t = x % y; x = y; y = t; }
    L = L / x * a->groups[i];
    if (L > 64) { L = a->frames; break; }
  }
  if (L > a->frames) L = a->frames;
  p.L = L; p.nseg = (a->frames + L - 1) / L;
  const long long items = a->batch * p.nseg * (a->d / (a->dtype == TVT_F32 ? 4 : 8));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (a->dtype == TVT_F32) pool::fwd_kernel<float><<<pool::grid_for(items, 128), 128, 0, s>>>(p);
  else pool::fwd_kernel<__nv_bfloat16><<<pool::grid_for(items, 128), 128, 0, s>>>(p);
  return check_launch("tvt_pyramid_pool_fwd");
}

extern "C" int tvt_pyramid_pool_bwd(const tvt_pyramid_pool_bwd_args* a, void* stream) {
  using namespace tvt;
  TVT_REQUIRE(a != nullptr && a->dx, "tvt_pyramid_pool_bwd: null pointer");
  TVT_REQUIRE(a->batch > 0 && a->frames > 0 && a->d > 0 && a->d % 8 == 0, "tvt_pyramid_pool_bwd: bad shape");
  TVT_REQUIRE(a->num_scales >= 1 && a->num_scales <= TVT_MAX_POOL_SCALES, "tvt_pyramid_pool_bwd: num_scales out of range");
  TVT_REQUIRE(a->dtype == TVT_BF16 || a->dtype == TVT_F32, "tvt_pyramid_pool_bwd: bad dtype");
  TVT_REQUIRE(a->dx_batch_stride % 8 == 0 && a->dx_frame_stride % 8 == 0 && al16(a->dx), "tvt_pyramid_pool_bwd: strides / base must keep 16-byte alignment");
  pool::BwdParams p{};
  for (int i = 0; i < a->num_scales; ++i) {
    TVT_REQUIRE(a->groups[i] >= 1 && a->groups[i] <= a->frames, "tvt_pyramid_pool_bwd: group size out of range");
    TVT_REQUIRE(a->dout[i] && al16(a->dout[i]) && (!a->relu || (a->out[i] && al16(a->out[i]))), "tvt_pyramid_pool_bwd: bad dout/out pointer");
    p.g[i] = a->groups[i]; p.dout[i] = a->dout[i]; p.out[i] = a->out[i];
  }
  int rc = require_sm100();
  if (rc != TVT_OK) return rc;
  p.dx = a->dx; p.B = a->batch; p.T = a->frames; p.d = a->d; p.dbs = a->dx_batch_stride; p.dfs = a->dx_frame_stride;
  p.ns = a->num_scales; p.relu = a->relu; p.accumulate = a->accumulate;
  const long long items = a->batch * a->frames * (a->d / (a->dtype == TVT_F32 ? 4 : 8));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (a->dtype == TVT_F32) pool::bwd_kernel<float><<<pool::grid_for(items, 256), 256, 0, s>>>(p);
  else pool::bwd_kernel<__nv_bfloat16><<<pool::grid_for(items, 256), 256, 0, s>>>(p);
  return check_launch("tvt_pyramid_pool_bwd");
}

extern "C" int tvt_spatial_pool_fwd(const tvt_spatial_pool_args* a, void* stream) {
  using namespace tvt;
  TVT_REQUIRE(a != nullptr && a->x && a->out, "tvt_spatial_pool_fwd: null pointer");
  TVT_REQUIRE(a->frames > 0 && a->channels > 0 && a->hw > 0, "tvt_spatial_pool_fwd: bad shape");
  TVT_REQUIRE(a->ld_out >= a->col_offset + a->channels, "tvt_spatial_pool_fwd: ld_out too small");
  TVT_REQUIRE((a->dtype == TVT_BF16 || a->dtype == TVT_F32) && (a->out_dtype == TVT_BF16 || a->out_dtype == TVT_F32), "tvt_spatial_pool_fwd: bad dtype");
  int rc = require_sm100();
  if (rc != TVT_OK) return rc;
  pool::SpatialParams p{a->x, a->out, a->frames, a->channels, a->hw, a->ld_out, a->col_offset, a->out_dtype == TVT_F32};
  const long long warps = a->frames * a->channels;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (a->dtype == TVT_F32) pool::spatial_kernel<float><<<pool::grid_for(warps * 32, 256), 256, 0, s>>>(p);
  else pool::spatial_kernel<__nv_bfloat16><<<pool::grid_for(warps * 32, 256), 256, 0, s>>>(p);
  return check_launch("tvt_spatial_pool_fwd");
}
