// Spatial-temporal pyramid pooling kernels (see include/tvt.h).
//   tvt_pyramid_pool_fwd / _bwd : every sum_group scale of Reasoning.forward (reference
//                                 src/models/TPN.py:64-72,106-110) in ONE pass over the frame tokens,
//                                 with the leading ReLU of each relation MLP (TPN.py:89) fused.
//   tvt_spatial_pool_fwd / _bwd : AvgPool2d to 1x1 of Feature_Pyramid_{low,Mid,High} (TPN.py:5,19,32) and its gradient.
// Pure bandwidth kernels: 16-byte vector IO, each input element read exactly once.
#include "tvt_common.cuh"
#include "tvt_ptx.cuh"

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

struct SpatialParams {
  const void* x; void* out; long long frames, C, hw, ld_out, col_off; int out_f32;
  long long rows_per_tile, ntiles;   // tiled kernel only
};

// Fallback (unaligned base, rows longer than a tile): one warp per (frame, channel), coalesced scalar reads, shuffle reduce.
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

// The bandwidth kernel of the spatial pyramid (175 616 elements read per frame, 896 written).  The
// [frames * C, HW] map is a flat stream of short rows (HW = 784 / 196 / 49: 3136 / 784 / 196 bytes in fp32, the last
// not even 16-byte aligned per row), so instead of shaping the loads around rows, persistent CTAs pull TILES of whole
// rows (<= kSpatialTileBytes, a multiple of 16 bytes by construction) into a 3-slot shared-memory ring with
// cp.async.bulk (one thread issues, mbarrier complete_tx; bytes in flight do not depend on registers or on the warps'
// instruction streams) and reduce the rows out of shared memory: a warp per row for long rows, a LANE per row when
// HW is short and odd (lane stride HW words is then conflict-free) so no lane idles on 49-element rows.
constexpr int kSpatialTileBytes = 24 * 1024, kSpatialStages = 3, kSpatialThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kSpatialThreads) spatial_tile_kernel(const SpatialParams p) {
  extern __shared__ __align__(128) unsigned char sp_smem[];
  __shared__ __align__(8) uint64_t full_bar[kSpatialStages];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long rows = p.frames * p.C, R = p.rows_per_tile;
  const int hw = static_cast<int>(p.hw);
  const float inv = 1.0f / static_cast<float>(hw);
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kSpatialStages; ++s) mbar_init(smem_u32(&full_bar[s]), 1);
    fence_mbar_init();
  }
  __syncthreads();
  const unsigned char* gx = reinterpret_cast<const unsigned char*>(p.x);
  auto issue = [&](int slot, long long tile) {
    const long long r0 = tile * R;
    const long long nr = rows - r0 < R ? rows - r0 : R;
    const uint32_t bytes = static_cast<uint32_t>(nr * hw * static_cast<long long>(sizeof(T)));   // host: every tile a multiple of 16
    const uint32_t bar = smem_u32(&full_bar[slot]);
    mbar_arrive_expect_tx(bar, bytes);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(sp_smem + slot * kSpatialTileBytes)), "l"(gx + r0 * hw * static_cast<long long>(sizeof(T))), "r"(bytes), "r"(bar)
                 : "memory");
  };
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kSpatialStages; ++s) {
      const long long t = blockIdx.x + static_cast<long long>(s) * gridDim.x;
      if (t < p.ntiles) issue(s, t);
    }
  }
  const bool lane_rows = hw < 64 && (hw & 1);
  int it = 0;
  for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
    const int slot = it % kSpatialStages;
    mbar_wait(smem_u32(&full_bar[slot]), (it / kSpatialStages) & 1);
    const T* t = reinterpret_cast<const T*>(sp_smem + slot * kSpatialTileBytes);
    const long long r0 = tile * R;
    const int nr = static_cast<int>(rows - r0 < R ? rows - r0 : R);
    const long long f0 = r0 / p.C;                      // one 64-bit division per tile; rows step from it
    const int c_first = static_cast<int>(r0 - f0 * p.C), Cc = static_cast<int>(p.C);
    auto emit = [&](int r, float acc) {
      int c = c_first + r;
      const int df = c / Cc;                              // 32-bit
      c -= df * Cc;
      const long long o = (f0 + df) * p.ld_out + p.col_off + c;
      if (p.out_f32) reinterpret_cast<float*>(p.out)[o] = acc * inv;
      else reinterpret_cast<__nv_bfloat16*>(p.out)[o] = __float2bfloat16_rn(acc * inv);
    };
    if (lane_rows) {
      for (int r = tid; r < nr; r += kSpatialThreads) {
        const T* row = t + r * hw;
        float a0 = 0.0f, a1 = 0.0f;
        int i = 0;
        for (; i + 1 < hw; i += 2) { a0 += Elem<T>::to_f(row[i]); a1 += Elem<T>::to_f(row[i + 1]); }
        if (i < hw) a0 += Elem<T>::to_f(row[i]);
        emit(r, a0 + a1);
      }
    } else if (hw < 512) {
      // medium rows (14 x 14 = 196): EIGHT lanes per row, four rows per warp pass - a warp per row would spend more
      // instructions on the shuffle tree and the output address than on the row's few loads.  Each lane sums a contiguous
      // eighth of the row (chunk starts 25 words apart: spread over the banks)
      const int sub = lane >> 3, l8 = lane & 7;
      const bool words = sizeof(T) == 2 && (hw & 1) == 0;   // bf16 rows of even length start 4-byte aligned: two elements per load
      const int n = words ? hw >> 1 : hw, seg = (n + 7) >> 3;
      for (int base = warp * 4; base < nr; base += (kSpatialThreads / 32) * 4) {
        const int r = base + sub;
        const bool valid = r < nr;
        float a0 = 0.0f, a1 = 0.0f;
        if (valid) {
          const int i0 = l8 * seg, i1 = i0 + seg < n ? i0 + seg : n;
          if (words) {
            const uint32_t* rw = reinterpret_cast<const uint32_t*>(t + r * hw);
            for (int i = i0; i < i1; ++i) {
              const uint32_t w = rw[i];
              a0 += __uint_as_float(w << 16);
              a1 += __uint_as_float(w & 0xFFFF0000u);
            }
          } else {
            const T* row = t + r * hw;
            int i = i0;
            for (; i + 1 < i1; i += 2) { a0 += Elem<T>::to_f(row[i]); a1 += Elem<T>::to_f(row[i + 1]); }
            if (i < i1) a0 += Elem<T>::to_f(row[i]);
          }
        }
        float acc = a0 + a1;
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        if (valid && l8 == 0) emit(r, acc);
      }
    } else {
      for (int r = warp; r < nr; r += kSpatialThreads / 32) {
        const T* row = t + r * hw;
        float a0 = 0.0f, a1 = 0.0f;
        if (sizeof(T) == 2 && (hw & 1) == 0) {            // bf16 rows of even length start 4-byte aligned: two elements per load
          const uint32_t* rw = reinterpret_cast<const uint32_t*>(row);
          const int nw = hw >> 1;
          for (int i = lane; i < nw; i += 32) {
            const uint32_t w = rw[i];
            a0 += __uint_as_float(w << 16);
            a1 += __uint_as_float(w & 0xFFFF0000u);
          }
        } else {
          int i = lane;
          for (; i + 32 < hw; i += 64) { a0 += Elem<T>::to_f(row[i]); a1 += Elem<T>::to_f(row[i + 32]); }
          if (i < hw) a0 += Elem<T>::to_f(row[i]);
        }
        const float acc = warp_sum(a0 + a1);
        if (lane == 0) emit(r, acc);
      }
    }
    __syncthreads();                       // every warp is done with the slot
    const long long next = tile + static_cast<long long>(kSpatialStages) * gridDim.x;
    if (tid == 0 && next < p.ntiles) {
      fence_proxy_async_smem();            // generic-proxy reads of the slot before the async-proxy refill
      issue(slot, next);
    }
  }
}

// Backward of the average pool: dx[f, c, :] = dpooled[f, col_off + c] / HW — a pure write stream.  A CTA owns a
// contiguous chunk of rows whose start is 16-byte aligned (host), stages the chunk's pooled gradients in shared memory
// and writes 16-byte vectors; the row of a vector's first element is one 32-bit division, its neighbours step from it.
struct SpatialBwdParams { const float* dpooled; void* dx; long long frames, C, hw, ld, col_off, rows_per_chunk, nchunks; };
constexpr int kSpatialBwdRows = 512;

template <typename T>
__global__ void __launch_bounds__(256) spatial_bwd_kernel(const SpatialBwdParams p) {
  constexpr int V = Vec16<T>::kN;
  __shared__ float g[kSpatialBwdRows];
  const long long rows = p.frames * p.C;
  const uint32_t hw = static_cast<uint32_t>(p.hw);
  const float inv = 1.0f / static_cast<float>(hw);
  for (long long chunk = blockIdx.x; chunk < p.nchunks; chunk += gridDim.x) {
    const long long r0 = chunk * p.rows_per_chunk;
    const int nr = static_cast<int>(rows - r0 < p.rows_per_chunk ? rows - r0 : p.rows_per_chunk);
    __syncthreads();
    for (int r = threadIdx.x; r < nr; r += blockDim.x) {
      const long long fc = r0 + r, f = fc / p.C, c = fc - f * p.C;
      g[r] = p.dpooled[f * p.ld + p.col_off + c] * inv;
    }
    __syncthreads();
    const uint32_t elems = static_cast<uint32_t>(nr) * hw;
    T* dst = reinterpret_cast<T*>(p.dx) + r0 * p.hw;
    const uint32_t nvec = elems / V;
    for (uint32_t v = threadIdx.x; v < nvec; v += blockDim.x) {
      const uint32_t e = v * V;
      uint32_t r = e / hw, rem = e - r * hw;
      float o[V];
#pragma unroll
      for (int i = 0; i < V; ++i) {
        o[i] = g[r];
        if (++rem == hw) { rem = 0; ++r; }
      }
      Vec16<T>::store(dst + e, o);
    }
    for (uint32_t e = nvec * V + threadIdx.x; e < elems; e += blockDim.x) dst[e] = Elem<T>::from_f(g[e / hw]);   // < V tail elements
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
  pool::SpatialParams p{a->x, a->out, a->frames, a->channels, a->hw, a->ld_out, a->col_offset, a->out_dtype == TVT_F32, 0, 0};
  const long long rows = a->frames * a->channels;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // tile = the largest multiple of r0 rows within kSpatialTileBytes, r0 = rows per 16-byte period of the row pitch
  const long long row_bytes = a->hw * (a->dtype == TVT_F32 ? 4 : 2);
  long long gcd = row_bytes, y = 16;
  while (y) { const long long t = gcd % y; gcd = y; y = t; }
  const long long r0 = 16 / gcd;
  long long R = (pool::kSpatialTileBytes / (r0 * row_bytes)) * r0;
  // the final (partial) tile must also be a whole number of 16-byte units: true when rows % r0 == 0
  if (R > 0 && al16(a->x) && rows % r0 == 0 && a->hw < (1 << 20)) {
    if (R >= 64) R -= R % 32;                                    // whole passes of the 8 warps (4 rows each for medium rows)
    else if (R >= 16) R -= R % 8;
    p.rows_per_tile = R;
    p.ntiles = (rows + R - 1) / R;
    static bool attr_set = false;
    const int smem = pool::kSpatialStages * pool::kSpatialTileBytes;
    if (!attr_set) {
      cudaFuncSetAttribute(pool::spatial_tile_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      cudaFuncSetAttribute(pool::spatial_tile_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      attr_set = true;
    }
    const long long cap = static_cast<long long>(num_sms()) * 3;   // 72 KB per CTA: three resident per SM
    const int grid = static_cast<int>(p.ntiles < cap ? p.ntiles : cap);
    if (a->dtype == TVT_F32) pool::spatial_tile_kernel<float><<<grid, pool::kSpatialThreads, smem, s>>>(p);
    else pool::spatial_tile_kernel<__nv_bfloat16><<<grid, pool::kSpatialThreads, smem, s>>>(p);
    return check_launch("tvt_spatial_pool_fwd");
  }
  if (a->dtype == TVT_F32) pool::spatial_kernel<float><<<pool::grid_for(rows * 32, 256), 256, 0, s>>>(p);
  else pool::spatial_kernel<__nv_bfloat16><<<pool::grid_for(rows * 32, 256), 256, 0, s>>>(p);
  return check_launch("tvt_spatial_pool_fwd");
}

extern "C" int tvt_spatial_pool_bwd(const tvt_spatial_pool_bwd_args* a, void* stream) {
  using namespace tvt;
  TVT_REQUIRE(a != nullptr && a->dpooled && a->dx, "tvt_spatial_pool_bwd: null pointer");
  TVT_REQUIRE(a->frames > 0 && a->channels > 0 && a->hw > 0 && a->hw < (1 << 20), "tvt_spatial_pool_bwd: bad shape");
  TVT_REQUIRE(a->ld >= a->col_offset + a->channels, "tvt_spatial_pool_bwd: ld too small");
  TVT_REQUIRE(a->dtype == TVT_BF16 || a->dtype == TVT_F32, "tvt_spatial_pool_bwd: bad dtype");
  TVT_REQUIRE(al16(a->dx), "tvt_spatial_pool_bwd: dx must be 16-byte aligned");
  int rc = require_sm100();
  if (rc != TVT_OK) return rc;
  const long long rows = a->frames * a->channels;
  const long long row_bytes = a->hw * (a->dtype == TVT_F32 ? 4 : 2);
  long long gcd = row_bytes, y = 16;
  while (y) { const long long t = gcd % y; gcd = y; y = t; }
  const long long r0 = 16 / gcd;                                  // chunk starts stay 16-byte aligned
  long long R = (pool::kSpatialBwdRows / r0) * r0;
  while (R > r0 && R * a->hw > (1LL << 16)) R -= r0;              // ~64 Ki elements per chunk
  pool::SpatialBwdParams p{static_cast<const float*>(a->dpooled), a->dx, a->frames, a->channels, a->hw, a->ld, a->col_offset, R, (rows + R - 1) / R};
  const long long cap = static_cast<long long>(num_sms()) * 8;
  const int grid = static_cast<int>(p.nchunks < cap ? p.nchunks : cap);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (a->dtype == TVT_F32) pool::spatial_bwd_kernel<float><<<grid, 256, 0, s>>>(p);
  else pool::spatial_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(p);
  return check_launch("tvt_spatial_pool_bwd");
}
