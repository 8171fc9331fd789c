// Shared host/device helpers for the tvt kernels (status handling, vector IO, reductions, RNG).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/tvt.h"

namespace tvt {

// ---------------------------------------------------------------- host side
void set_last_error(const char* fmt, ...);
int num_sms();
// Returns TVT_OK or records the CUDA error string and returns TVT_ECUDA.
int check_launch(const char* what);
// TVT_EARCH unless the current device is compute capability 10.x.
int require_sm100();
// Device pointer registered with tvt_set_seed_source (nullptr when none): see mix_seed below.
const unsigned long long* seed_source();

#define TVT_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      ::tvt::set_last_error(__VA_ARGS__); \
      return TVT_EINVAL;                \
    }                                   \
  } while (0)

// ---------------------------------------------------------------- device side
#ifdef __CUDACC__

// Launch with programmatic stream serialisation (see pdl_wait / pdl_trigger in tvt_ptx.cuh); optional cluster width.
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream, int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  static const int allowed = getenv("TVT_NO_PDL") ? 0 : 1;
  attr[n].val.programmaticStreamSerializationAllowed = allowed;
  ++n;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

template <typename T>
struct Elem;
template <>
struct Elem<float> {
  static __device__ __forceinline__ float to_f(float v) { return v; }
  static __device__ __forceinline__ float from_f(float v) { return v; }
};
template <>
struct Elem<__nv_bfloat16> {
  static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};

// 16-byte vector load of kVec elements (8 bf16 / 4 fp32) converted to float.
template <typename T>
struct Vec16;
template <>
struct Vec16<float> {
  static constexpr int kN = 4;
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
  static __device__ __forceinline__ void unpack(const uint4& t, float (&v)[4]) {
    v[0] = __uint_as_float(t.x); v[1] = __uint_as_float(t.y); v[2] = __uint_as_float(t.z); v[3] = __uint_as_float(t.w);
  }
};
template <>
struct Vec16<__nv_bfloat16> {
  static constexpr int kN = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = t;
  }
  static __device__ __forceinline__ void unpack(const uint4& t, float (&v)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
};

// 16-byte streaming load that bypasses L1 allocation (each element of these kernels is read exactly once).
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Exact (erf) GELU and its derivative, matching nn.GELU() default.
__device__ __forceinline__ float gelu_f(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
}
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// Counter-based dropout mask: keep(element) is a pure function of (seed, element index), so the
// backward pass regenerates the forward mask instead of storing it.  One Philox-2x32 evaluation (5 rounds:
// the fewest whose keep-bits show no measurable lag / cross-seed correlation on 2^21-element streams, see
// DESIGN.md) yields 64 bits = 4 independent 16-bit lanes for 4 consecutive elements; ~15 integer
// instructions per 4 elements.
constexpr int kDropoutRounds = 5;
constexpr uint32_t kDropoutMul = 0xD256D193u, kDropoutWeyl = 0x9E3779B9u;
// Same hash with the per-round keys (seed_lo + r * kDropoutWeyl) and the starting words handed in, for callers
// that hoist them out of a hot loop; L0 = high word of elem_div4 ^ high word of the seed.
__device__ __forceinline__ void dropout_words(const uint32_t (&rk)[kDropoutRounds], uint32_t R, uint32_t L, uint32_t& w_lo, uint32_t& w_hi) {
#pragma unroll
  for (int r = 0; r < kDropoutRounds; ++r) {
    const uint64_t m = static_cast<uint64_t>(R) * kDropoutMul;
    R = static_cast<uint32_t>(m >> 32) ^ rk[r] ^ L;
    L = static_cast<uint32_t>(m);
  }
  w_lo = R;   // lanes 0, 1
  w_hi = L;   // lanes 2, 3
}
__device__ __forceinline__ uint64_t dropout_bits4(uint64_t seed, uint64_t elem_div4) {
  uint32_t rk[kDropoutRounds];
#pragma unroll
  for (int r = 0; r < kDropoutRounds; ++r) rk[r] = static_cast<uint32_t>(seed) + r * kDropoutWeyl;
  uint32_t lo, hi;
  dropout_words(rk, static_cast<uint32_t>(elem_div4), static_cast<uint32_t>(elem_div4 >> 32) ^ static_cast<uint32_t>(seed >> 32), lo, hi);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}
// Whole-step CUDA graphs bake every kernel argument, the dropout seeds included, so the step-to-step variation of the masks
// comes from DEVICE memory: when the library has a seed source (tvt_set_seed_source: a device-resident step counter that a
// one-thread kernel at the head of the captured step advances) its hash is folded into the high word of every seed.  Only the
// high word changes, so the round keys that some kernels hoist to launch constants (low word) stay valid.
__device__ __forceinline__ unsigned long long mix_seed(unsigned long long seed, const unsigned long long* src) {
  if (src == nullptr) return seed;
  const unsigned long long h = __ldg(src) * 0x9E3779B97F4A7C15ull;
  return seed ^ (h & 0xFFFFFFFF00000000ull);
}

// Lane i (0..3) survives dropout when its 16 random bits are >= thr16 (= p * 65536).
__device__ __forceinline__ bool dropout_keep_lane(uint64_t bits, int i, uint32_t thr16) {
  return (static_cast<uint32_t>(bits >> (16 * i)) & 0xFFFFu) >= thr16;
}
__device__ __forceinline__ bool dropout_keep(uint64_t seed, uint64_t elem, uint32_t thr16) {
  return dropout_keep_lane(dropout_bits4(seed, elem >> 2), static_cast<int>(elem & 3), thr16);
}

// Attention-probability dropout: the mask of (clip*head bh, query row i, key j) is lane (j & 3) of the
// 64 bits hashed from ((bh * Sq + i) * ceil(Sk / 4) + j / 4), so a thread that owns a query row needs one
// hash per 4 consecutive keys.  Shared by the CUDA-core and the tcgen05 kernels (forward of one may be
// paired with the backward of the other).
__device__ __forceinline__ uint64_t attn_drop_bits(uint64_t seed, uint64_t rowkey, int j4) {
  return dropout_bits4(seed, rowkey + static_cast<uint64_t>(j4));
}
__device__ __forceinline__ uint64_t attn_rowkey(long long bh, int sq, int sk, int i) {
  return (static_cast<uint64_t>(bh) * sq + i) * static_cast<uint64_t>((sk + 3) >> 2);
}

#endif  // __CUDACC__

}  // namespace tvt
