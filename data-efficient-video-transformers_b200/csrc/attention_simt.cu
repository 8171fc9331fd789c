// fp32 CUDA-core attention, forward and backward: the exact-arithmetic path used by the fp32 parity
// mode (and as a correctness anchor for the tcgen05 attention kernel in attention_sm100.cu).
// One CTA per (batch, head); Q / K / V / dO of that head are staged in shared memory as fp32, one warp
// per query row (forward, backward pass 1) or per key row (backward pass 2); nothing is materialised
// in global memory besides the log-sum-exp vector.  See include/tvt.h: tvt_attention_fwd / _bwd.
#include "tvt_common.cuh"

namespace tvt {
namespace attn_simt {

constexpr int kWarps = 8;

struct Params {
  const void* q; const void* k; const void* v; const void* o; const void* d_o;
  void* out; float* lse;
  void* dq; void* dk; void* dv;
  int B, H, Sq, Sk, hd;
  long long ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv;
  float scale;
  float dropout_scale; unsigned dropout_thr16; unsigned long long dropout_seed; const unsigned long long* seed_src;
};

// Stage rows [0, S) x [0, hd) of a head into smem (pitch hd + 4 floats) as fp32.
template <typename T>
__device__ __forceinline__ void stage(float* dst, const T* src, long long ld, int S, int hd, int pitch) {
  constexpr int V = Vec16<T>::kN;
  const int per_row = hd / V;
  for (int i = threadIdx.x; i < S * per_row; i += blockDim.x) {
    const int r = i / per_row, c = (i - r * per_row) * V;
    float v[V];
    Vec16<T>::load(src + r * ld + c, v);
#pragma unroll
    for (int j = 0; j < V; ++j) dst[r * pitch + c + j] = v[j];
  }
}

__device__ __forceinline__ float dot_row(const float* a, const float* b, int hd) {
  float acc = 0.0f;
  for (int c = 0; c < hd; c += 4) {
    const float4 x = *reinterpret_cast<const float4*>(a + c);
    const float4 y = *reinterpret_cast<const float4*>(b + c);
    acc += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
  }
  return acc;
}

// Dropout on attention probabilities (mask layout: tvt_common.cuh, attn_drop_bits).
__device__ __forceinline__ float drop_mul(const Params& p, long long bh, int i, int j) {
  if (!p.dropout_thr16) return 1.0f;
  const uint64_t bits = attn_drop_bits(mix_seed(p.dropout_seed, p.seed_src), attn_rowkey(bh, p.Sq, p.Sk, i), j >> 2);
  return dropout_keep_lane(bits, j & 3, p.dropout_thr16) ? p.dropout_scale : 0.0f;
}

template <typename T>
__global__ void __launch_bounds__(kWarps * 32) fwd_kernel(const Params p) {
  extern __shared__ __align__(16) float smem[];
  const int pitch = p.hd + 4;
  float* sQ = smem;
  float* sK = sQ + p.Sq * pitch;
  float* sV = sK + p.Sk * pitch;
  float* sP = sV + p.Sk * pitch;  // [kWarps][Sk]
  const int bh = blockIdx.x, b = bh / p.H, h = bh % p.H;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const T* q = reinterpret_cast<const T*>(p.q) + static_cast<long long>(b) * p.Sq * p.ldq + h * p.hd;
  const T* k = reinterpret_cast<const T*>(p.k) + static_cast<long long>(b) * p.Sk * p.ldk + h * p.hd;
  const T* v = reinterpret_cast<const T*>(p.v) + static_cast<long long>(b) * p.Sk * p.ldv + h * p.hd;
  stage<T>(sQ, q, p.ldq, p.Sq, p.hd, pitch);
  stage<T>(sK, k, p.ldk, p.Sk, p.hd, pitch);
  stage<T>(sV, v, p.ldv, p.Sk, p.hd, pitch);
  __syncthreads();
  float* myP = sP + warp * p.Sk;
  for (int i = warp; i < p.Sq; i += kWarps) {
    float mx = -INFINITY;
    for (int j = lane; j < p.Sk; j += 32) {
      const float s = dot_row(sQ + i * pitch, sK + j * pitch, p.hd) * p.scale;
      myP[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.0f;
    for (int j = lane; j < p.Sk; j += 32) {
      const float e = __expf(myP[j] - mx);
      sum += e;
      myP[j] = e * drop_mul(p, bh, i, j);
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    if (lane == 0 && p.lse) p.lse[static_cast<long long>(bh) * p.Sq + i] = mx + __logf(sum);
    __syncwarp();
    T* o = reinterpret_cast<T*>(p.out) + (static_cast<long long>(b) * p.Sq + i) * p.ldo + h * p.hd;
    for (int c = lane; c < p.hd; c += 32) {
      float acc = 0.0f;
      for (int j = 0; j < p.Sk; ++j) acc += myP[j] * sV[j * pitch + c];
      o[c] = Elem<T>::from_f(acc * inv);
    }
    __syncwarp();
  }
}

// Backward.  D_i = sum_c dO_ic * O_ic;  P_ij = exp(s_ij - lse_i);  with probability dropout m_ij:
//   dV_j = sum_i (P_ij m_ij) dO_i;  dP_ij = (dO_i . V_j) m_ij;  dS_ij = P_ij (dP_ij - D_i)
//   dQ_i = scale * sum_j dS_ij K_j;  dK_j = scale * sum_i dS_ij Q_i
template <typename T>
__global__ void __launch_bounds__(kWarps * 32) bwd_kernel(const Params p) {
  extern __shared__ __align__(16) float smem[];
  const int pitch = p.hd + 4;
  float* sQ = smem;
  float* sdO = sQ + p.Sq * pitch;
  float* sK = sdO + p.Sq * pitch;
  float* sV = sK + p.Sk * pitch;
  float* sD = sV + p.Sk * pitch;                         // [Sq]
  float* sL = sD + p.Sq;                                 // [Sq]
  float* sP = sL + p.Sq;                                 // [kWarps][max(Sq, Sk)]
  const int smax = p.Sq > p.Sk ? p.Sq : p.Sk;
  const int bh = blockIdx.x, b = bh / p.H, h = bh % p.H;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long qrow0 = static_cast<long long>(b) * p.Sq, krow0 = static_cast<long long>(b) * p.Sk;
  stage<T>(sQ, reinterpret_cast<const T*>(p.q) + qrow0 * p.ldq + h * p.hd, p.ldq, p.Sq, p.hd, pitch);
  stage<T>(sdO, reinterpret_cast<const T*>(p.d_o) + qrow0 * p.lddo + h * p.hd, p.lddo, p.Sq, p.hd, pitch);
  stage<T>(sK, reinterpret_cast<const T*>(p.k) + krow0 * p.ldk + h * p.hd, p.ldk, p.Sk, p.hd, pitch);
  stage<T>(sV, reinterpret_cast<const T*>(p.v) + krow0 * p.ldv + h * p.hd, p.ldv, p.Sk, p.hd, pitch);
  __syncthreads();
  // D_i from O (global) and dO (smem)
  for (int i = warp; i < p.Sq; i += kWarps) {
    const T* o = reinterpret_cast<const T*>(p.o) + (qrow0 + i) * p.ldo + h * p.hd;
    float acc = 0.0f;
    for (int c = lane; c < p.hd; c += 32) acc += Elem<T>::to_f(o[c]) * sdO[i * pitch + c];
    acc = warp_sum(acc);
    if (lane == 0) {
      sD[i] = acc;
      sL[i] = p.lse[static_cast<long long>(bh) * p.Sq + i];
    }
  }
  __syncthreads();
  float* my = sP + warp * smax;
  // pass 1: one warp per query row -> dQ_i
  for (int i = warp; i < p.Sq; i += kWarps) {
    for (int j = lane; j < p.Sk; j += 32) {
      const float s = dot_row(sQ + i * pitch, sK + j * pitch, p.hd) * p.scale;
      const float pr = __expf(s - sL[i]);
      const float dp = dot_row(sdO + i * pitch, sV + j * pitch, p.hd) * drop_mul(p, bh, i, j);
      my[j] = pr * (dp - sD[i]);
    }
    __syncwarp();
    T* dq = reinterpret_cast<T*>(p.dq) + (qrow0 + i) * p.lddq + h * p.hd;
    for (int c = lane; c < p.hd; c += 32) {
      float acc = 0.0f;
      for (int j = 0; j < p.Sk; ++j) acc += my[j] * sK[j * pitch + c];
      dq[c] = Elem<T>::from_f(acc * p.scale);
    }
    __syncwarp();
  }
  // pass 2: one warp per key row -> dK_j, dV_j (two smem vectors per warp: dS column and P*m column)
  float* my2 = sP + (kWarps + warp) * smax;
  for (int j = warp; j < p.Sk; j += kWarps) {
    for (int i = lane; i < p.Sq; i += 32) {
      const float s = dot_row(sQ + i * pitch, sK + j * pitch, p.hd) * p.scale;
      const float pr = __expf(s - sL[i]);
      const float m = drop_mul(p, bh, i, j);
      const float dp = dot_row(sdO + i * pitch, sV + j * pitch, p.hd) * m;
      my[i] = pr * (dp - sD[i]);
      my2[i] = pr * m;
    }
    __syncwarp();
    T* dk = reinterpret_cast<T*>(p.dk) + (krow0 + j) * p.lddk + h * p.hd;
    T* dv = reinterpret_cast<T*>(p.dv) + (krow0 + j) * p.lddv + h * p.hd;
    for (int c = lane; c < p.hd; c += 32) {
      float ak = 0.0f, av = 0.0f;
      for (int i = 0; i < p.Sq; ++i) {
        ak += my[i] * sQ[i * pitch + c];
        av += my2[i] * sdO[i * pitch + c];
      }
      dk[c] = Elem<T>::from_f(ak * p.scale);
      dv[c] = Elem<T>::from_f(av);
    }
    __syncwarp();
  }
}

static size_t fwd_smem(const Params& p) {
  return sizeof(float) * (static_cast<size_t>(p.Sq + 2 * p.Sk) * (p.hd + 4) + static_cast<size_t>(kWarps) * p.Sk);
}
static size_t bwd_smem(const Params& p) {
  const int smax = p.Sq > p.Sk ? p.Sq : p.Sk;
  return sizeof(float) * (static_cast<size_t>(2 * p.Sq + 2 * p.Sk) * (p.hd + 4) + 2 * p.Sq + static_cast<size_t>(2 * kWarps) * smax);
}

template <typename K>
static int set_smem(K kern, size_t bytes, const char* what) {
  if (bytes > 227 * 1024) {
    set_last_error("%s: sequence/head too large for the shared-memory resident kernel (%zu bytes)", what, bytes);
    return TVT_EINVAL;
  }
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
    if (e != cudaSuccess) {
      set_last_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
      return TVT_ECUDA;
    }
  }
  return TVT_OK;
}

int launch_fwd(const Params& p, bool f32, cudaStream_t s) {
  const size_t bytes = fwd_smem(p);
  int rc;
  if (f32) {
    if ((rc = set_smem(fwd_kernel<float>, bytes, "tvt_attention_fwd")) != TVT_OK) return rc;
    fwd_kernel<float><<<p.B * p.H, kWarps * 32, bytes, s>>>(p);
  } else {
    if ((rc = set_smem(fwd_kernel<__nv_bfloat16>, bytes, "tvt_attention_fwd")) != TVT_OK) return rc;
    fwd_kernel<__nv_bfloat16><<<p.B * p.H, kWarps * 32, bytes, s>>>(p);
  }
  return check_launch("tvt_attention_fwd");
}

int launch_bwd(const Params& p, bool f32, cudaStream_t s) {
  const size_t bytes = bwd_smem(p);
  int rc;
  if (f32) {
    if ((rc = set_smem(bwd_kernel<float>, bytes, "tvt_attention_bwd")) != TVT_OK) return rc;
    bwd_kernel<float><<<p.B * p.H, kWarps * 32, bytes, s>>>(p);
  } else {
    if ((rc = set_smem(bwd_kernel<__nv_bfloat16>, bytes, "tvt_attention_bwd")) != TVT_OK) return rc;
    bwd_kernel<__nv_bfloat16><<<p.B * p.H, kWarps * 32, bytes, s>>>(p);
  }
  return check_launch("tvt_attention_bwd");
}

}  // namespace attn_simt
}  // namespace tvt

namespace tvt {
namespace attn_tc {  // attention_sm100.cu
bool supported(long long sq, long long sk, long long hd);
int launch_fwd(const tvt_attention_fwd_args* a, cudaStream_t s);
int launch_bwd(const tvt_attention_bwd_args* a, cudaStream_t s);
}  // namespace attn_tc
}  // namespace tvt

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// impl 0: tcgen05 kernel when it applies (bf16, head_dim 64, short sequences), else the CUDA-core kernel.
static int pick_impl(int impl, int dtype, long long sq, long long sk, long long hd, const char* who) {
  const bool tc_ok = dtype == TVT_BF16 && tvt::attn_tc::supported(sq, sk, hd);
  if (impl == 2 && !tc_ok) {
    tvt::set_last_error("%s: impl=2 (tcgen05) needs bf16, head_dim 64, sq <= 256, sk <= 272", who);
    return -1;
  }
  if (impl < 0 || impl > 2) {
    tvt::set_last_error("%s: bad impl %d", who, impl);
    return -1;
  }
  return impl == 0 ? (tc_ok ? 2 : 1) : impl;
}

extern "C" int tvt_attention_fwd(const tvt_attention_fwd_args* a, void* stream) {
  using namespace tvt;
  TVT_REQUIRE(a != nullptr, "tvt_attention_fwd: null args");
  TVT_REQUIRE(a->q && a->k && a->v && a->o, "tvt_attention_fwd: null pointer");
  TVT_REQUIRE(a->batch > 0 && a->heads > 0 && a->sq > 0 && a->sk > 0, "tvt_attention_fwd: empty problem");
  TVT_REQUIRE(a->head_dim > 0 && a->head_dim % 8 == 0, "tvt_attention_fwd: head_dim must be a multiple of 8");
  TVT_REQUIRE(a->dtype == TVT_BF16 || a->dtype == TVT_F32, "tvt_attention_fwd: bad dtype");
  const int64_t w = a->heads * a->head_dim;
  TVT_REQUIRE(a->ldq >= w && a->ldk >= w && a->ldv >= w && a->ldo >= w, "tvt_attention_fwd: row pitch smaller than heads*head_dim");
  TVT_REQUIRE(a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0 && a->ldo % 8 == 0, "tvt_attention_fwd: row pitches must be multiples of 8");
  TVT_REQUIRE(al16(a->q) && al16(a->k) && al16(a->v) && al16(a->o), "tvt_attention_fwd: pointers must be 16-byte aligned");
  TVT_REQUIRE(a->dropout_p >= 0.0f && a->dropout_p < 1.0f, "tvt_attention_fwd: dropout_p must be in [0,1)");
  TVT_REQUIRE(a->batch * a->heads < (1ll << 31), "tvt_attention_fwd: batch*heads too large");
  const int impl = pick_impl(a->impl, a->dtype, a->sq, a->sk, a->head_dim, "tvt_attention_fwd");
  if (impl < 0) return TVT_EINVAL;
  int rc = require_sm100();
  if (rc != TVT_OK) return rc;
  if (impl == 2) return attn_tc::launch_fwd(a, static_cast<cudaStream_t>(stream));
  attn_simt::Params p{};
  p.q = a->q; p.k = a->k; p.v = a->v; p.out = a->o; p.lse = a->lse;
  p.B = (int)a->batch; p.H = (int)a->heads; p.Sq = (int)a->sq; p.Sk = (int)a->sk; p.hd = (int)a->head_dim;
  p.ldq = a->ldq; p.ldk = a->ldk; p.ldv = a->ldv; p.ldo = a->ldo; p.scale = a->scale;
  if (a->dropout_p > 0.0f) {
    p.dropout_thr16 = (unsigned)(a->dropout_p * 65536.0f + 0.5f);
    p.dropout_scale = 65536.0f / (65536.0f - (float)p.dropout_thr16);
    p.dropout_seed = a->dropout_seed;
    p.seed_src = seed_source();
  }
  return attn_simt::launch_fwd(p, a->dtype == TVT_F32, static_cast<cudaStream_t>(stream));
}

extern "C" int tvt_attention_bwd(const tvt_attention_bwd_args* a, void* stream) {
  using namespace tvt;
  TVT_REQUIRE(a != nullptr, "tvt_attention_bwd: null args");
  TVT_REQUIRE(a->q && a->k && a->v && a->o && a->d_o && a->lse && a->dq && a->dk && a->dv, "tvt_attention_bwd: null pointer");
  TVT_REQUIRE(a->batch > 0 && a->heads > 0 && a->sq > 0 && a->sk > 0, "tvt_attention_bwd: empty problem");
  TVT_REQUIRE(a->head_dim > 0 && a->head_dim % 8 == 0, "tvt_attention_bwd: head_dim must be a multiple of 8");
  TVT_REQUIRE(a->dtype == TVT_BF16 || a->dtype == TVT_F32, "tvt_attention_bwd: bad dtype");
  const int64_t w = a->heads * a->head_dim;
  TVT_REQUIRE(a->ldq >= w && a->ldk >= w && a->ldv >= w && a->ldo >= w && a->lddo >= w && a->lddq >= w && a->lddk >= w && a->lddv >= w,
              "tvt_attention_bwd: row pitch smaller than heads*head_dim");
  TVT_REQUIRE((a->ldq | a->ldk | a->ldv | a->ldo | a->lddo | a->lddq | a->lddk | a->lddv) % 8 == 0, "tvt_attention_bwd: row pitches must be multiples of 8");
  TVT_REQUIRE(al16(a->q) && al16(a->k) && al16(a->v) && al16(a->o) && al16(a->d_o) && al16(a->dq) && al16(a->dk) && al16(a->dv),
              "tvt_attention_bwd: pointers must be 16-byte aligned");
  TVT_REQUIRE(a->dropout_p >= 0.0f && a->dropout_p < 1.0f, "tvt_attention_bwd: dropout_p must be in [0,1)");
  const int impl = pick_impl(a->impl, a->dtype, a->sq, a->sk, a->head_dim, "tvt_attention_bwd");
  if (impl < 0) return TVT_EINVAL;
  int rc = require_sm100();
  if (rc != TVT_OK) return rc;
  if (impl == 2) return attn_tc::launch_bwd(a, static_cast<cudaStream_t>(stream));
  attn_simt::Params p{};
  p.q = a->q; p.k = a->k; p.v = a->v; p.o = a->o; p.d_o = a->d_o; p.lse = const_cast<float*>(a->lse);
  p.dq = a->dq; p.dk = a->dk; p.dv = a->dv;
  p.B = (int)a->batch; p.H = (int)a->heads; p.Sq = (int)a->sq; p.Sk = (int)a->sk; p.hd = (int)a->head_dim;
  p.ldq = a->ldq; p.ldk = a->ldk; p.ldv = a->ldv; p.ldo = a->ldo; p.lddo = a->lddo; p.lddq = a->lddq; p.lddk = a->lddk; p.lddv = a->lddv;
  p.scale = a->scale;
  if (a->dropout_p > 0.0f) {
    p.dropout_thr16 = (unsigned)(a->dropout_p * 65536.0f + 0.5f);
    p.dropout_scale = 65536.0f / (65536.0f - (float)p.dropout_thr16);
    p.dropout_seed = a->dropout_seed;
    p.seed_src = seed_source();
  }
  return attn_simt::launch_bwd(p, a->dtype == TVT_F32, static_cast<cudaStream_t>(stream));
}
