// tcgen05 attention for head_dim 64 (bf16), forward and backward, self- and cross-modal (Sq != Sk).
//
// One CTA per (clip, head): 4 worker warps (thread == one query / key row == one TMEM lane) + 1 issue warp.
// Q, K, V (and dO) tiles arrive by TMA as 128-byte-swizzled rows of 64 bf16; because one head row is
// exactly one swizzle atom wide, the SAME shared-memory tile serves as a K-major operand (rows x hd
// contraction) and as an MN-major operand (contraction over rows), so no transposes are ever made:
//
//   forward   S  = Q K^T           A = Q  (K-major)   B = K  (K-major)    acc S  [128 x Sk]  TMEM
//             O  = P V             A = P  (K-major)   B = V  (MN-major)   acc O  [128 x 64]  TMEM
//   backward  S  = Q K^T, dP = dO V^T                 B = V  (K-major)
//             dV = P^T dO          A = P  (MN-major)  B = dO (MN-major)
//             dK = dS^T Q          A = dS (MN-major)  B = Q  (MN-major)
//             dQ = dS K            A = dS (K-major)   B = K  (MN-major)
//
// Softmax runs on the TMEM rows (tcgen05.ld, fp32, exp2 with the scale folded in); P / dS are written
// as bf16 into swizzled shared memory and consumed by the next MMA.  Rows / columns beyond the sequence
// are masked (the sequences here are 2^k + 1 tokens long: 17, 33, 65, 129).
#include <cuda.h>
#include <cstdlib>

#include <mutex>
#include <type_traits>
#include <set>

#include "tvt_common.cuh"
#include "tvt_ptx.cuh"

namespace tvt {
namespace attn_tc {

constexpr int HD = 64;        // head dim: one 128-byte swizzle atom
constexpr int kThreads = 160;     // backward: warps 0-3 one row per thread; warp 4 TMA / MMA issue
constexpr int kThreadsFwd = 192;  // forward: + warp 5 for the tail query rows
constexpr float kLog2e = 1.4426950408889634f;

struct Params {
  int B, H, Sq, Sk;
  int sk_pad;                 // Sk rounded up to 16
  int kv_box;                 // rows per K / V TMA box (divides sk_pad)
  const __nv_bfloat16* q_in; long long ldq;   // raw Q rows for the CUDA-core tail path
  float scale;
  float dropout_scale; unsigned dropout_thr16; unsigned long long dropout_seed; const unsigned long long* seed_src;
  __nv_bfloat16* o; long long ldo;
  float* lse;
  // backward
  const __nv_bfloat16* o_in; const __nv_bfloat16* do_in; long long lddo;
  __nv_bfloat16* dq; __nv_bfloat16* dk; __nv_bfloat16* dv; long long lddq, lddk, lddv;
  int use_tu;                 // fwd_small: tail-key / tail-row score vectors from two N = 16 MMAs instead of CUDA-core dot products
  int prefetch_stride;        // fwd_small / bwd3: CTA bh prefetches the operands of CTA bh + prefetch_stride into L2 (= resident CTAs of the grid)
};

// Byte offset of the 16-byte chunk `chunk` (0..7) of row `row` inside a [rows x 128 B] swizzled tile.
__device__ __forceinline__ uint32_t sw128(int row, int chunk) {
  return static_cast<uint32_t>(row) * 128u + (static_cast<uint32_t>(chunk ^ (row & 7)) << 4);
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lds_f32(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr) : "memory");
  return v;
}

__device__ __forceinline__ float drop_mul(const Params& p, long long bh, int i, int j) {
  if (!p.dropout_thr16) return 1.0f;
  const uint64_t bits = attn_drop_bits(mix_seed(p.dropout_seed, p.seed_src), attn_rowkey(bh, p.Sq, p.Sk, i), j >> 2);
  return dropout_keep_lane(bits, j & 3, p.dropout_thr16) ? p.dropout_scale : 0.0f;
}
// Dropout multipliers of 16 consecutive keys [c0, c0 + 16) of one query row (c0 % 16 == 0): 4 hashes.
__device__ __forceinline__ void drop_mul16(const Params& p, uint64_t rowkey, int c0, float (&m)[16]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const uint64_t bits = attn_drop_bits(mix_seed(p.dropout_seed, p.seed_src), rowkey, (c0 >> 2) + g);
#pragma unroll
    for (int i = 0; i < 4; ++i) m[4 * g + i] = dropout_keep_lane(bits, i, p.dropout_thr16) ? p.dropout_scale : 0.0f;
  }
}

// Issue D[tmem] = A(K-major tile, 128 rows) * B(K-major tile, n rows)^T over the 64-wide contraction.
__device__ __forceinline__ void mma_kk(uint32_t d_tmem, uint32_t a_smem, uint32_t b_smem, int n) {
  const uint32_t idesc = make_idesc_bf16(128, n, false, false);
#pragma unroll
  for (int k = 0; k < HD / 16; ++k)
    tc_mma_f16_ss(d_tmem, make_smem_desc_sw128(a_smem + k * 32, 16, 1024), make_smem_desc_sw128(b_smem + k * 32, 16, 1024), idesc, k > 0);
}

// ------------------------------------------------------------------------------------------- forward
// Warps 0-3: one query row per thread (TMEM lane == thread).  Warp 4, lane 0: issues every TMA and MMA.
// Warp 5: "tail" query rows on CUDA cores.  The sequences are 2^k + 1 tokens long, so a second 128-row
// MMA tile would hold a single valid row (S = 129); instead the <= kMaxTail rows beyond the last full
// tile are computed by warp 5 straight from the K / V tiles in shared memory while the tensor cores
// work on the main tile.
// smem: Q tile 16 KB | K sk_pad*128 | V sk_pad*128 | P ceil(sk_pad/64)*16 KB | barriers | tail scratch
constexpr int kMaxTail = 8;
constexpr int KC = 144;   // padded key capacity of one chunk: 128 + kMaxTail rounded up to 16

__device__ __forceinline__ int main_tiles(int sq) {   // number of 128-row tensor-core tiles
  const int full = sq / 128, rem = sq - full * 128;
  return (full >= 1 && rem <= kMaxTail) ? full : (sq + 127) / 128;
}

// bf16 element (row, col) of a [rows x 64] 128B-swizzled tile
__device__ __forceinline__ const __nv_bfloat16* sw_elem(const uint8_t* tile, int row, int col) {
  return reinterpret_cast<const __nv_bfloat16*>(tile + sw128(row, col >> 3)) + (col & 7);
}

__global__ void __launch_bounds__(kThreadsFwd) fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                          const __grid_constant__ CUtensorMap tmV, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int kv_bytes = p.sk_pad * 128;
  const int kblocks = (p.sk_pad + 63) / 64;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + 128 * 128;
  uint8_t* sV = sK + ((kv_bytes + 1023) & ~1023);
  uint8_t* sP = sV + ((kv_bytes + 1023) & ~1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + kblocks * 16384);
  uint64_t* bar_kv = bars;      // K and V landed
  uint64_t* bar_q = bars + 1;   // Q tile landed (one phase per m-tile)
  uint64_t* bar_s = bars + 2;   // S = Q K^T complete (one phase per m-tile)
  uint64_t* bar_o = bars + 3;   // O = P V complete   (one phase per m-tile)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  float* tail_q = reinterpret_cast<float*>(bars + 6);   // [64]
  float* tail_p = tail_q + HD;                          // [sk_pad]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool issuer = warp == 4 && lane == 0;
  const int bh = blockIdx.x, b = bh / p.H, h = bh % p.H;
  const int o_col = (p.sk_pad + 63) & ~63;           // O accumulator starts on a 64-column boundary
  const uint32_t tmem_cols = o_col + HD <= 128 ? 128u : (o_col + HD <= 256 ? 256u : 512u);
  const int m_tiles = main_tiles(p.Sq);

  if (issuer) {
    mbar_init(smem_u32(bar_kv), 1);
    mbar_init(smem_u32(bar_q), 1);
    mbar_init(smem_u32(bar_s), 1);
    mbar_init(smem_u32(bar_o), 1);
    fence_mbar_init();
  }
  if (warp == 4) {
    tmem_alloc(smem_u32(tmem_slot), tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;                 // [128 x sk_pad] fp32
  const uint32_t tmem_O = tmem_base + o_col;         // [128 x 64] fp32
  const float sl2 = p.scale * kLog2e;

  if (issuer) {
    mbar_arrive_expect_tx(smem_u32(bar_kv), 2 * kv_bytes);
    for (int r = 0; r < p.sk_pad; r += p.kv_box) {
      tma_load_2d(smem_u32(sK + r * 128), &tmK, smem_u32(bar_kv), h * HD, b * p.Sk + r);
      tma_load_2d(smem_u32(sV + r * 128), &tmV, smem_u32(bar_kv), h * HD, b * p.Sk + r);
    }
  }

  if (warp == 5) {
    // ---- tail query rows [128 * m_tiles, Sq): scores, softmax and P V on CUDA cores
    const int t0 = m_tiles * 128;
    if (t0 < p.Sq) {
      mbar_wait(smem_u32(bar_kv), 0);
      for (int row = t0; row < p.Sq; ++row) {
        const __nv_bfloat16* qrow = p.q_in + (static_cast<long long>(b) * p.Sq + row) * p.ldq + h * HD;
        const float2 q2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(qrow + 2 * lane));
        tail_q[2 * lane] = q2.x;
        tail_q[2 * lane + 1] = q2.y;
        __syncwarp();
        float mx = -INFINITY;
        for (int j = lane; j < p.Sk; j += 32) {
          float acc = 0.0f;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float kf[8];
            Vec16<__nv_bfloat16>::unpack(*reinterpret_cast<const uint4*>(sK + sw128(j, c)), kf);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc += tail_q[c * 8 + i] * kf[i];
          }
          tail_p[j] = acc;
          mx = fmaxf(mx, acc);
        }
        mx = warp_max(mx);
        float sum = 0.0f;
        for (int j = lane; j < p.Sk; j += 32) {
          const float e = exp2f((tail_p[j] - mx) * sl2);
          sum += e;
          tail_p[j] = e * drop_mul(p, bh, row, j);
        }
        sum = warp_sum(sum);
        __syncwarp();
        float o0 = 0.0f, o1 = 0.0f;
        for (int j = 0; j < p.Sk; ++j) {
          const float pj = tail_p[j];
          const float2 v2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sw_elem(sV, j, 2 * lane)));
          o0 += pj * v2.x;
          o1 += pj * v2.y;
        }
        const float inv = 1.0f / sum;
        __nv_bfloat16* orow = p.o + (static_cast<long long>(b) * p.Sq + row) * p.ldo + h * HD;
        *reinterpret_cast<__nv_bfloat162*>(orow + 2 * lane) = __floats2bfloat162_rn(o0 * inv, o1 * inv);
        if (lane == 0 && p.lse) p.lse[static_cast<long long>(bh) * p.Sq + row] = mx * p.scale + __logf(sum);
        __syncwarp();
      }
    }
  }

  uint32_t phase = 0;
  for (int mt = 0; mt < m_tiles; ++mt, phase ^= 1) {
    if (issuer) {
      mbar_arrive_expect_tx(smem_u32(bar_q), 128 * 128);
      tma_load_2d(smem_u32(sQ), &tmQ, smem_u32(bar_q), h * HD, b * p.Sq + mt * 128);
      if (mt == 0) mbar_wait(smem_u32(bar_kv), 0);
      mbar_wait(smem_u32(bar_q), phase);
      tc_fence_after();
      for (int n0 = 0; n0 < p.sk_pad; n0 += 256) {   // S = Q K^T in column chunks of <= 256 keys
        const int n = p.sk_pad - n0 < 256 ? p.sk_pad - n0 : 256;
        mma_kk(tmem_S + n0, smem_u32(sQ), smem_u32(sK + n0 * 128), n);
      }
      tc_commit(smem_u32(bar_s));
    }
    const int row = mt * 128 + tid;                          // query row of this thread (warps 0-3)
    const bool row_ok = warp < 4 && row < p.Sq;
    const bool warp_ok = warp < 4 && mt * 128 + warp * 32 < p.Sq;   // warp-uniform: some valid row in this warp
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
    float mx = -INFINITY, sum = 0.0f;
    if (warp_ok) {
      mbar_wait(smem_u32(bar_s), phase);
      tc_fence_after();
      // pass 1: row maximum (32 columns per TMEM load, 16 for the ragged end)
      int c0 = 0;
      for (; c0 + 32 <= p.sk_pad; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_S + lane_addr + c0, r);
        tmem_ld_wait_dep(r);
        if (c0 + 32 <= p.Sk) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i < p.Sk) mx = fmaxf(mx, __uint_as_float(r[i]));
        }
      }
      if (c0 < p.sk_pad) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(tmem_S + lane_addr + c0, r);
        tmem_ld_wait_dep(r);
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (c0 + i < p.Sk) mx = fmaxf(mx, __uint_as_float(r[i]));
      }
      const float mxs = mx * sl2;
      const uint64_t rowkey = attn_rowkey(bh, p.Sq, p.Sk, row_ok ? row : 0);
      // pass 2: P = exp2(s * scale * log2e - max), bf16, into the swizzled A-operand tile
      for (c0 = 0; c0 < p.sk_pad; c0 += 16) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(tmem_S + lane_addr + c0, r);
        tmem_ld_wait_dep(r);
        uint32_t packed[8];
        if (c0 + 16 <= p.Sk && !p.dropout_thr16) {
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const float e0 = exp2f(__uint_as_float(r[i]) * sl2 - mxs), e1 = exp2f(__uint_as_float(r[i + 1]) * sl2 - mxs);
            sum += e0 + e1;
            packed[i >> 1] = pack_bf16x2(e0, e1);
          }
        } else {
          float m[16];
          if (p.dropout_thr16) {
            drop_mul16(p, rowkey, c0, m);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) m[i] = 1.0f;
          }
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            float e0 = 0.0f, e1 = 0.0f;
            if (c0 + i < p.Sk) { e0 = exp2f(__uint_as_float(r[i]) * sl2 - mxs); sum += e0; e0 *= m[i]; }
            if (c0 + i + 1 < p.Sk) { e1 = exp2f(__uint_as_float(r[i + 1]) * sl2 - mxs); sum += e1; e1 *= m[i + 1]; }
            packed[i >> 1] = pack_bf16x2(e0, e1);
          }
        }
        const uint32_t blk = smem_u32(sP + (c0 >> 6) * 16384);
        const int ch = (c0 & 63) >> 3;
        sts128(blk + sw128(tid, ch), packed[0], packed[1], packed[2], packed[3]);
        sts128(blk + sw128(tid, ch + 1), packed[4], packed[5], packed[6], packed[7]);
      }
      fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
    }
    // (rows of warps beyond the sequence keep stale P: each P row only feeds its own, never stored, O row)
    tc_fence_before();
    __syncthreads();
    if (issuer) {
      tc_fence_after();
      const uint32_t idesc = make_idesc_bf16(128, HD, false, true);
      for (int k = 0; k < p.sk_pad / 16; ++k) {     // O = P V: contraction over keys in steps of 16
        const uint32_t a = smem_u32(sP + (k >> 2) * 16384) + (k & 3) * 32;
        const uint32_t bb = smem_u32(sV) + k * 16 * 128;
        tc_mma_f16_ss(tmem_O, make_smem_desc_sw128(a, 16, 1024), make_smem_desc_sw128(bb, 8192, 1024), idesc, k > 0);
      }
      tc_commit(smem_u32(bar_o));
    }
    if (warp_ok) {
      mbar_wait(smem_u32(bar_o), phase);
      tc_fence_after();
      const float inv = 1.0f / sum;
      __nv_bfloat16* orow = p.o + (static_cast<long long>(b) * p.Sq + (row_ok ? row : 0)) * p.ldo + h * HD;
#pragma unroll
      for (int c0 = 0; c0 < HD; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_O + lane_addr + c0, r);
        tmem_ld_wait_dep(r);
        if (row_ok) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint32_t pk[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) pk[i] = pack_bf16x2(__uint_as_float(r[8 * g + 2 * i]) * inv, __uint_as_float(r[8 * g + 2 * i + 1]) * inv);
            *reinterpret_cast<uint4*>(orow + c0 + 8 * g) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
      }
      if (row_ok && p.lse) p.lse[static_cast<long long>(bh) * p.Sq + row] = mx * p.scale + __logf(sum);
    }
    tc_fence_before();
    __syncthreads();   // TMEM S/O, sQ and sP are reused by the next m-tile
    tc_fence_after();
  }
  if (m_tiles == 0) __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------------------------------- forward, small
// Sq <= 128 + kMaxTail AND Sk <= 128 + kSmallTailKeys (every self-attention of the model family: S = 17 / 33 / 65 / 129).
// Same roles as fwd_kernel, but sized for FOUR resident CTAs per SM (the kernel is a latency chain, so
// residency is what buys throughput): the <= 8 tail KEYS are handled per thread on CUDA cores as well, which
// keeps S at 128 TMEM columns (O then reuses S's first 64 columns: 128-column allocation), and P overwrites
// the Q tile and the first 16 KB of the K tile once S has been computed (48 KB of tiles per CTA).
constexpr int kSmallTailKeys = 2;   // per-thread tail keys are unrolled: keep the kernel small (S = 2^k + 1 needs 1)

__global__ void __launch_bounds__(kThreads, 4) fwd_small_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmK128,
                                                                   const __grid_constant__ CUtensorMap tmK8, const __grid_constant__ CUtensorMap tmQ8,
                                                                   const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                       // later: P block 0 (keys 0..63)
  uint8_t* sK = sQ + 16384;                 // later: P block 1 (keys 64..127) over rows 0..127; tail key rows stay
  uint8_t* sV = sK + KC * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + KC * 128);
  uint64_t* bar_kv = bars;
  uint64_t* bar_q = bars + 1;
  uint64_t* bar_s = bars + 2;
  uint64_t* bar_o = bars + 3;
  uint64_t* bar_t = bars + 4;     // tail-score MMAs done (see below)
  uint64_t* bar_tu = bars + 6;    // ... and read out of TMEM by the four worker warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  float* tail_q = reinterpret_cast<float*>(bars + 8);   // [8 + 4 * 64] reduction scratch of the tail-row path
  float* tail_p = tail_q + 8 + 4 * HD;                  // [KC] probabilities of the current tail row

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool issuer = warp == 4 && lane == 0;
  const int bh = blockIdx.x, b = bh / p.H, h = bh % p.H;
  const int n_rows = p.Sq < 128 ? p.Sq : 128;      // query rows on the tensor cores
  const int tq_rows = p.Sq - n_rows;               // tail query rows (warp 5)
  const int n_keys = p.Sk < 128 ? p.Sk : 128;      // keys on the tensor cores
  const int tk = p.Sk - n_keys;                    // tail keys (per thread)
  const int n_mma = (n_keys + 15) & ~15;
  const float sl2 = p.scale * kLog2e;
  // The "+1" of S = 2^k + 1: the tail key's score column (q_i . k_128 for every main row i) and the tail query row's score row
  // (k_j . q_128 for every main key j) are dot products over head_dim that cost a quarter of this kernel's instructions on CUDA
  // cores.  With use_tu the tensor cores produce both from ONE 16-row B block = { K rows 128..135 | Q rows 128..135 } (rows 128..143
  // of the K tile): U = Q_main B^T (columns 0..7 = tail keys) and T = K_main B^T (columns 8..15 = tail query rows), two N = 16
  // MMAs issued ahead of the main S tile into TMEM columns [0, 32), read into registers by the workers, then overwritten by S.
  const bool tu = p.use_tu != 0;
  // 32-bit shared-space addresses of the tiles: the CUDA-core paths below read them with ld.shared (a generic-pointer
  // dereference costs 64-bit address arithmetic per load, which was a fifth of this kernel's instructions)
  const uint32_t sQ_s = smem_u32(sQ), sK_s = smem_u32(sK), sV_s = smem_u32(sV), tail_p_s = smem_u32(tail_p);

  if (issuer) {
    mbar_init(smem_u32(bar_kv), 1);
    mbar_init(smem_u32(bar_q), 1);
    mbar_init(smem_u32(bar_s), 1);
    mbar_init(smem_u32(bar_o), 1);
    mbar_init(smem_u32(bar_t), 1);
    mbar_init(smem_u32(bar_tu), 4);
    fence_mbar_init();
  }
  if (warp == 4) {
    tmem_alloc(smem_u32(tmem_slot), 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
  pdl_wait();      // barrier init and TMEM allocation overlapped the previous kernel's tail
  pdl_trigger();

  if (issuer) {
    if (tu) {   // K rows 0..135 (main + tail keys), Q tail rows into rows 136..143 of the K tile, V as usual
      mbar_arrive_expect_tx(smem_u32(bar_kv), (136 + p.sk_pad) * 128);
      tma_load_2d(smem_u32(sK), &tmK128, smem_u32(bar_kv), h * HD, b * p.Sk);
      tma_load_2d(smem_u32(sK + 128 * 128), &tmK8, smem_u32(bar_kv), h * HD, b * p.Sk + 128);
      for (int r = 0; r < p.sk_pad; r += p.kv_box) tma_load_2d(smem_u32(sV + r * 128), &tmV, smem_u32(bar_kv), h * HD, b * p.Sk + r);
      mbar_arrive_expect_tx(smem_u32(bar_q), (128 + 8) * 128);
      tma_load_2d(smem_u32(sQ), &tmQ, smem_u32(bar_q), h * HD, b * p.Sq);
      tma_load_2d(smem_u32(sK + 136 * 128), &tmQ8, smem_u32(bar_q), h * HD, b * p.Sq + 128);
    } else {
      mbar_arrive_expect_tx(smem_u32(bar_kv), 2 * p.sk_pad * 128);
      for (int r = 0; r < p.sk_pad; r += p.kv_box) {
        tma_load_2d(smem_u32(sK + r * 128), &tmK, smem_u32(bar_kv), h * HD, b * p.Sk + r);
        tma_load_2d(smem_u32(sV + r * 128), &tmV, smem_u32(bar_kv), h * HD, b * p.Sk + r);
      }
      mbar_arrive_expect_tx(smem_u32(bar_q), 128 * 128);
      tma_load_2d(smem_u32(sQ), &tmQ, smem_u32(bar_q), h * HD, b * p.Sq);
    }
    // pull the tiles of the CTA that will take this one's place on the SM into L2, so that the successor's TMA loads complete at L2
    // rather than DRAM latency.  Measured at C5 (tools/attn_time.py, inputs larger than L2): 66-73 us with or without it - the chain
    // is not bound by that latency; kept because it is four instructions (TVT_ATTN_FWD_PREFETCH=0 turns it off)
    const int nb = bh + p.prefetch_stride;
    if (p.prefetch_stride > 0 && nb < static_cast<int>(gridDim.x)) {
      const int b2 = nb / p.H, h2 = nb % p.H;
      if (tu) tma_prefetch_2d(&tmK128, h2 * HD, b2 * p.Sk);
      else for (int r = 0; r < p.sk_pad; r += p.kv_box) tma_prefetch_2d(&tmK, h2 * HD, b2 * p.Sk + r);
      for (int r = 0; r < p.sk_pad; r += p.kv_box) tma_prefetch_2d(&tmV, h2 * HD, b2 * p.Sk + r);
      tma_prefetch_2d(&tmQ, h2 * HD, b2 * p.Sq);
    }
    mbar_wait(smem_u32(bar_kv), 0);
    mbar_wait(smem_u32(bar_q), 0);
    tc_fence_after();
    if (tu) {
      mma_kk(tmem_base, smem_u32(sQ), smem_u32(sK + 128 * 128), 16);        // U: columns [0, 16)
      mma_kk(tmem_base + 16, smem_u32(sK), smem_u32(sK + 128 * 128), 16);   // T: columns [16, 32)
      tc_commit(smem_u32(bar_t));
      mbar_wait(smem_u32(bar_tu), 0);                                       // the workers hold U / T in registers
      tc_fence_after();
    }
    mma_kk(tmem_base, smem_u32(sQ), smem_u32(sK), n_mma);      // S = Q K^T over the main keys
    tc_commit(smem_u32(bar_s));
  }

  // tail scores from the tensor cores: this thread's row of U (tail keys) and of T (tail query rows)
  float u_tail[kSmallTailKeys], t_tail[kMaxTail];
#pragma unroll
  for (int t = 0; t < kSmallTailKeys; ++t) u_tail[t] = 0.0f;
#pragma unroll
  for (int t = 0; t < kMaxTail; ++t) t_tail[t] = 0.0f;
  if (tu && warp < 4) {
    mbar_wait(smem_u32(bar_t), 0);
    tc_fence_after();
    uint32_t ru[16], rt[16];
    tmem_ld_32x32b_x16(tmem_base + lane_addr, ru);
    tmem_ld_32x32b_x16(tmem_base + lane_addr + 16, rt);
    tmem_ld_wait_dep(ru, rt);
#pragma unroll
    for (int t = 0; t < kSmallTailKeys; ++t) u_tail[t] = __uint_as_float(ru[t]);
#pragma unroll
    for (int t = 0; t < kMaxTail; ++t) t_tail[t] = __uint_as_float(rt[8 + t]);
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(bar_tu));
  }

  if (warp < 4 && tq_rows > 0) {
    // ---- tail query rows, key-parallel on the 128 worker threads (thread j owns key j, and tail key 128 + j),
    //      done while the tensor cores compute S for the main tile.  Spreading this over the four worker warps
    //      (instead of a dedicated warp) keeps the four SM sub-partitions evenly loaded: every CTA's extra warp
    //      would land on the same sub-partition, which became the bottleneck with four resident CTAs.
    float* red = tail_q;                 // [8]   cross-warp max / sum
    float* part = tail_q + 8;            // [4][64] P V partials of the four warps
    // the tail query row comes straight from global memory: fetch it before waiting for the K / V tiles
    uint4 qraw[8];
    const bool need_q = !tu || tid < tk;       // with U / T only the tail-key x tail-row corner is still a CUDA-core dot product
    if (need_q) {
      const __nv_bfloat16* qrow = p.q_in + (static_cast<long long>(b) * p.Sq + n_rows) * p.ldq + h * HD;
#pragma unroll
      for (int c = 0; c < 8; ++c) qraw[c] = __ldg(reinterpret_cast<const uint4*>(qrow) + c);
    }
    mbar_wait(smem_u32(bar_kv), 0);
    for (int t = 0; t < tq_rows; ++t) {
      const int row = n_rows + t;
      if (t > 0 && need_q) {
        const __nv_bfloat16* qrow = p.q_in + (static_cast<long long>(b) * p.Sq + row) * p.ldq + h * HD;
#pragma unroll
        for (int c = 0; c < 8; ++c) qraw[c] = __ldg(reinterpret_cast<const uint4*>(qrow) + c);
      }
      float s0 = 0.0f, s1 = 0.0f;
      if (tu) {
        s0 = t_tail[0];
#pragma unroll
        for (int i = 1; i < kMaxTail; ++i) s0 = t == i ? t_tail[i] : s0;
        if (tid < tk) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float qf[8], kf[8];
            Vec16<__nv_bfloat16>::unpack(qraw[c], qf);
            Vec16<__nv_bfloat16>::unpack(lds128(sK_s + sw128(128 + tid, c)), kf);
#pragma unroll
            for (int i = 0; i < 8; ++i) s1 += qf[i] * kf[i];
          }
        }
      } else
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float qf[8], kf[8];
        Vec16<__nv_bfloat16>::unpack(qraw[c], qf);
        Vec16<__nv_bfloat16>::unpack(lds128(sK_s + sw128(tid, c)), kf);
#pragma unroll
        for (int i = 0; i < 8; ++i) s0 += qf[i] * kf[i];
        if (tid < tk) {
          Vec16<__nv_bfloat16>::unpack(lds128(sK_s + sw128(128 + tid, c)), kf);
#pragma unroll
          for (int i = 0; i < 8; ++i) s1 += qf[i] * kf[i];
        }
      }
      const bool k0ok = tid < n_keys, k1ok = tid < tk;
      float mxt = warp_max(fmaxf(k0ok ? s0 : -INFINITY, k1ok ? s1 : -INFINITY));
      if (lane == 0) red[warp] = mxt;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mxt = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
      const float e0 = k0ok ? ex2_approx((s0 - mxt) * sl2) : 0.0f, e1 = k1ok ? ex2_approx((s1 - mxt) * sl2) : 0.0f;
      const float ws = warp_sum(e0 + e1);
      if (lane == 0) red[4 + warp] = ws;
      if (k0ok) tail_p[tid] = e0 * drop_mul(p, bh, row, tid);
      if (k1ok) tail_p[128 + tid] = e1 * drop_mul(p, bh, row, 128 + tid);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const float sumt = red[4] + red[5] + red[6] + red[7];
      // P V: lane = (key group, 16-byte column chunk); a thread takes every 16th key with a 128-bit V read, the warp's four
      // key groups are folded with two shuffle steps and the four warps through shared memory
      {
        const int ch = lane & 7;
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
        for (int j = warp * 4 + (lane >> 3); j < p.Sk; j += 16) {
          const float pj = lds_f32(tail_p_s + 4 * j);
          float vf[8];
          Vec16<__nv_bfloat16>::unpack(lds128(sV_s + sw128(j, ch)), vf);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = fmaf(pj, vf[i], acc[i]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
          acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
        }
        if (lane < 8) {
          *reinterpret_cast<float4*>(part + warp * HD + ch * 8) = make_float4(acc[0], acc[1], acc[2], acc[3]);
          *reinterpret_cast<float4*>(part + warp * HD + ch * 8 + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (tid < HD) {
        __nv_bfloat16* orow = p.o + (static_cast<long long>(b) * p.Sq + row) * p.ldo + h * HD;
        orow[tid] = __float2bfloat16_rn((part[tid] + part[HD + tid] + part[2 * HD + tid] + part[3 * HD + tid]) / sumt);
        if (tid == 0 && p.lse) p.lse[static_cast<long long>(bh) * p.Sq + row] = mxt * p.scale + __logf(sumt);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");   // scratch reused by the next tail row
    }
  }

  const bool row_ok = warp < 4 && tid < n_rows;
  const bool warp_ok = warp < 4 && warp * 32 < n_rows;
  float mx = -INFINITY, sum = 0.0f;
  float st[kSmallTailKeys];  // raw scores, then probabilities, of the tail keys for this thread's row
  if (warp_ok) {
    mbar_wait(smem_u32(bar_s), 0);
    tc_fence_after();
    // tail-key scores on CUDA cores: q_i (own row of the Q tile) . k_t
#pragma unroll
    for (int t = 0; t < kSmallTailKeys; ++t) {
      st[t] = -INFINITY;
      if (t < tk) {
        float acc = 0.0f;
        if (tu) {
          acc = u_tail[t];
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float qf[8], kf[8];
            Vec16<__nv_bfloat16>::unpack(lds128(sQ_s + sw128(tid, c)), qf);
            Vec16<__nv_bfloat16>::unpack(lds128(sK_s + sw128(128 + t, c)), kf);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc += qf[i] * kf[i];
          }
        }
        st[t] = acc;
        mx = fmaxf(mx, acc);
      }
    }
    {                                              // pass 1: row maximum over the main keys (full blocks carry no per-key predicate)
      int c0 = 0;
#pragma unroll 1
      for (; c0 + 32 <= n_keys; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + lane_addr + c0, r);
        tmem_ld_wait_dep(r);
#pragma unroll
        for (int i = 0; i < 32; i += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(r[i]), __uint_as_float(r[i + 1])));
      }
      if (c0 < n_keys) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + lane_addr + c0, r);
        tmem_ld_wait_dep(r);
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c0 + i < n_keys) mx = fmaxf(mx, __uint_as_float(r[i]));
      }
    }
    const float mxs = mx * sl2;
    const uint64_t rowkey = attn_rowkey(bh, p.Sq, p.Sk, row_ok ? tid : 0);
#pragma unroll
    for (int t = 0; t < kSmallTailKeys; ++t) {
      if (t < tk) {
        const float e = ex2_approx(fmaf(st[t], sl2, -mxs));
        sum += e;
        st[t] = e * drop_mul(p, bh, row_ok ? tid : 0, 128 + t);
      } else {
        st[t] = 0.0f;
      }
    }
    // pass 2: P (bf16) into the A-operand tiles, 32 columns per TMEM round trip
    auto emit16 = [&](int c0, const uint32_t* r, auto full_tag) {
      constexpr bool kFull = decltype(full_tag)::value;   // full blocks carry no per-key select (two instructions per key otherwise)
      uint32_t packed[8];
      float e[16];
      const int lim = n_keys - c0;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        e[i] = ex2_approx(fmaf(__uint_as_float(r[i]), sl2, -mxs));
        if constexpr (!kFull) e[i] = i < lim ? e[i] : 0.0f;   // padding keys of the last block
      }
#pragma unroll
      for (int i = 0; i < 16; i += 4) sum += (e[i] + e[i + 1]) + (e[i + 2] + e[i + 3]);
      if (p.dropout_thr16) {
        float m[16];
        drop_mul16(p, rowkey, c0, m);
#pragma unroll
        for (int i = 0; i < 16; ++i) e[i] *= m[i];
      }
#pragma unroll
      for (int i = 0; i < 16; i += 2) packed[i >> 1] = pack_bf16x2(e[i], e[i + 1]);
      const uint32_t blk = c0 < 64 ? sQ_s : sK_s;
      const int ch = (c0 & 63) >> 3;
      sts128(blk + sw128(tid, ch), packed[0], packed[1], packed[2], packed[3]);
      sts128(blk + sw128(tid, ch + 1), packed[4], packed[5], packed[6], packed[7]);
    };
    {
      int c0 = 0;
#pragma unroll 1
      for (; c0 + 16 <= n_keys; c0 += 16) {    // one copy of the full-block body: the kernel is instruction-cache sensitive
        uint32_t r[16];
        tmem_ld_32x32b_x16(tmem_base + lane_addr + c0, r);
        tmem_ld_wait_dep(r);
        emit16(c0, r, std::true_type{});
      }
      if (c0 < n_mma) {                         // the partial last block (never at S = 129: 128 main keys)
        uint32_t r[16];
        tmem_ld_32x32b_x16(tmem_base + lane_addr + c0, r);
        tmem_ld_wait_dep(r);
        emit16(c0, r, std::false_type{});
      }
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  if (issuer) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, HD, false, true);
    for (int k = 0; k < n_mma / 16; ++k) {     // O = P V over the main keys (accumulator reuses S's columns 0..63)
      const uint32_t a = smem_u32(k < 4 ? sQ : sK) + (k & 3) * 32;
      tc_mma_f16_ss(tmem_base, make_smem_desc_sw128(a, 16, 1024), make_smem_desc_sw128(smem_u32(sV) + k * 2048, 8192, 1024), idesc, k > 0);
    }
    tc_commit(smem_u32(bar_o));
  }
  if (warp_ok) {
    mbar_wait(smem_u32(bar_o), 0);
    tc_fence_after();
    const float inv = 1.0f / sum;
    // thread == row for the TMEM read; the warp's 32 rows are staged in its 4 KB of the consumed P block 0 and stored with
    // 8 lanes per 128-byte row (a row-per-lane store costs 32 line requests per instruction)
    const uint32_t stage_s = smem_u32(sQ) + warp * 4096;
#pragma unroll 1
    for (int c0 = 0; c0 < HD; c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_base + lane_addr + c0, r);
      tmem_ld_wait_dep(r);
      {
        float o[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(r[i]);
#pragma unroll
        for (int t = 0; t < kSmallTailKeys; ++t) {
          if (t < tk) {          // tail keys: O_i += p_it * v_t
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float vf[8];
              Vec16<__nv_bfloat16>::unpack(lds128(sV_s + sw128(128 + t, (c0 >> 3) + g)), vf);
#pragma unroll
              for (int i = 0; i < 8; ++i) o[8 * g + i] += st[t] * vf[i];
            }
          }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t pk[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) pk[i] = pack_bf16x2(o[8 * g + 2 * i] * inv, o[8 * g + 2 * i + 1] * inv);
          sts128(stage_s + sw128(lane, (c0 >> 3) + g), pk[0], pk[1], pk[2], pk[3]);
        }
      }
    }
    __syncwarp();
    {
      __nv_bfloat16* obase = p.o + (static_cast<long long>(b) * p.Sq + warp * 32) * p.ldo + h * HD;
      const int nvalid = n_rows - warp * 32;
#pragma unroll 1
      for (int it = 0; it < 8; ++it) {
        const int rr = 4 * it + (lane >> 3);
        if (rr < nvalid) *reinterpret_cast<uint4*>(obase + rr * p.ldo + 8 * (lane & 7)) = lds128(stage_s + sw128(rr, lane & 7));
      }
    }
    if (row_ok && p.lse) p.lse[static_cast<long long>(bh) * p.Sq + tid] = mx * p.scale + __logf(sum);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 128);
}

// ------------------------------------------------------------------------------------------- backward
// Outer loop over key chunks of 128, inner loop over query tiles of 128 (<= 2).
// TMEM columns: S [0,128) | dP [128,256) | dK [256,320) | dV [320,384) | dQ tile t at [384 + 64 t, +64)
// smem: Q tiles (m_tiles x 16 KB) | dO tiles (m_tiles x 16 KB) | K chunk 16 KB | V chunk 16 KB | P 2x16 KB | dS 2x16 KB
__global__ void __launch_bounds__(kThreads) bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                       const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                                                       const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int m_tiles = (p.Sq + 127) / 128;
  uint8_t* sQ = smem;
  uint8_t* sdO = sQ + m_tiles * 16384;
  uint8_t* sK = sdO + m_tiles * 16384;
  uint8_t* sV = sK + 16384;
  uint8_t* sP = sV + 16384;       // [128 queries x 128 keys] as two 64-key blocks
  uint8_t* sdS = sP + 32768;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + 32768);
  uint64_t* bar_q = bars;         // all Q and dO tiles landed
  uint64_t* bar_kv = bars + 1;    // K/V chunk landed (one phase per chunk)
  uint64_t* bar_s = bars + 2;     // S and dP complete       (one phase per (chunk, tile))
  uint64_t* bar_g = bars + 3;     // dV, dK, dQ MMAs complete (one phase per (chunk, tile))
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  float* sD = reinterpret_cast<float*>(bars + 6);          // [m_tiles * 128]
  float* sL = sD + m_tiles * 128;                          // lse * log2e

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool issuer = warp == 4 && lane == 0;
  const int bh = blockIdx.x, b = bh / p.H, h = bh % p.H;
  if (issuer) {
    mbar_init(smem_u32(bar_q), 1);
    mbar_init(smem_u32(bar_kv), 1);
    mbar_init(smem_u32(bar_s), 1);
    mbar_init(smem_u32(bar_g), 1);
    fence_mbar_init();
  }
  if (warp == 4) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tdP = tmem_base + 128, tdK = tmem_base + 256, tdV = tmem_base + 320, tdQ = tmem_base + 384;
  const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;

  if (issuer) {
    mbar_arrive_expect_tx(smem_u32(bar_q), 2 * m_tiles * 16384);
    for (int r = 0; r < m_tiles * 128; r += 16) {
      tma_load_2d(smem_u32(sQ + r * 128), &tmQ, smem_u32(bar_q), h * HD, b * p.Sq + r);
      tma_load_2d(smem_u32(sdO + r * 128), &tmdO, smem_u32(bar_q), h * HD, b * p.Sq + r);
    }
  }
  // D_i = sum_c dO_ic O_ic and lse_i, one query row per thread, straight from global memory
  if (warp < 4) {
    for (int mt = 0; mt < m_tiles; ++mt) {
      const int row = mt * 128 + tid;
      float acc = 0.0f, l = 0.0f;
      if (row < p.Sq) {
        const __nv_bfloat16* orow = p.o_in + (static_cast<long long>(b) * p.Sq + row) * p.ldo + h * HD;
        const __nv_bfloat16* drow = p.do_in + (static_cast<long long>(b) * p.Sq + row) * p.lddo + h * HD;
#pragma unroll
        for (int c = 0; c < HD; c += 8) {
          float a[8], d[8];
          Vec16<__nv_bfloat16>::load(orow + c, a);
          Vec16<__nv_bfloat16>::load(drow + c, d);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc += a[i] * d[i];
        }
        l = p.lse[static_cast<long long>(bh) * p.Sq + row] * kLog2e;
      }
      sD[row] = acc;
      sL[row] = l;
    }
  }
  __syncthreads();

  const float sl2 = p.scale * kLog2e;
  uint32_t kv_phase = 0, it_phase = 0;
  const int n_chunks = (p.Sk + 127) / 128;
  for (int kc = 0; kc < n_chunks; ++kc, kv_phase ^= 1) {
    const int k0 = kc * 128;
    const int nk = p.sk_pad - k0 < 128 ? p.sk_pad - k0 : 128;   // padded keys in this chunk (multiple of 16)
    if (issuer) {
      mbar_arrive_expect_tx(smem_u32(bar_kv), 2 * nk * 128);
      for (int r = 0; r < nk; r += 16) {
        tma_load_2d(smem_u32(sK + r * 128), &tmK, smem_u32(bar_kv), h * HD, b * p.Sk + k0 + r);
        tma_load_2d(smem_u32(sV + r * 128), &tmV, smem_u32(bar_kv), h * HD, b * p.Sk + k0 + r);
      }
    }
    for (int mt = 0; mt < m_tiles; ++mt, it_phase ^= 1) {
      if (issuer) {
        if (kc == 0 && mt == 0) mbar_wait(smem_u32(bar_q), 0);
        if (mt == 0) mbar_wait(smem_u32(bar_kv), kv_phase);
        tc_fence_after();
        mma_kk(tS, smem_u32(sQ + mt * 16384), smem_u32(sK), nk);     // S  = Q K^T
        mma_kk(tdP, smem_u32(sdO + mt * 16384), smem_u32(sV), nk);   // dP = dO V^T
        tc_commit(smem_u32(bar_s));
      }
      const int row = mt * 128 + tid;
      const bool row_ok = warp < 4 && row < p.Sq;
      const bool warp_ok = warp < 4 && mt * 128 + warp * 32 < p.Sq;
      if (warp < 4) {
        mbar_wait(smem_u32(bar_s), it_phase);
        tc_fence_after();
        if (warp_ok) {
          const float Di = sD[row], Li = sL[row];
          for (int c0 = 0; c0 < nk; c0 += 16) {
            uint32_t rs[16], rp[16];
            tmem_ld_32x32b_x16(tS + lane_addr + c0, rs);
            tmem_ld_32x32b_x16(tdP + lane_addr + c0, rp);
            tmem_ld_wait_dep(rs, rp);
            uint32_t pp[8], pd[8];
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              float pr[2], ds[2];
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                const int key = k0 + c0 + i + u;
                float prob = 0.0f, dsv = 0.0f;
                if (row_ok && key < p.Sk) {
                  const float m = drop_mul(p, bh, row, key);
                  prob = exp2f(__uint_as_float(rs[i + u]) * sl2 - Li);
                  dsv = prob * (__uint_as_float(rp[i + u]) * m - Di) * p.scale;
                  prob *= m;
                }
                pr[u] = prob;
                ds[u] = dsv;
              }
              pp[i >> 1] = pack_bf16x2(pr[0], pr[1]);
              pd[i >> 1] = pack_bf16x2(ds[0], ds[1]);
            }
            const int blk = (c0 >> 6) * 16384, ch = (c0 & 63) >> 3;
            *reinterpret_cast<uint4*>(sP + blk + sw128(tid, ch)) = make_uint4(pp[0], pp[1], pp[2], pp[3]);
            *reinterpret_cast<uint4*>(sP + blk + sw128(tid, ch + 1)) = make_uint4(pp[4], pp[5], pp[6], pp[7]);
            *reinterpret_cast<uint4*>(sdS + blk + sw128(tid, ch)) = make_uint4(pd[0], pd[1], pd[2], pd[3]);
            *reinterpret_cast<uint4*>(sdS + blk + sw128(tid, ch + 1)) = make_uint4(pd[4], pd[5], pd[6], pd[7]);
          }
        } else {
          // query rows beyond the sequence feed the contraction of dV / dK: they must be zero
          for (int c0 = 0; c0 < nk; c0 += 16) {
            const int blk = (c0 >> 6) * 16384, ch = (c0 & 63) >> 3;
            *reinterpret_cast<uint4*>(sP + blk + sw128(tid, ch)) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(sP + blk + sw128(tid, ch + 1)) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(sdS + blk + sw128(tid, ch)) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(sdS + blk + sw128(tid, ch + 1)) = make_uint4(0, 0, 0, 0);
          }
        }
        // (key columns [nk, 128) stay stale: column j of P / dS only feeds row j of dV / dK, never stored)
        fence_proxy_async_smem();
      }
      tc_fence_before();
      __syncthreads();
      if (issuer) {
        tc_fence_after();
        // dV += P^T dO, dK += dS^T Q : M = 128 keys (two 64-key column blocks, LBO 16 KB), contraction over the 128 queries
        const uint32_t idesc_mn = make_idesc_bf16(128, HD, true, true);
        for (int k = 0; k < 8; ++k) {
          const uint64_t bd = make_smem_desc_sw128(smem_u32(sdO + mt * 16384) + k * 2048, 8192, 1024);
          const uint64_t bq = make_smem_desc_sw128(smem_u32(sQ + mt * 16384) + k * 2048, 8192, 1024);
          tc_mma_f16_ss(tdV, make_smem_desc_sw128(smem_u32(sP) + k * 2048, 16384, 1024), bd, idesc_mn, (mt > 0 || k > 0));
          tc_mma_f16_ss(tdK, make_smem_desc_sw128(smem_u32(sdS) + k * 2048, 16384, 1024), bq, idesc_mn, (mt > 0 || k > 0));
        }
        // dQ_tile += dS K : contraction over the nk keys of this chunk
        const uint32_t idesc_q = make_idesc_bf16(128, HD, false, true);
        for (int k = 0; k < nk / 16; ++k) {
          const uint32_t a = smem_u32(sdS + (k >> 2) * 16384) + (k & 3) * 32;
          tc_mma_f16_ss(tdQ + mt * HD, make_smem_desc_sw128(a, 16, 1024), make_smem_desc_sw128(smem_u32(sK) + k * 2048, 8192, 1024), idesc_q,
                        (kc > 0 || k > 0));
        }
        tc_commit(smem_u32(bar_g));
      }
      if (warp < 4) {
        mbar_wait(smem_u32(bar_g), it_phase);   // P / dS (smem) and S / dP (TMEM) may be overwritten after this
        tc_fence_after();
      }
    }
    // dK / dV rows of this key chunk -> global (thread == key row)
    const int key = k0 + tid;
    if (warp < 4 && k0 + warp * 32 < p.Sk) {
      __nv_bfloat16* dkrow = p.dk + (static_cast<long long>(b) * p.Sk + (key < p.Sk ? key : 0)) * p.lddk + h * HD;
      __nv_bfloat16* dvrow = p.dv + (static_cast<long long>(b) * p.Sk + (key < p.Sk ? key : 0)) * p.lddv + h * HD;
#pragma unroll
      for (int c0 = 0; c0 < HD; c0 += 16) {
        uint32_t rk[16], rv[16];
        tmem_ld_32x32b_x16(tdK + lane_addr + c0, rk);
        tmem_ld_32x32b_x16(tdV + lane_addr + c0, rv);
        tmem_ld_wait_dep(rk, rv);
        if (key < p.Sk) {
          uint32_t a[8], c[8];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            a[i >> 1] = pack_bf16x2(__uint_as_float(rk[i]), __uint_as_float(rk[i + 1]));
            c[i >> 1] = pack_bf16x2(__uint_as_float(rv[i]), __uint_as_float(rv[i + 1]));
          }
          *reinterpret_cast<uint4*>(dkrow + c0) = make_uint4(a[0], a[1], a[2], a[3]);
          *reinterpret_cast<uint4*>(dkrow + c0 + 8) = make_uint4(a[4], a[5], a[6], a[7]);
          *reinterpret_cast<uint4*>(dvrow + c0) = make_uint4(c[0], c[1], c[2], c[3]);
          *reinterpret_cast<uint4*>(dvrow + c0 + 8) = make_uint4(c[4], c[5], c[6], c[7]);
        }
      }
    }
    tc_fence_before();
    __syncthreads();   // K / V chunk buffers and the dK / dV accumulators are reused by the next chunk
    tc_fence_after();
  }
  // dQ tiles -> global
  if (warp < 4) {
    for (int mt = 0; mt < m_tiles; ++mt) {
      const int row = mt * 128 + tid;
      if (mt * 128 + warp * 32 < p.Sq) {
        __nv_bfloat16* dqrow = p.dq + (static_cast<long long>(b) * p.Sq + (row < p.Sq ? row : 0)) * p.lddq + h * HD;
#pragma unroll
        for (int c0 = 0; c0 < HD; c0 += 16) {
          uint32_t r[16];
          tmem_ld_32x32b_x16(tdQ + mt * HD + lane_addr + c0, r);
          tmem_ld_wait_dep(r);
          if (row < p.Sq) {
            uint32_t a[8];
#pragma unroll
            for (int i = 0; i < 16; i += 2) a[i >> 1] = pack_bf16x2(__uint_as_float(r[i]), __uint_as_float(r[i + 1]));
            *reinterpret_cast<uint4*>(dqrow + c0) = make_uint4(a[0], a[1], a[2], a[3]);
            *reinterpret_cast<uint4*>(dqrow + c0 + 8) = make_uint4(a[4], a[5], a[6], a[7]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------- backward, v2
// Sq <= 128 + kMaxTail: ONE 128-row tensor-core tile + tail query rows on CUDA cores (warp 5); keys in chunks
// of <= 144 (the 129-token sequences are one chunk).  Per chunk:
//   issuer   S = Q K^T, dP = dO V^T                         (N = nk <= 144)
//   workers  P, dS -> bf16 swizzled smem tiles (3 k-blocks)  | tail warp: p, ds of the tail rows (smem vectors)
//   issuer   dV = P^T dO, dK = dS^T Q (keys 0..127 of the chunk), dQ += dS K
//   workers  drain dK / dV rows, add the tail rows' rank-1 terms, store
//   if nk > 128: issuer recomputes dV / dK for keys 128..nk-1 from the third k-block (same TMEM columns)
// TMEM: S [0,144) | dP [160,304) | dK [320,384) | dV [384,448) | dQ [448,512)
// smem: Q 16K | dO 16K | P 48K | dS 48K | K 18K | V 18K | barriers | D, lse | tail scratch
constexpr int kThreadsBwd2 = 320;   // 8 worker warps (two threads per row), issue warp, tail warp

__global__ void __launch_bounds__(kThreadsBwd2) bwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                            const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                                                            const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sdO = sQ + 16384;
  uint8_t* sP = sdO + 16384;        // 3 k-blocks
  uint8_t* sdS = sP + 3 * 16384;    // 3 k-blocks; block 2 + 16 KB falls into sK (valid, finite data)
  uint8_t* sK = sdS + 3 * 16384;
  uint8_t* sV = sK + KC * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + KC * 128);
  uint64_t* bar_q = bars;           // Q and dO tiles landed
  uint64_t* bar_kv = bars + 1;      // K / V chunk landed          (one phase per chunk)
  uint64_t* bar_s = bars + 2;       // S and dP complete           (one phase per chunk)
  uint64_t* bar_g = bars + 3;       // dV, dK, dQ MMAs complete    (one phase per chunk)
  uint64_t* bar_t = bars + 4;       // tail-key dV, dK complete    (one phase per chunk that has tail keys)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  float* sDh = reinterpret_cast<float*>(bars + 8);  // [2][128] per-half partial D
  float* sL = sDh + 256;                            // [128] lse * log2e
  float* tq = sL + 128;                             // [kMaxTail][64] tail query rows
  float* tdo = tq + kMaxTail * HD;                  // [kMaxTail][64] tail dO rows
  float* tdq = tdo + kMaxTail * HD;                 // [kMaxTail][64] tail dQ accumulators
  float* tp = tdq + kMaxTail * HD;                  // [kMaxTail][KC] P * mask of the tail rows, this chunk
  float* tds = tp + kMaxTail * KC;                  // [kMaxTail][KC] dS of the tail rows, this chunk
  float* tDL = tds + kMaxTail * KC;                 // [kMaxTail][2]  D_t, lse_t * log2e

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool worker = warp < 8;
  const int row = tid & 127, half = (tid >> 7) & 1;   // two worker threads per row: they split the columns
  const bool issuer = warp == 8 && lane == 0;
  const int bh = blockIdx.x, b = bh / p.H, h = bh % p.H;
  const int n_main = p.Sq < 128 ? p.Sq : 128;       // valid rows of the tensor-core tile
  const int ntail = p.Sq - n_main;                  // rows handled by warp 5
  if (issuer) {
    mbar_init(smem_u32(bar_q), 1);
    mbar_init(smem_u32(bar_kv), 1);
    mbar_init(smem_u32(bar_s), 1);
    mbar_init(smem_u32(bar_g), 1);
    mbar_init(smem_u32(bar_t), 1);
    fence_mbar_init();
  }
  if (warp == 8) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tdP = tmem_base + 160, tdK = tmem_base + 320, tdV = tmem_base + 384, tdQ = tmem_base + 448;
  const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const float sl2 = p.scale * kLog2e;

  if (issuer) {
    mbar_arrive_expect_tx(smem_u32(bar_q), 2 * 16384);
    tma_load_2d(smem_u32(sQ), &tmQ, smem_u32(bar_q), h * HD, b * p.Sq);
    tma_load_2d(smem_u32(sdO), &tmdO, smem_u32(bar_q), h * HD, b * p.Sq);
  }
  if (worker) {
    // D_i = sum_c dO_ic O_ic (each half sums 32 of the 64 columns) and lse_i for the main rows, from global memory
    float acc = 0.0f, l = 0.0f;
    if (row < n_main) {
      const __nv_bfloat16* orow = p.o_in + (static_cast<long long>(b) * p.Sq + row) * p.ldo + h * HD + 32 * half;
      const __nv_bfloat16* drow = p.do_in + (static_cast<long long>(b) * p.Sq + row) * p.lddo + h * HD + 32 * half;
#pragma unroll
      for (int c = 0; c < 32; c += 8) {
        float a[8], d[8];
        Vec16<__nv_bfloat16>::load(orow + c, a);
        Vec16<__nv_bfloat16>::load(drow + c, d);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += a[i] * d[i];
      }
      l = p.lse[static_cast<long long>(bh) * p.Sq + row] * kLog2e;
    }
    sDh[half * 128 + row] = acc;
    sL[row] = l;
  } else if (warp == 9) {
    for (int t = 0; t < ntail; ++t) {
      const long long grow = static_cast<long long>(b) * p.Sq + n_main + t;
      const float2 q2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p.q_in + grow * p.ldq + h * HD + 2 * lane));
      const float2 d2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p.do_in + grow * p.lddo + h * HD + 2 * lane));
      const float2 o2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p.o_in + grow * p.ldo + h * HD + 2 * lane));
      tq[t * HD + 2 * lane] = q2.x; tq[t * HD + 2 * lane + 1] = q2.y;
      tdo[t * HD + 2 * lane] = d2.x; tdo[t * HD + 2 * lane + 1] = d2.y;
      tdq[t * HD + 2 * lane] = 0.0f; tdq[t * HD + 2 * lane + 1] = 0.0f;
      const float Dt = warp_sum(d2.x * o2.x + d2.y * o2.y);
      if (lane == 0) {
        tDL[2 * t] = Dt;
        tDL[2 * t + 1] = p.lse[static_cast<long long>(bh) * p.Sq + n_main + t] * kLog2e;
      }
    }
  }
  __syncthreads();

  uint32_t phase = 0, t_phase = 0;
  for (int k0 = 0; k0 < p.Sk; k0 += KC, phase ^= 1) {
    const int nk = p.sk_pad - k0 < KC ? p.sk_pad - k0 : KC;   // padded keys in this chunk (multiple of 16)
    if (issuer) {
      mbar_arrive_expect_tx(smem_u32(bar_kv), 2 * nk * 128);
      for (int r = 0; r < nk; r += 16) {
        tma_load_2d(smem_u32(sK + r * 128), &tmK, smem_u32(bar_kv), h * HD, b * p.Sk + k0 + r);
        tma_load_2d(smem_u32(sV + r * 128), &tmV, smem_u32(bar_kv), h * HD, b * p.Sk + k0 + r);
      }
      if (k0 == 0) mbar_wait(smem_u32(bar_q), 0);
      mbar_wait(smem_u32(bar_kv), phase);
      tc_fence_after();
      mma_kk(tS, smem_u32(sQ), smem_u32(sK), nk);      // S  = Q K^T
      mma_kk(tdP, smem_u32(sdO), smem_u32(sV), nk);    // dP = dO V^T
      tc_commit(smem_u32(bar_s));
    }
    if (worker) {
      mbar_wait(smem_u32(bar_s), phase);
      tc_fence_after();
      const bool row_ok = row < n_main;
      if ((warp & 3) * 32 < n_main) {
        const float Di = sDh[row] + sDh[128 + row], Li = sL[row];
        const uint64_t rowkey = attn_rowkey(bh, p.Sq, p.Sk, row_ok ? row : 0);
        // 16 columns of P / dS from raw S / dP accumulators -> swizzled bf16 tiles
        auto emit16 = [&](int c0, const uint32_t* rs, const uint32_t* rp) {
          float m[16];
          if (p.dropout_thr16) {
            drop_mul16(p, rowkey, k0 + c0, m);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) m[i] = 1.0f;
          }
          uint32_t pp[8], pd[8];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            float pr[2], ds[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              float prob = 0.0f, dsv = 0.0f;
              if (row_ok && k0 + c0 + i + u < p.Sk) {
                prob = exp2f(__uint_as_float(rs[i + u]) * sl2 - Li);
                dsv = prob * (__uint_as_float(rp[i + u]) * m[i + u] - Di) * p.scale;
                prob *= m[i + u];
              }
              pr[u] = prob;
              ds[u] = dsv;
            }
            pp[i >> 1] = pack_bf16x2(pr[0], pr[1]);
            pd[i >> 1] = pack_bf16x2(ds[0], ds[1]);
          }
          const uint32_t blk = (c0 >> 6) * 16384;
          const int ch = (c0 & 63) >> 3;
          sts128(smem_u32(sP) + blk + sw128(row, ch), pp[0], pp[1], pp[2], pp[3]);
          sts128(smem_u32(sP) + blk + sw128(row, ch + 1), pp[4], pp[5], pp[6], pp[7]);
          sts128(smem_u32(sdS) + blk + sw128(row, ch), pd[0], pd[1], pd[2], pd[3]);
          sts128(smem_u32(sdS) + blk + sw128(row, ch + 1), pd[4], pd[5], pd[6], pd[7]);
        };
#pragma unroll 1
        for (int c0 = 16 * half; c0 < nk; c0 += 32) {   // the two halves interleave 16-column chunks (one copy of the body)
          uint32_t rs[16], rp[16];
          tmem_ld_32x32b_x16(tS + lane_addr + c0, rs);
          tmem_ld_32x32b_x16(tdP + lane_addr + c0, rp);
          tmem_ld_wait_dep(rs, rp);
          emit16(c0, rs, rp);
        }
      } else {
        // query rows beyond the sequence feed the contraction of dV / dK: they must be zero
        for (int c0 = 8 * half; c0 < nk; c0 += 16) {
          const uint32_t off = (c0 >> 6) * 16384 + sw128(row, (c0 & 63) >> 3);
          sts128(smem_u32(sP) + off, 0, 0, 0, 0);
          sts128(smem_u32(sdS) + off, 0, 0, 0, 0);
        }
      }
      fence_proxy_async_smem();
    } else if (warp == 9 && ntail > 0) {
      // tail query rows against this key chunk (CUDA cores), K / V read from the swizzled smem tiles
      mbar_wait(smem_u32(bar_kv), phase);
      for (int t = 0; t < ntail; ++t) {
        const float Dt = tDL[2 * t], Lt = tDL[2 * t + 1];
        const uint64_t rowkey = attn_rowkey(bh, p.Sq, p.Sk, n_main + t);
        for (int j = lane; j < nk; j += 32) {
          float pm = 0.0f, dsv = 0.0f;
          if (k0 + j < p.Sk) {
            float sacc = 0.0f, dacc = 0.0f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              float kf[8], vf[8];
              Vec16<__nv_bfloat16>::unpack(*reinterpret_cast<const uint4*>(sK + sw128(j, c)), kf);
              Vec16<__nv_bfloat16>::unpack(*reinterpret_cast<const uint4*>(sV + sw128(j, c)), vf);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                sacc += tq[t * HD + c * 8 + i] * kf[i];
                dacc += tdo[t * HD + c * 8 + i] * vf[i];
              }
            }
            float m = 1.0f;
            if (p.dropout_thr16) {
              const uint64_t bits = attn_drop_bits(mix_seed(p.dropout_seed, p.seed_src), rowkey, (k0 + j) >> 2);
              m = dropout_keep_lane(bits, (k0 + j) & 3, p.dropout_thr16) ? p.dropout_scale : 0.0f;
            }
            const float prob = exp2f(sacc * sl2 - Lt);
            dsv = prob * (dacc * m - Dt) * p.scale;
            pm = prob * m;
          }
          tp[t * KC + j] = pm;
          tds[t * KC + j] = dsv;
        }
        __syncwarp();
        float a0 = tdq[t * HD + 2 * lane], a1 = tdq[t * HD + 2 * lane + 1];
        for (int j = 0; j < nk; ++j) {
          const float dsj = tds[t * KC + j];
          const float2 k2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sw_elem(sK, j, 2 * lane)));
          a0 += dsj * k2.x;
          a1 += dsj * k2.y;
        }
        tdq[t * HD + 2 * lane] = a0;
        tdq[t * HD + 2 * lane + 1] = a1;
      }
    }
    tc_fence_before();
    __syncthreads();
    if (issuer) {
      tc_fence_after();
      // dV = P^T dO, dK = dS^T Q for keys 0..127 of the chunk: A MN-major (two 64-key blocks, LBO 16 KB), contraction over queries
      const uint32_t idesc_mn = make_idesc_bf16(128, HD, true, true);
      for (int k = 0; k < 8; ++k) {
        const uint64_t bd = make_smem_desc_sw128(smem_u32(sdO) + k * 2048, 8192, 1024);
        const uint64_t bq = make_smem_desc_sw128(smem_u32(sQ) + k * 2048, 8192, 1024);
        tc_mma_f16_ss(tdV, make_smem_desc_sw128(smem_u32(sP) + k * 2048, 16384, 1024), bd, idesc_mn, k > 0);
        tc_mma_f16_ss(tdK, make_smem_desc_sw128(smem_u32(sdS) + k * 2048, 16384, 1024), bq, idesc_mn, k > 0);
      }
      // dQ += dS K: contraction over the nk keys of this chunk
      const uint32_t idesc_q = make_idesc_bf16(128, HD, false, true);
      for (int k = 0; k < nk / 16; ++k) {
        const uint32_t a = smem_u32(sdS + (k >> 2) * 16384) + (k & 3) * 32;
        tc_mma_f16_ss(tdQ, make_smem_desc_sw128(a, 16, 1024), make_smem_desc_sw128(smem_u32(sK) + k * 2048, 8192, 1024), idesc_q,
                      (k0 > 0 || k > 0));
      }
      tc_commit(smem_u32(bar_g));
    }
    // drain dK / dV (thread == key row of the chunk), add the tail query rows' rank-1 terms
    auto drain_keys = [&](int jbase) {        // half 0 drains dK (+ dS_t q_t), half 1 drains dV (+ p_t dO_t)
      const int j = jbase + row;                 // key index inside the chunk
      const int key = k0 + j;
      const bool ok = key < p.Sk;
      __nv_bfloat16* grow = (half ? p.dv : p.dk) + (static_cast<long long>(b) * p.Sk + (ok ? key : 0)) * (half ? p.lddv : p.lddk) + h * HD;
      const float* tw = half ? tp : tds;         // per-key weight of the tail query rows
      const float* tr = half ? tdo : tq;         // their row vectors
      const uint32_t tacc = half ? tdV : tdK;
#pragma unroll 1
      for (int c0 = 0; c0 < HD; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tacc + lane_addr + c0, r);
        tmem_ld_wait_dep(r);
        if (ok) {
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(r[i]);
          for (int t = 0; t < ntail; ++t) {
            const float wj = tw[t * KC + j];
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] += wj * tr[t * HD + c0 + i];
          }
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<uint4*>(grow + c0 + 8 * g) = make_uint4(pack_bf16x2(f[8 * g], f[8 * g + 1]), pack_bf16x2(f[8 * g + 2], f[8 * g + 3]),
                                                                        pack_bf16x2(f[8 * g + 4], f[8 * g + 5]), pack_bf16x2(f[8 * g + 6], f[8 * g + 7]));
        }
      }
    };
    if (worker) {
      mbar_wait(smem_u32(bar_g), phase);
      tc_fence_after();
      if (k0 + (warp & 3) * 32 < p.Sk) drain_keys(0);
    }
    if (nk > 128) {
      tc_fence_before();
      __syncthreads();          // dK / dV accumulators drained: reuse them for keys 128.. of the chunk
      if (issuer) {
        tc_fence_after();
        const uint32_t idesc_mn = make_idesc_bf16(128, HD, true, true);
        for (int k = 0; k < 8; ++k) {
          const uint64_t bd = make_smem_desc_sw128(smem_u32(sdO) + k * 2048, 8192, 1024);
          const uint64_t bq = make_smem_desc_sw128(smem_u32(sQ) + k * 2048, 8192, 1024);
          tc_mma_f16_ss(tdV, make_smem_desc_sw128(smem_u32(sP) + 2 * 16384 + k * 2048, 16384, 1024), bd, idesc_mn, k > 0);
          tc_mma_f16_ss(tdK, make_smem_desc_sw128(smem_u32(sdS) + 2 * 16384 + k * 2048, 16384, 1024), bq, idesc_mn, k > 0);
        }
        tc_commit(smem_u32(bar_t));
      }
      if (worker) {
        mbar_wait(smem_u32(bar_t), t_phase);
        tc_fence_after();
        if ((warp & 3) * 32 < nk - 128 && k0 + 128 + (warp & 3) * 32 < p.Sk) drain_keys(128);
      }
      t_phase ^= 1;
    }
    tc_fence_before();
    __syncthreads();   // K / V / P / dS buffers and the dK / dV accumulators are reused by the next chunk
    tc_fence_after();
  }
  // dQ: main rows from TMEM, tail rows from the CUDA-core accumulators
  if (worker && (warp & 3) * 32 < n_main) {
    __nv_bfloat16* dqrow = p.dq + (static_cast<long long>(b) * p.Sq + (row < n_main ? row : 0)) * p.lddq + h * HD + 32 * half;
    uint32_t r[32];
    tmem_ld_32x32b_x32(tdQ + lane_addr + 32 * half, r);
    tmem_ld_wait_dep(r);
    if (row < n_main) {
#pragma unroll
      for (int g = 0; g < 4; ++g)
        *reinterpret_cast<uint4*>(dqrow + 8 * g) =
            make_uint4(pack_bf16x2(__uint_as_float(r[8 * g]), __uint_as_float(r[8 * g + 1])), pack_bf16x2(__uint_as_float(r[8 * g + 2]), __uint_as_float(r[8 * g + 3])),
                       pack_bf16x2(__uint_as_float(r[8 * g + 4]), __uint_as_float(r[8 * g + 5])), pack_bf16x2(__uint_as_float(r[8 * g + 6]), __uint_as_float(r[8 * g + 7])));
    }
  } else if (warp == 9) {
    for (int t = 0; t < ntail; ++t) {
      __nv_bfloat16* dqrow = p.dq + (static_cast<long long>(b) * p.Sq + n_main + t) * p.lddq + h * HD;
      *reinterpret_cast<__nv_bfloat162*>(dqrow + 2 * lane) = __floats2bfloat162_rn(tdq[t * HD + 2 * lane], tdq[t * HD + 2 * lane + 1]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------- backward, v3
// Single key chunk (Sk <= 144) and Sq <= 128 + kMaxTail: every self-attention of the model family.  Same tiles, TMEM
// columns and tail-row scheme as bwd2_kernel, re-scheduled after a phase-stamp profile of that kernel (tools/attn_stamps.py:
// 31 k clk per CTA, of which D-from-global 4.4 k, K/V TMA issue 2.3 k, P/dS pass 8.2 k, unrolled MMA issue 3.7 k, drains
// 5.7 k, the one-key tail pass 3.9 k):
//   * Q, dO, O, K, V all arrive by TMA issued at t = 0 (one box each; O lands in the not-yet-used P block 0 and
//     D_i = sum_c dO_ic O_ic is computed from shared memory);
//   * five tail-row warps split the keys (warp 8's lane 0 is also the TMA / MMA issuer); the next CTA's operands are
//     prefetched into L2 while this one computes;
//   * lean P / dS arithmetic: ex2.approx, masking only as a bitwise fix-up of the last (partial) block, softmax scale
//     applied when dK / dQ leave instead of per dS element;
//   * dV / dK of the tail keys are issued with the main gradient MMAs into the (by then free) S columns, so there is no
//     second sync / commit / drain round; two commits let the dK / dV drains overlap the dQ and tail-key MMAs;
//   * rolled issue loops and ONE out-of-line drain routine: the kernel is straight-line code executed once per CTA, so
//     its instruction footprint is fetch time.
// TMEM: S [0,144) | dP [160,304) | dK [320,384) | dV [384,448) | dQ [448,512); tail keys: dV_t [0,64) | dK_t [64,128)
// smem: Q 16K | dO 16K | P 48K (block 0 first holds O) | dS 48K | K 18K | V 18K | barriers | D partials | tail scratch
constexpr int kThreadsBwd3 = 416;   // warps 0-7 workers (two threads per row), warps 8-12 tail query rows (warp 12: keys >= 128)

#ifdef TVT_ATTN_STAMPS
__device__ long long g_stamps[3][32];
#define TVT_STAMP(slot)                                                                                                  \
  do {                                                                                                                   \
    const int sb_ = blockIdx.x == 0 ? 0 : (blockIdx.x == gridDim.x / 2 ? 1 : (blockIdx.x == gridDim.x - 1 ? 2 : -1));    \
    if (sb_ >= 0) g_stamps[sb_][slot] = clock64();                                                                       \
  } while (0)
#else
#define TVT_STAMP(slot) do {} while (0)
#endif


// One warp's 32 rows of a [128 x 64] fp32 accumulator -> bf16 global rows: out = (acc + sum_t w_t * vec_t) * mul, the rank-1
// terms being the tail query rows' contributions (w_t = this lane's row weight at w_s + t * KC floats, vec_t = 64-float vector
// at vec_s + t * HD floats).  Thread == row for the TMEM read, but a row-per-lane global store costs 32 line requests per
// instruction (the LSU was the drain's bottleneck), so the warp's rows are staged through 4 KB of swizzled shared memory and
// stored with 8 lanes per 128-byte row.  Rows [0, nvalid) of the warp are written; row r goes to dst + r * ld.
// Deliberately not inlined: one copy serves dK, dV, their tail keys and dQ.
__device__ __noinline__ void drain_rows(uint32_t taddr, __nv_bfloat16* dst, long long ld, int nvalid, float mul, int ntail, uint32_t w_s,
                                        uint32_t vec_s, uint32_t stage_s) {
  const int lane = threadIdx.x & 31;
#pragma unroll 1
  for (int pc = 0; pc < 2; ++pc) {
    uint32_t r[32];
    tmem_ld_32x32b_x32(taddr + 32 * pc, r);
    tmem_ld_wait_dep(r);
    float f[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(r[i]);
#pragma unroll 1
    for (int t = 0; t < ntail; ++t) {
      const float wj = lds_f32(w_s + t * KC * 4);
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const uint4 v = lds128(vec_s + (t * HD + 32 * pc + i) * 4);
        f[i] += wj * __uint_as_float(v.x);
        f[i + 1] += wj * __uint_as_float(v.y);
        f[i + 2] += wj * __uint_as_float(v.z);
        f[i + 3] += wj * __uint_as_float(v.w);
      }
    }
#pragma unroll
    for (int g = 0; g < 4; ++g)
      sts128(stage_s + sw128(lane, 4 * pc + g), pack_bf16x2(f[8 * g] * mul, f[8 * g + 1] * mul), pack_bf16x2(f[8 * g + 2] * mul, f[8 * g + 3] * mul),
             pack_bf16x2(f[8 * g + 4] * mul, f[8 * g + 5] * mul), pack_bf16x2(f[8 * g + 6] * mul, f[8 * g + 7] * mul));
  }
  __syncwarp();
#pragma unroll 1
  for (int it = 0; it < 8; ++it) {
    const int rr = 4 * it + (lane >> 3);
    if (rr < nvalid) *reinterpret_cast<uint4*>(dst + rr * ld + 8 * (lane & 7)) = lds128(stage_s + sw128(rr, lane & 7));
  }
  __syncwarp();
}

__global__ void __launch_bounds__(kThreadsBwd3) bwd3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                            const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                                                            const __grid_constant__ CUtensorMap tmO, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sdO = sQ + 16384;
  uint8_t* sP = sdO + 16384;        // 3 k-blocks; block 0 holds the O tile until D has been computed
  uint8_t* sdS = sP + 3 * 16384;    // 3 k-blocks; the MN-major tail-key operand runs 16 KB past block 2 into sK (finite data, rows never drained)
  uint8_t* sK = sdS + 3 * 16384;
  uint8_t* sV = sK + KC * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + KC * 128);
  uint64_t* bar_ld = bars;          // Q, dO, O, K, V landed
  uint64_t* bar_s = bars + 1;       // S and dP complete
  uint64_t* bar_g1 = bars + 2;      // dV, dK (keys 0..127) complete
  uint64_t* bar_g2 = bars + 3;      // dQ and the tail keys' dV, dK complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  float* sDh = reinterpret_cast<float*>(bars + 8);  // [2][128] per-half partial D
  float* tq = sDh + 256;                            // [kMaxTail][64] tail query rows
  float* tdo = tq + kMaxTail * HD;                  // [kMaxTail][64] tail dO rows
  float* tdq = tdo + kMaxTail * HD;                 // [kMaxTail][64] tail dQ accumulators (shared atomics of the four tail warps)
  float* tp = tdq + kMaxTail * HD;                  // [kMaxTail][KC] P * mask of the tail rows
  float* tds = tp + kMaxTail * KC;                  // [kMaxTail][KC] dS / scale of the tail rows
  float* tDL = tds + kMaxTail * KC;                 // [kMaxTail][2]  D_t, lse_t * log2e

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool worker = warp < 8;
  const int row = tid & 127, half = (tid >> 7) & 1;   // two worker threads per row: they split the columns
  const bool issuer = warp == 8 && lane == 0;
  const int bh = blockIdx.x, b = bh / p.H, h = bh % p.H;
  const int n_main = p.Sq < 128 ? p.Sq : 128;       // valid rows of the tensor-core tile
  const int ntail = p.Sq - n_main;                  // rows handled by the tail warps
  const int nk = p.sk_pad;                          // padded keys (multiple of 16, <= KC)
  pdl_wait();      // the operand TMAs are issued in the prologue, so wait first: only the launch latency is overlapped
  pdl_trigger();
  if (tid == 0) TVT_STAMP(0);
  if (issuer) {
    mbar_init(smem_u32(bar_ld), 1);
    mbar_init(smem_u32(bar_s), 1);
    mbar_init(smem_u32(bar_g1), 1);
    mbar_init(smem_u32(bar_g2), 1);
    fence_mbar_init();
    mbar_arrive_expect_tx(smem_u32(bar_ld), 3 * 16384 + 2 * nk * 128);
    tma_load_2d(smem_u32(sQ), &tmQ, smem_u32(bar_ld), h * HD, b * p.Sq);
    tma_load_2d(smem_u32(sK), &tmK, smem_u32(bar_ld), h * HD, b * p.Sk);
    tma_load_2d(smem_u32(sdO), &tmdO, smem_u32(bar_ld), h * HD, b * p.Sq);
    tma_load_2d(smem_u32(sV), &tmV, smem_u32(bar_ld), h * HD, b * p.Sk);
    tma_load_2d(smem_u32(sP), &tmO, smem_u32(bar_ld), h * HD, b * p.Sq);
  }
  if (warp == 8) {
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  float Li = 0.0f;
  if (worker && row < n_main) Li = p.lse[static_cast<long long>(bh) * p.Sq + row] * kLog2e;
  tc_fence_before();
  __syncthreads();                  // #1: TMEM base address visible
  tc_fence_after();
  if (tid == 0) TVT_STAMP(1);
  if (warp == 10 && lane == 0) {
    // the CTA that follows this one on the SM (one resident CTA per SM, so roughly blockIdx + #SMs): pull its operands into L2
    const int nb = bh + p.prefetch_stride;
    if (nb < static_cast<int>(gridDim.x)) {
      const int b2 = nb / p.H, h2 = nb % p.H;
      tma_prefetch_2d(&tmQ, h2 * HD, b2 * p.Sq);
      tma_prefetch_2d(&tmK, h2 * HD, b2 * p.Sk);
      tma_prefetch_2d(&tmdO, h2 * HD, b2 * p.Sq);
      tma_prefetch_2d(&tmV, h2 * HD, b2 * p.Sk);
      tma_prefetch_2d(&tmO, h2 * HD, b2 * p.Sq);
    }
  }
  if (warp == 9) {
    for (int t = 0; t < ntail; ++t) {
      const long long grow = static_cast<long long>(b) * p.Sq + n_main + t;
      const float2 q2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p.q_in + grow * p.ldq + h * HD + 2 * lane));
      const float2 d2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p.do_in + grow * p.lddo + h * HD + 2 * lane));
      const float2 o2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p.o_in + grow * p.ldo + h * HD + 2 * lane));
      tq[t * HD + 2 * lane] = q2.x; tq[t * HD + 2 * lane + 1] = q2.y;
      tdo[t * HD + 2 * lane] = d2.x; tdo[t * HD + 2 * lane + 1] = d2.y;
      tdq[t * HD + 2 * lane] = 0.0f; tdq[t * HD + 2 * lane + 1] = 0.0f;
      const float Dt = warp_sum(d2.x * o2.x + d2.y * o2.y);
      if (lane == 0) {
        tDL[2 * t] = Dt;
        tDL[2 * t + 1] = p.lse[static_cast<long long>(bh) * p.Sq + n_main + t] * kLog2e;
      }
    }
  }
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tdP = tmem_base + 160, tdK = tmem_base + 320, tdV = tmem_base + 384, tdQ = tmem_base + 448;
  const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const float sl2 = p.scale * kLog2e;

  if (warp == 8) {
    if (lane == 0) {
      mbar_wait(smem_u32(bar_ld), 0);
      TVT_STAMP(16);
      tc_fence_after();
      const uint32_t idesc = make_idesc_bf16(128, nk, false, false);
#pragma unroll 1
      for (int k = 0; k < HD / 16; ++k) {
        tc_mma_f16_ss(tS, make_smem_desc_sw128(smem_u32(sQ) + k * 32, 16, 1024), make_smem_desc_sw128(smem_u32(sK) + k * 32, 16, 1024), idesc, k > 0);
        tc_mma_f16_ss(tdP, make_smem_desc_sw128(smem_u32(sdO) + k * 32, 16, 1024), make_smem_desc_sw128(smem_u32(sV) + k * 32, 16, 1024), idesc, k > 0);
      }
      tc_commit(smem_u32(bar_s));
      TVT_STAMP(17);
    }
    __syncwarp();
  }
  if (worker) {
    // D_i = sum_c dO_ic O_ic from the shared-memory tiles (each half sums 32 of the 64 columns)
    mbar_wait(smem_u32(bar_ld), 0);
    float acc = 0.0f;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      float a[8], d[8];
      Vec16<__nv_bfloat16>::unpack(lds128(smem_u32(sP) + sw128(row, 4 * half + ch)), a);
      Vec16<__nv_bfloat16>::unpack(lds128(smem_u32(sdO) + sw128(row, 4 * half + ch)), d);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc += a[i] * d[i];
    }
    sDh[half * 128 + row] = acc;
  }
  __syncthreads();                  // #2: D partials and tail scratch visible; O has been read, P block 0 may be overwritten
  if (tid == 0) TVT_STAMP(2);

  if (worker) {
    mbar_wait(smem_u32(bar_s), 0);
    tc_fence_after();
    if (tid == 0) TVT_STAMP(3);
    const bool row_ok = row < n_main;
    const int lim = row_ok ? p.Sk : 0;          // valid keys of this row: query rows beyond the sequence must come out as zeros
    const float Di = sDh[row] + sDh[128 + row];
    const uint64_t rowkey = attn_rowkey(bh, p.Sq, p.Sk, row_ok ? row : 0);
#pragma unroll 1
    for (int c0 = 16 * half; c0 < nk; c0 += 32) {   // the two halves interleave 16-column blocks
      uint32_t rs[16], rp[16];
      tmem_ld_32x32b_x16(tS + lane_addr + c0, rs);
      tmem_ld_32x32b_x16(tdP + lane_addr + c0, rp);
      float m[16];
      if (p.dropout_thr16) {
        drop_mul16(p, rowkey, c0, m);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) m[i] = 1.0f;
      }
      tmem_ld_wait_dep(rs, rp);
      uint32_t pp[8], pd[8];
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        const float p0 = ex2_approx(fmaf(__uint_as_float(rs[i]), sl2, -Li));
        const float p1 = ex2_approx(fmaf(__uint_as_float(rs[i + 1]), sl2, -Li));
        const float d0 = p0 * fmaf(__uint_as_float(rp[i]), m[i], -Di);
        const float d1 = p1 * fmaf(__uint_as_float(rp[i + 1]), m[i + 1], -Di);
        pp[i >> 1] = pack_bf16x2(p0 * m[i], p1 * m[i + 1]);
        pd[i >> 1] = pack_bf16x2(d0, d1);
      }
      if (c0 + 16 > lim) {          // partial block (padding keys hold other rows' data) or a row beyond the sequence
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          const int c = c0 + 2 * w;
          const uint32_t keep = c + 1 < lim ? 0xFFFFFFFFu : (c < lim ? 0x0000FFFFu : 0u);
          pp[w] &= keep;
          pd[w] &= keep;
        }
      }
      const uint32_t blk = (c0 >> 6) * 16384;
      const int ch = (c0 & 63) >> 3;
      sts128(smem_u32(sP) + blk + sw128(row, ch), pp[0], pp[1], pp[2], pp[3]);
      sts128(smem_u32(sP) + blk + sw128(row, ch + 1), pp[4], pp[5], pp[6], pp[7]);
      sts128(smem_u32(sdS) + blk + sw128(row, ch), pd[0], pd[1], pd[2], pd[3]);
      sts128(smem_u32(sdS) + blk + sw128(row, ch + 1), pd[4], pd[5], pd[6], pd[7]);
    }
    fence_proxy_async_smem();
    if (tid == 0) TVT_STAMP(4);
  } else if (ntail > 0) {
    // tail query rows on CUDA cores: tail warp tw takes keys [32 tw, 32 tw + 32) (warp 12: the tail keys 128..)
    const int tw = warp - 8;
    mbar_wait(smem_u32(bar_ld), 0);
    if (lane == 0 && tw < 4) TVT_STAMP(24 + tw);
    for (int t = 0; t < ntail; ++t) {
      const float Dt = tDL[2 * t], Lt = tDL[2 * t + 1];
      const uint64_t rowkey = attn_rowkey(bh, p.Sq, p.Sk, n_main + t);
      float a0 = 0.0f, a1 = 0.0f;
      for (int jb = 32 * tw; jb < nk; jb += 160) {
        const int j = jb + lane;
        if (j < nk) {
          float pm = 0.0f, dsv = 0.0f;
          if (j < p.Sk) {
            float sacc = 0.0f, dacc = 0.0f;
#pragma unroll 1
            for (int c = 0; c < 8; ++c) {
              float kf[8], vf[8];
              Vec16<__nv_bfloat16>::unpack(lds128(smem_u32(sK) + sw128(j, c)), kf);
              Vec16<__nv_bfloat16>::unpack(lds128(smem_u32(sV) + sw128(j, c)), vf);
#pragma unroll
              for (int i = 0; i < 8; i += 4) {
                const uint4 q4 = lds128(smem_u32(tq) + (t * HD + c * 8 + i) * 4);
                const uint4 o4 = lds128(smem_u32(tdo) + (t * HD + c * 8 + i) * 4);
                sacc += __uint_as_float(q4.x) * kf[i] + __uint_as_float(q4.y) * kf[i + 1] + __uint_as_float(q4.z) * kf[i + 2] + __uint_as_float(q4.w) * kf[i + 3];
                dacc += __uint_as_float(o4.x) * vf[i] + __uint_as_float(o4.y) * vf[i + 1] + __uint_as_float(o4.z) * vf[i + 2] + __uint_as_float(o4.w) * vf[i + 3];
              }
            }
            float m = 1.0f;
            if (p.dropout_thr16) {
              const uint64_t bits = attn_drop_bits(mix_seed(p.dropout_seed, p.seed_src), rowkey, j >> 2);
              m = dropout_keep_lane(bits, j & 3, p.dropout_thr16) ? p.dropout_scale : 0.0f;
            }
            const float prob = ex2_approx(fmaf(sacc, sl2, -Lt));
            dsv = prob * (dacc * m - Dt);
            pm = prob * m;
          }
          tp[t * KC + j] = pm;
          tds[t * KC + j] = dsv;
        }
        __syncwarp();
        const int jn = nk - jb < 32 ? nk - jb : 32;
#pragma unroll 4
        for (int jj = 0; jj < jn; ++jj) {
          const float dsj = tds[t * KC + jb + jj];
          const float2 k2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sw_elem(sK, jb + jj, 2 * lane)));
          a0 += dsj * k2.x;
          a1 += dsj * k2.y;
        }
      }
      atomicAdd(&tdq[t * HD + 2 * lane], a0);
      atomicAdd(&tdq[t * HD + 2 * lane + 1], a1);
    }
    if (lane == 0 && tw < 4) TVT_STAMP(28 + tw);
  }
  tc_fence_before();
  __syncthreads();                  // #3: P, dS tiles and the tail rows' p / ds vectors complete; S / dP columns free
  if (tid == 0) TVT_STAMP(5);
  if (issuer) {
    tc_fence_after();
    // dV = P^T dO, dK = dS^T Q for keys 0..127: A MN-major (two 64-key blocks, LBO 16 KB), contraction over queries
    const uint32_t idesc_mn = make_idesc_bf16(128, HD, true, true);
#pragma unroll 1
    for (int k = 0; k < 8; ++k) {
      tc_mma_f16_ss(tdV, make_smem_desc_sw128(smem_u32(sP) + k * 2048, 16384, 1024), make_smem_desc_sw128(smem_u32(sdO) + k * 2048, 8192, 1024),
                    idesc_mn, k > 0);
      tc_mma_f16_ss(tdK, make_smem_desc_sw128(smem_u32(sdS) + k * 2048, 16384, 1024), make_smem_desc_sw128(smem_u32(sQ) + k * 2048, 8192, 1024),
                    idesc_mn, k > 0);
    }
    tc_commit(smem_u32(bar_g1));
    TVT_STAMP(18);
    // dQ = dS K: contraction over the nk keys
    const uint32_t idesc_q = make_idesc_bf16(128, HD, false, true);
#pragma unroll 1
    for (int k = 0; k < nk / 16; ++k)
      tc_mma_f16_ss(tdQ, make_smem_desc_sw128(smem_u32(sdS + (k >> 2) * 16384) + (k & 3) * 32, 16, 1024),
                    make_smem_desc_sw128(smem_u32(sK) + k * 2048, 8192, 1024), idesc_q, k > 0);
    if (nk > 128) {
      // tail keys 128..nk-1 from the third k-block, into the S columns: dV_t [0,64) | dK_t [64,128)
#pragma unroll 1
      for (int k = 0; k < 8; ++k) {
        tc_mma_f16_ss(tS, make_smem_desc_sw128(smem_u32(sP) + 2 * 16384 + k * 2048, 16384, 1024),
                      make_smem_desc_sw128(smem_u32(sdO) + k * 2048, 8192, 1024), idesc_mn, k > 0);
        tc_mma_f16_ss(tS + 64, make_smem_desc_sw128(smem_u32(sdS) + 2 * 16384 + k * 2048, 16384, 1024),
                      make_smem_desc_sw128(smem_u32(sQ) + k * 2048, 8192, 1024), idesc_mn, k > 0);
      }
    }
    tc_commit(smem_u32(bar_g2));
    TVT_STAMP(19);
  }
  if (worker) {
    // half 0 drains dK (+ ds_t q_t, x scale) then dQ; half 1 drains dV (+ p_t dO_t) then the tail keys.  thread == row;
    // each warp stages its 32 rows in its own 4 KB of the (by then consumed) P blocks 0 / 1
    __nv_bfloat16* gbase = (half ? p.dv : p.dk) + static_cast<long long>(b) * p.Sk * (half ? p.lddv : p.lddk) + h * HD;
    const long long gld = half ? p.lddv : p.lddk;
    const uint32_t w_s = smem_u32(half ? tp : tds), vec_s = smem_u32(half ? tdo : tq);
    const uint32_t stage_s = smem_u32(sP) + warp * 4096;
    const int r0 = (warp & 3) * 32;               // first row of this warp's TMEM quadrant
    mbar_wait(smem_u32(bar_g1), 0);
    tc_fence_after();
    if (tid == 0) TVT_STAMP(6);
    if (r0 < p.Sk) drain_rows((half ? tdV : tdK) + lane_addr, gbase + r0 * gld, gld, p.Sk - r0, half ? 1.0f : p.scale, ntail, w_s + row * 4, vec_s, stage_s);
    if (tid == 0) TVT_STAMP(7);
    mbar_wait(smem_u32(bar_g2), 0);
    tc_fence_after();
    if (tid == 0) TVT_STAMP(8);
    if (half == 0) {
      if (r0 < n_main)
        drain_rows(tdQ + lane_addr, p.dq + (static_cast<long long>(b) * p.Sq + r0) * p.lddq + h * HD, p.lddq, n_main - r0, p.scale, 0, 0, 0, stage_s);
    } else if (warp == 4 && nk > 128) {           // tail keys 128..: lanes of the first TMEM quadrant, dV_t at S[0,64)
      const int key = 128 + lane < KC ? 128 + lane : KC - 1;
      drain_rows(tS, gbase + 128 * gld, gld, p.Sk - 128, 1.0f, ntail, smem_u32(tp) + key * 4, smem_u32(tdo), stage_s);
    }
    if (tid == 0) TVT_STAMP(9);
  } else if (warp == 12) {
    if (nk > 128) {                               // dK_t at S[64,128): the other idle warp that may read TMEM lanes 0..31
      mbar_wait(smem_u32(bar_g2), 0);
      tc_fence_after();
      const int key = 128 + lane < KC ? 128 + lane : KC - 1;
      drain_rows(tS + 64, p.dk + (static_cast<long long>(b) * p.Sk + 128) * p.lddk + h * HD, p.lddk, p.Sk - 128, p.scale, ntail,
                 smem_u32(tds) + key * 4, smem_u32(tq), smem_u32(sP) + 8 * 4096);
    }
  } else if (warp == 9) {
    for (int t = 0; t < ntail; ++t) {
      __nv_bfloat16* dqrow = p.dq + (static_cast<long long>(b) * p.Sq + n_main + t) * p.lddq + h * HD;
      *reinterpret_cast<__nv_bfloat162*>(dqrow + 2 * lane) =
          __floats2bfloat162_rn(tdq[t * HD + 2 * lane] * p.scale, tdq[t * HD + 2 * lane + 1] * p.scale);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) TVT_STAMP(10);
  if (warp == 8) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// [rows, width] bf16 matrix with row pitch ld, boxes of 16 rows x 64 columns, 128B swizzle.
static int make_map(CUtensorMap* m, const void* ptr, long long rows, long long width, long long ld, int box_rows = 16) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return TVT_ECUDA;
  }
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(width), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("attention: cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
    return TVT_ECUDA;
  }
  return TVT_OK;
}

bool supported(long long sq, long long sk, long long hd) { return hd == HD && sq <= 256 && sk <= 272 && sq >= 1 && sk >= 1; }

// Opt a kernel in to the full 227 KB of dynamic shared memory, once per kernel per process.
template <typename K>
static int set_smem(K kern, size_t bytes, const char* what) {
  static std::mutex mu;
  static std::set<const void*> done;
  if (bytes > 227 * 1024) {
    set_last_error("%s: needs %zu bytes of shared memory", what, bytes);
    return TVT_EINVAL;
  }
  std::lock_guard<std::mutex> lock(mu);
  if (done.count(reinterpret_cast<const void*>(kern))) return TVT_OK;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) {
    set_last_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
    return TVT_ECUDA;
  }
  done.insert(reinterpret_cast<const void*>(kern));
  return TVT_OK;
}

int launch_fwd(const tvt_attention_fwd_args* a, cudaStream_t s) {
  Params p{};
  p.B = (int)a->batch; p.H = (int)a->heads; p.Sq = (int)a->sq; p.Sk = (int)a->sk; p.sk_pad = (p.Sk + 15) & ~15;
  p.scale = a->scale; p.o = (__nv_bfloat16*)a->o; p.ldo = a->ldo; p.lse = a->lse;
  if (a->dropout_p > 0.0f) {
    p.dropout_thr16 = (unsigned)(a->dropout_p * 65536.0f + 0.5f);
    p.dropout_scale = 65536.0f / (65536.0f - (float)p.dropout_thr16);
    p.dropout_seed = a->dropout_seed;
    p.seed_src = seed_source();
  }
  CUtensorMap tq, tk, tv;
  const long long w = a->heads * HD;
  int rc;
  p.kv_box = p.sk_pad % 48 == 0 ? 48 : (p.sk_pad % 32 == 0 ? 32 : 16);
  if (p.sk_pad % 128 == 0) p.kv_box = 128;
  p.q_in = (const __nv_bfloat16*)a->q; p.ldq = a->ldq;
  if ((rc = make_map(&tq, a->q, a->batch * a->sq, w, a->ldq, 128)) != TVT_OK) return rc;
  if ((rc = make_map(&tk, a->k, a->batch * a->sk, w, a->ldk, p.kv_box)) != TVT_OK) return rc;
  if ((rc = make_map(&tv, a->v, a->batch * a->sk, w, a->ldv, p.kv_box)) != TVT_OK) return rc;
  if (p.Sq <= 128 + kMaxTail && p.Sk <= 128 + kSmallTailKeys) {
    const size_t bytes_s = 1024 + 16384 + 2 * (size_t)KC * 128 + 64 + (8 + 4 * HD + KC) * 4;   // 55 KB: 4 CTAs / SM
    if ((rc = set_smem(fwd_small_kernel, bytes_s, "tvt_attention_fwd")) != TVT_OK) return rc;
    CUtensorMap tk128 = tk, tk8 = tk, tq8 = tq;
    static const int tu_enabled = [] { const char* e = getenv("TVT_ATTN_TU"); return e ? atoi(e) : 1; }();
    p.use_tu = tu_enabled && (p.Sq > 128 || p.Sk > 128) && p.Sq >= 128 && p.Sk >= 128;
    static const int pf_enabled = [] { const char* e = getenv("TVT_ATTN_FWD_PREFETCH"); return e ? atoi(e) : 1; }();
    p.prefetch_stride = pf_enabled ? 4 * num_sms() : 0;   // four resident CTAs per SM
    if (p.use_tu) {
      if ((rc = make_map(&tk128, a->k, a->batch * a->sk, w, a->ldk, 128)) != TVT_OK) return rc;
      if ((rc = make_map(&tk8, a->k, a->batch * a->sk, w, a->ldk, 8)) != TVT_OK) return rc;
      if ((rc = make_map(&tq8, a->q, a->batch * a->sq, w, a->ldq, 8)) != TVT_OK) return rc;
    }
    if (launch_pdl(fwd_small_kernel, p.B * p.H, kThreads, bytes_s, s, 1, tq, tk, tv, tk128, tk8, tq8, p) != cudaSuccess) return check_launch("tvt_attention_fwd");
    return check_launch("tvt_attention_fwd");
  }
  const int kv = ((p.sk_pad * 128) + 1023) & ~1023;
  const size_t bytes = 1024 + 16384 + 2 * (size_t)kv + (size_t)((p.sk_pad + 63) / 64) * 16384 + 64 + (HD + p.sk_pad) * 4;
  if ((rc = set_smem(fwd_kernel, bytes, "tvt_attention_fwd")) != TVT_OK) return rc;
  fwd_kernel<<<p.B * p.H, kThreadsFwd, bytes, s>>>(tq, tk, tv, p);
  return check_launch("tvt_attention_fwd");
}

int launch_bwd(const tvt_attention_bwd_args* a, cudaStream_t s) {
  Params p{};
  p.B = (int)a->batch; p.H = (int)a->heads; p.Sq = (int)a->sq; p.Sk = (int)a->sk; p.sk_pad = (p.Sk + 15) & ~15;
  p.scale = a->scale; p.lse = const_cast<float*>(a->lse);
  p.o_in = (const __nv_bfloat16*)a->o; p.ldo = a->ldo; p.do_in = (const __nv_bfloat16*)a->d_o; p.lddo = a->lddo;
  p.dq = (__nv_bfloat16*)a->dq; p.dk = (__nv_bfloat16*)a->dk; p.dv = (__nv_bfloat16*)a->dv;
  p.lddq = a->lddq; p.lddk = a->lddk; p.lddv = a->lddv;
  if (a->dropout_p > 0.0f) {
    p.dropout_thr16 = (unsigned)(a->dropout_p * 65536.0f + 0.5f);
    p.dropout_scale = 65536.0f / (65536.0f - (float)p.dropout_thr16);
    p.dropout_seed = a->dropout_seed;
    p.seed_src = seed_source();
  }
  CUtensorMap tq, tk, tv, tdo;
  const long long w = a->heads * HD;
  int rc;
  if ((rc = make_map(&tq, a->q, a->batch * a->sq, w, a->ldq)) != TVT_OK) return rc;
  if ((rc = make_map(&tk, a->k, a->batch * a->sk, w, a->ldk)) != TVT_OK) return rc;
  if ((rc = make_map(&tv, a->v, a->batch * a->sk, w, a->ldv)) != TVT_OK) return rc;
  if ((rc = make_map(&tdo, a->d_o, a->batch * a->sq, w, a->lddo)) != TVT_OK) return rc;
  if (p.Sq <= 128 + kMaxTail && p.sk_pad <= KC) {
    CUtensorMap tq128, tdo128, to128, tk1, tv1;     // one box per operand
    if ((rc = make_map(&tq128, a->q, a->batch * a->sq, w, a->ldq, 128)) != TVT_OK) return rc;
    if ((rc = make_map(&tdo128, a->d_o, a->batch * a->sq, w, a->lddo, 128)) != TVT_OK) return rc;
    if ((rc = make_map(&to128, a->o, a->batch * a->sq, w, a->ldo, 128)) != TVT_OK) return rc;
    if ((rc = make_map(&tk1, a->k, a->batch * a->sk, w, a->ldk, p.sk_pad)) != TVT_OK) return rc;
    if ((rc = make_map(&tv1, a->v, a->batch * a->sk, w, a->ldv, p.sk_pad)) != TVT_OK) return rc;
    p.q_in = (const __nv_bfloat16*)a->q; p.ldq = a->ldq;
    const size_t bytes3 = 1024 + 2 * 16384 + 6 * 16384 + 2 * (size_t)KC * 128 + 64 + 256 * 4 + (3 * kMaxTail * HD + 2 * kMaxTail * KC + 2 * kMaxTail) * 4;
    if ((rc = set_smem(bwd3_kernel, bytes3, "tvt_attention_bwd")) != TVT_OK) return rc;
    p.prefetch_stride = num_sms();
    launch_pdl(bwd3_kernel, p.B * p.H, kThreadsBwd3, bytes3, s, 1, tq128, tk1, tv1, tdo128, to128, p);
    return check_launch("tvt_attention_bwd");
  }
  if (p.Sq <= 128 + kMaxTail) {
    CUtensorMap tq128, tdo128;
    if ((rc = make_map(&tq128, a->q, a->batch * a->sq, w, a->ldq, 128)) != TVT_OK) return rc;
    if ((rc = make_map(&tdo128, a->d_o, a->batch * a->sq, w, a->lddo, 128)) != TVT_OK) return rc;
    p.q_in = (const __nv_bfloat16*)a->q; p.ldq = a->ldq;
    const size_t bytes2 = 1024 + 2 * 16384 + 6 * 16384 + 2 * (size_t)KC * 128 + 64 + 384 * 4 + (3 * kMaxTail * HD + 2 * kMaxTail * KC + 2 * kMaxTail) * 4;
    if ((rc = set_smem(bwd2_kernel, bytes2, "tvt_attention_bwd")) != TVT_OK) return rc;
    bwd2_kernel<<<p.B * p.H, kThreadsBwd2, bytes2, s>>>(tq128, tk, tv, tdo128, p);
    return check_launch("tvt_attention_bwd");
  }
  const int m_tiles = (p.Sq + 127) / 128;
  const size_t bytes = 1024 + (size_t)m_tiles * 32768 + 2 * 16384 + 2 * 32768 + 64 + (size_t)m_tiles * 128 * 8;
  if ((rc = set_smem(bwd_kernel, bytes, "tvt_attention_bwd")) != TVT_OK) return rc;
  bwd_kernel<<<p.B * p.H, kThreads, bytes, s>>>(tq, tk, tv, tdo, p);
  return check_launch("tvt_attention_bwd");
}

}  // namespace attn_tc
}  // namespace tvt

#ifdef TVT_ATTN_STAMPS
extern "C" __attribute__((visibility("default"))) int tvt_debug_attn_stamps(long long* out) {
  return cudaMemcpyFromSymbol(out, tvt::attn_tc::g_stamps, sizeof(long long) * 3 * 32) == cudaSuccess ? 0 : -1;
}
#endif
