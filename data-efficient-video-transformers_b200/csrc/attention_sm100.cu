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

#include "tvt_common.cuh"
#include "tvt_ptx.cuh"

namespace tvt {
namespace attn_tc {

constexpr int HD = 64;        // head dim: one 128-byte swizzle atom
constexpr int kThreads = 160;  // warps 0-3: one row per thread; warp 4: TMA / MMA issue
constexpr float kLog2e = 1.4426950408889634f;

struct Params {
  int B, H, Sq, Sk;
  int sk_pad;                 // Sk rounded up to 16
  float scale;
  float dropout_scale; unsigned dropout_thr16; unsigned long long dropout_seed;
  __nv_bfloat16* o; long long ldo;
  float* lse;
  // backward
  const __nv_bfloat16* o_in; const __nv_bfloat16* do_in; long long lddo;
  __nv_bfloat16* dq; __nv_bfloat16* dk; __nv_bfloat16* dv; long long lddq, lddk, lddv;
};

// Byte offset of the 16-byte chunk `chunk` (0..7) of row `row` inside a [rows x 128 B] swizzled tile.
__device__ __forceinline__ uint32_t sw128(int row, int chunk) {
  return static_cast<uint32_t>(row) * 128u + (static_cast<uint32_t>(chunk ^ (row & 7)) << 4);
}

__device__ __forceinline__ float drop_mul(const Params& p, long long bh, int i, int j) {
  if (!p.dropout_thr16) return 1.0f;
  const unsigned long long e = (static_cast<unsigned long long>(bh) * p.Sq + i) * p.Sk + j;
  return dropout_keep(p.dropout_seed, e, p.dropout_thr16) ? p.dropout_scale : 0.0f;
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
// smem: Q tile 16 KB | K sk_pad*128 | V sk_pad*128 | P ceil(sk_pad/64)*16 KB | barriers
__global__ void __launch_bounds__(kThreads) fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
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

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool issuer = warp == 4 && lane == 0;
  const int bh = blockIdx.x, b = bh / p.H, h = bh % p.H;
  const int o_col = (p.sk_pad + 63) & ~63;           // O accumulator starts on a 64-column boundary
  const uint32_t tmem_cols = o_col + HD <= 128 ? 128u : (o_col + HD <= 256 ? 256u : 512u);

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

  if (issuer) {
    mbar_arrive_expect_tx(smem_u32(bar_kv), 2 * kv_bytes);
    for (int r = 0; r < p.sk_pad; r += 16) {
      tma_load_2d(smem_u32(sK + r * 128), &tmK, smem_u32(bar_kv), h * HD, b * p.Sk + r);
      tma_load_2d(smem_u32(sV + r * 128), &tmV, smem_u32(bar_kv), h * HD, b * p.Sk + r);
    }
  }
  const int m_tiles = (p.Sq + 127) / 128;
  uint32_t phase = 0;
  const float sl2 = p.scale * kLog2e;
  for (int mt = 0; mt < m_tiles; ++mt, phase ^= 1) {
    if (issuer) {
      mbar_arrive_expect_tx(smem_u32(bar_q), 128 * 128);
      for (int r = 0; r < 128; r += 16) tma_load_2d(smem_u32(sQ + r * 128), &tmQ, smem_u32(bar_q), h * HD, b * p.Sq + mt * 128 + r);
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
      for (int c0 = 0; c0 < p.sk_pad; c0 += 16) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(tmem_S + lane_addr + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (c0 + i < p.Sk) mx = fmaxf(mx, __uint_as_float(r[i]));
      }
      const float mxs = mx * sl2;
      const int drow = row_ok ? row : 0;
      for (int c0 = 0; c0 < p.sk_pad; c0 += 16) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(tmem_S + lane_addr + c0, r);
        tmem_ld_wait();
        uint32_t packed[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          float e0 = 0.0f, e1 = 0.0f;
          if (c0 + i < p.Sk) { e0 = exp2f(__uint_as_float(r[i]) * sl2 - mxs); sum += e0; e0 *= drop_mul(p, bh, drow, c0 + i); }
          if (c0 + i + 1 < p.Sk) { e1 = exp2f(__uint_as_float(r[i + 1]) * sl2 - mxs); sum += e1; e1 *= drop_mul(p, bh, drow, c0 + i + 1); }
          packed[i >> 1] = pack_bf16x2(e0, e1);
        }
        // 16 columns = two 16-byte chunks of row `tid` in k-block c0/64
        uint8_t* blk = sP + (c0 >> 6) * 16384;
        const int ch = (c0 & 63) >> 3;
        *reinterpret_cast<uint4*>(blk + sw128(tid, ch)) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        *reinterpret_cast<uint4*>(blk + sw128(tid, ch + 1)) = make_uint4(packed[4], packed[5], packed[6], packed[7]);
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
      for (int c0 = 0; c0 < HD; c0 += 16) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(tmem_O + lane_addr + c0, r);
        tmem_ld_wait();
        if (row_ok) {
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 16; i += 2) pk[i >> 1] = pack_bf16x2(__uint_as_float(r[i]) * inv, __uint_as_float(r[i + 1]) * inv);
          *reinterpret_cast<uint4*>(orow + c0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(orow + c0 + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      }
      if (row_ok && p.lse) p.lse[static_cast<long long>(bh) * p.Sq + row] = mx * p.scale + __logf(sum);
    }
    tc_fence_before();
    __syncthreads();   // TMEM S/O, sQ and sP are reused by the next m-tile
    tc_fence_after();
  }
  if (warp == 4) tmem_dealloc(tmem_base, tmem_cols);
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
            tmem_ld_wait();
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
        tmem_ld_wait();
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
          tmem_ld_wait();
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
static int make_map(CUtensorMap* m, const void* ptr, long long rows, long long width, long long ld) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return TVT_ECUDA;
  }
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(width), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {64, 16};
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

template <typename K>
static int set_smem(K kern, size_t bytes, const char* what) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
  if (e != cudaSuccess) {
    set_last_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
    return TVT_ECUDA;
  }
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
  }
  CUtensorMap tq, tk, tv;
  const long long w = a->heads * HD;
  int rc;
  if ((rc = make_map(&tq, a->q, a->batch * a->sq, w, a->ldq)) != TVT_OK) return rc;
  if ((rc = make_map(&tk, a->k, a->batch * a->sk, w, a->ldk)) != TVT_OK) return rc;
  if ((rc = make_map(&tv, a->v, a->batch * a->sk, w, a->ldv)) != TVT_OK) return rc;
  const int kv = ((p.sk_pad * 128) + 1023) & ~1023;
  const size_t bytes = 1024 + 16384 + 2 * (size_t)kv + (size_t)((p.sk_pad + 63) / 64) * 16384 + 64;
  if ((rc = set_smem(fwd_kernel, bytes, "tvt_attention_fwd")) != TVT_OK) return rc;
  fwd_kernel<<<p.B * p.H, kThreads, bytes, s>>>(tq, tk, tv, p);
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
  }
  CUtensorMap tq, tk, tv, tdo;
  const long long w = a->heads * HD;
  int rc;
  if ((rc = make_map(&tq, a->q, a->batch * a->sq, w, a->ldq)) != TVT_OK) return rc;
  if ((rc = make_map(&tk, a->k, a->batch * a->sk, w, a->ldk)) != TVT_OK) return rc;
  if ((rc = make_map(&tv, a->v, a->batch * a->sk, w, a->ldv)) != TVT_OK) return rc;
  if ((rc = make_map(&tdo, a->d_o, a->batch * a->sq, w, a->lddo)) != TVT_OK) return rc;
  const int m_tiles = (p.Sq + 127) / 128;
  const size_t bytes = 1024 + (size_t)m_tiles * 32768 + 2 * 16384 + 2 * 32768 + 64 + (size_t)m_tiles * 128 * 8;
  if ((rc = set_smem(bwd_kernel, bytes, "tvt_attention_bwd")) != TVT_OK) return rc;
  bwd_kernel<<<p.B * p.H, kThreads, bytes, s>>>(tq, tk, tv, tdo, p);
  return check_launch("tvt_attention_bwd");
}

}  // namespace attn_tc
}  // namespace tvt
