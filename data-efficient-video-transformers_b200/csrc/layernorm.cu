// LayerNorm forward / backward (see include/tvt.h: tvt_layernorm_fwd / tvt_layernorm_bwd), one warp
// per token row, the row held in registers (16-byte loads), fp32 statistics via warp shuffles.
// "Embed" row mapping fuses SimpleTransformer.add_pos_cls (reference src/models/transformer.py:74-82):
// CLS concat + positional-encoding add + (dropout) + LayerNorm in one pass, and its backward scatter.
#include "tvt_common.cuh"
#include "tvt_ptx.cuh"

namespace tvt {
namespace ln {

constexpr int kWarps = 8;

struct FwdParams {
  const void* x;      // [rows, d]  (or, embed mode: frame features [B, T, d])
  const void* cls;    // embed mode: [B, d] CLS rows, same dtype as x
  const float* pe;    // embed mode: [S, d] fp32 positional encoding
  const float* gamma; const float* beta;
  void* y;            // [rows, d]
  void* pre;          // embed mode: optional copy of the pre-LN row (x + pe), for backward
  float* mean; float* rstd;
  long long rows; int d; int S;  // S > 0 selects embed mode (rows = B * S)
  float eps;
  float dropout_scale; unsigned dropout_thr16; unsigned long long dropout_seed; const unsigned long long* seed_src;
  int y_seq; long long y_pitch;   // ring kernels, optional: output row r goes to y + (r / y_seq) * y_pitch + (r % y_seq) * d elements
                      // (the rows of one clip land inside a wider per-clip block: no concatenation pass afterwards)
  int reverse;        // ring kernels: walk the rows from the LAST pair to the first.  The producer (a GEMM epilogue) has just
                      // written these rows in ascending order through an L2 smaller than its traffic, so the rows written last
                      // are the ones still resident: reading them first turns DRAM reads into L2 hits
};

template <typename T, int kChunks>
__global__ void __launch_bounds__(kWarps * 32, kChunks <= 4 ? 3 : 1) ln_fwd_kernel(const FwdParams p) {
  constexpr int V = Vec16<T>::kN;
  const int lane = threadIdx.x & 31;
  const long long warp0 = static_cast<long long>(blockIdx.x) * kWarps + (threadIdx.x >> 5);
  const long long nwarps = static_cast<long long>(gridDim.x) * kWarps;
  // gamma / beta live in shared memory (each lane re-reads its own columns per row, conflict-free): holding them in
  // registers cost 16 * kChunks registers per thread, and this kernel lives on occupancy (bytes in flight)
  __shared__ __align__(16) float gam_s[kChunks * 32 * V], bet_s[kChunks * 32 * V];
  for (int t = threadIdx.x; t < kChunks * 32 * V; t += kWarps * 32) {
    gam_s[t] = t < p.d ? p.gamma[t] : 0.0f;
    bet_s[t] = t < p.d ? p.beta[t] : 0.0f;
  }
  __syncthreads();
  auto src_of = [&](long long row, const float*& pe_row) -> const T* {
    if (p.S > 0) {
      const long long b = row / p.S;
      const int s = static_cast<int>(row - b * p.S);
      pe_row = p.pe + static_cast<long long>(s) * p.d;
      return s == 0 ? reinterpret_cast<const T*>(p.cls) + b * p.d
                    : reinterpret_cast<const T*>(p.x) + (b * (p.S - 1) + (s - 1)) * p.d;
    }
    pe_row = nullptr;
    return reinterpret_cast<const T*>(p.x) + row * p.d;
  };
  uint4 raw[kChunks], nxt[kChunks];
  const float* pe_row = nullptr;
  if (warp0 < p.rows) {
    const T* src = src_of(warp0, pe_row);
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int col = (c * 32 + lane) * V;
      raw[c] = col < p.d ? ldg_stream(src + col) : make_uint4(0, 0, 0, 0);
    }
  }
  for (long long row = warp0; row < p.rows; row += nwarps) {
    // software pipeline: the next row's 16-byte loads are in flight while this row is reduced
    const float* pe_next = nullptr;
    if (row + nwarps < p.rows) {
      const T* src = src_of(row + nwarps, pe_next);
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        const int col = (c * 32 + lane) * V;
        nxt[c] = col < p.d ? ldg_stream(src + col) : make_uint4(0, 0, 0, 0);
      }
    }
    float v[kChunks][V];
    float sum = 0.0f;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int col = (c * 32 + lane) * V;
      Vec16<T>::unpack(raw[c], v[c]);
      if (col < p.d) {
        if (pe_row) {
#pragma unroll
          for (int i = 0; i < V; ++i) v[c][i] += pe_row[col + i];
          if (p.dropout_thr16) {   // d % 8 == 0: the row's element index is a multiple of 4, so one hash serves 4 elements
            const unsigned long long e4 = (static_cast<unsigned long long>(row) * p.d + col) >> 2;
#pragma unroll
            for (int g = 0; g < V / 4; ++g) {
              const uint64_t bits = dropout_bits4(mix_seed(p.dropout_seed, p.seed_src), e4 + g);
#pragma unroll
              for (int i = 0; i < 4; ++i)
                v[c][4 * g + i] = dropout_keep_lane(bits, i, p.dropout_thr16) ? v[c][4 * g + i] * p.dropout_scale : 0.0f;
            }
          }
          if (p.pre) Vec16<T>::store(reinterpret_cast<T*>(p.pre) + row * p.d + col, v[c]);
        }
#pragma unroll
        for (int i = 0; i < V; ++i) sum += v[c][i];
      }
    }
    const float mean = warp_sum(sum) / p.d;
    float sq = 0.0f;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int col = (c * 32 + lane) * V;
      if (col < p.d) {
#pragma unroll
        for (int i = 0; i < V; ++i) { const float t = v[c][i] - mean; sq += t * t; }
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) / p.d + p.eps);
    if (lane == 0) {
      if (p.mean) p.mean[row] = mean;
      if (p.rstd) p.rstd[row] = rstd;
    }
    T* dst = reinterpret_cast<T*>(p.y) + row * p.d;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int col = (c * 32 + lane) * V;
      if (col < p.d) {
        float o[V];
#pragma unroll
        for (int i = 0; i < V; i += 4) {
          const float4 g4 = *reinterpret_cast<const float4*>(gam_s + col + i);
          const float4 b4 = *reinterpret_cast<const float4*>(bet_s + col + i);
          o[i] = (v[c][i] - mean) * rstd * g4.x + b4.x;
          o[i + 1] = (v[c][i + 1] - mean) * rstd * g4.y + b4.y;
          o[i + 2] = (v[c][i + 2] - mean) * rstd * g4.z + b4.z;
          o[i + 3] = (v[c][i + 3] - mean) * rstd * g4.w + b4.w;
        }
        Vec16<T>::store(dst + col, o);
      }
    }
#pragma unroll
    for (int c = 0; c < kChunks; ++c) raw[c] = nxt[c];
    pe_row = pe_next;
  }
}

// ---- plain-mode forward with the rows staged by the bulk-copy engine -------------------------------------------------
// ncu on the register-prefetch kernel above: half issue-bound (423 instructions per row at 52 % issue utilisation) and half
// latency-bound on the two dependent shuffle trees of every row (short-scoreboard stalls), with loads in flight limited by
// registers.  Here every warp owns a private ring of kRing slots in shared memory, each slot holding a PAIR of consecutive
// rows fetched by one cp.async.bulk (lane 0 issues, one mbarrier per slot), and normalises both rows of a slot together:
//   * bytes in flight no longer depend on registers or on the warp's instruction stream;
//   * each of the two reduction passes (mean, centred sum of squares: torch's arithmetic) runs ONE shuffle tree for both
//     rows of the pair, i.e. two trees per pair instead of four (a shifted single pass was tried: 5x larger rstd error,
//     enough to push one fp32-mode gradient of the cross-attention model past 1e-3);
//   * gamma / beta are read from shared memory once per pair; no column predicates when d fills the chunks exactly.
// No block-level synchronisation after the prologue.
constexpr int kRing = 3;
constexpr int kRingWarps = 4;

__device__ __forceinline__ void bulk_load_rows(uint32_t dst_s, const void* src, uint32_t bytes, uint32_t bar_s) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_s), "l"(src), "r"(bytes), "r"(bar_s) : "memory");
}
__device__ __forceinline__ void bar_wait(uint32_t bar_s, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar_s), "r"(parity) : "memory");
}

template <typename T, int kChunks, bool kExact>
__global__ void __launch_bounds__(kRingWarps * 32) ln_fwd_ring_kernel(const FwdParams p) {
  constexpr int V = Vec16<T>::kN;
  extern __shared__ __align__(128) uint8_t ring_raw[];
  __shared__ __align__(16) float gam_s[kChunks * 32 * V], bet_s[kChunks * 32 * V];
  __shared__ __align__(8) unsigned long long bars[kRingWarps][kRing];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long pair0 = static_cast<long long>(blockIdx.x) * kRingWarps + warp;
  const long long npairs = (p.rows + 1) >> 1, stride = static_cast<long long>(gridDim.x) * kRingWarps;
  const uint32_t row_bytes = static_cast<uint32_t>(p.d) * sizeof(T);
  const uint32_t ring_s = static_cast<uint32_t>(__cvta_generic_to_shared(ring_raw)) + warp * kRing * 2 * row_bytes;
  const uint32_t bar_s = static_cast<uint32_t>(__cvta_generic_to_shared(&bars[warp][0]));
  auto fetch = [&](long long qi, int slot) {     // rows 2q, 2q + 1 (the last pair may be a single row)
    const long long q = p.reverse ? npairs - 1 - qi : qi;
    const uint32_t bytes = 2 * q + 1 < p.rows ? 2 * row_bytes : row_bytes;
    bulk_load_rows(ring_s + slot * 2 * row_bytes, reinterpret_cast<const T*>(p.x) + 2 * q * p.d, bytes, bar_s + 8 * slot);
  };
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kRing; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s + 8 * s) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  pdl_wait();
  pdl_trigger();
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kRing; ++s)
      if (pair0 + s * stride < npairs) fetch(pair0 + s * stride, s);
  }
  for (int t = threadIdx.x; t < kChunks * 32 * V; t += kRingWarps * 32) {
    gam_s[t] = t < p.d ? p.gamma[t] : 0.0f;
    bet_s[t] = t < p.d ? p.beta[t] : 0.0f;
  }
  __syncthreads();
  const float inv_d = 1.0f / p.d;
  int slot = 0;
  uint32_t parity = 0;
  for (long long qi = pair0; qi < npairs; qi += stride) {
    const long long q = p.reverse ? npairs - 1 - qi : qi;     // see FwdParams::reverse
    bar_wait(bar_s + 8 * slot, parity);
    float v[2][kChunks][V];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        const int col = (c * 32 + lane) * V;
        uint4 raw = make_uint4(0, 0, 0, 0);
        if (kExact || col < p.d)
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w)
                       : "r"(ring_s + (slot * 2 + r) * row_bytes + col * static_cast<uint32_t>(sizeof(T))) : "memory");
        Vec16<T>::unpack(raw, v[r][c]);
      }
    }
    // two passes over the registers (mean, then centred sum of squares: same arithmetic as torch), one shuffle tree per
    // pass carrying both rows of the pair
    float s1[2] = {0.0f, 0.0f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
#pragma unroll
        for (int i = 0; i < V; ++i) s1[r] += v[r][c][i];   // columns beyond d hold zeros
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1[0] += __shfl_xor_sync(0xffffffffu, s1[0], o);
      s1[1] += __shfl_xor_sync(0xffffffffu, s1[1], o);
    }
    // every lane's slot data went into the shuffle tree above, so the slot can be refilled: kRing pairs ahead
    if (lane == 0 && qi + kRing * stride < npairs) fetch(qi + kRing * stride, slot);
    const float mean[2] = {s1[0] * inv_d, s1[1] * inv_d};
    float s2[2] = {0.0f, 0.0f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        const int col = (c * 32 + lane) * V;
        if (kExact || col < p.d) {
#pragma unroll
          for (int i = 0; i < V; ++i) {
            const float t = v[r][c][i] - mean[r];
            s2[r] = fmaf(t, t, s2[r]);
          }
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s2[0] += __shfl_xor_sync(0xffffffffu, s2[0], o);
      s2[1] += __shfl_xor_sync(0xffffffffu, s2[1], o);
    }
    float rstd[2], shift[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      rstd[r] = rsqrtf(s2[r] * inv_d + p.eps);
      shift[r] = -mean[r] * rstd[r];
      if (lane == 0 && 2 * q + r < p.rows) {
        if (p.mean) p.mean[2 * q + r] = mean[r];
        if (p.rstd) p.rstd[2 * q + r] = rstd[r];
      }
    }
    T* dst = reinterpret_cast<T*>(p.y) + 2 * q * p.d;
    T* dst1 = dst + p.d;
    if (p.y_seq > 0) {
      const long long r0 = 2 * q, r1 = 2 * q + 1;
      dst = reinterpret_cast<T*>(p.y) + (r0 / p.y_seq) * p.y_pitch + (r0 % p.y_seq) * p.d;
      dst1 = reinterpret_cast<T*>(p.y) + (r1 / p.y_seq) * p.y_pitch + (r1 % p.y_seq) * p.d;
    }
    const bool two = 2 * q + 1 < p.rows;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int col = (c * 32 + lane) * V;
      if (kExact || col < p.d) {
        float o0[V], o1[V];
#pragma unroll
        for (int i = 0; i < V; i += 4) {
          const float4 g4 = *reinterpret_cast<const float4*>(gam_s + col + i);
          const float4 b4 = *reinterpret_cast<const float4*>(bet_s + col + i);
          const float gm[4] = {g4.x, g4.y, g4.z, g4.w}, bt[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if constexpr (sizeof(T) == 4) {      // fp32 parity mode: torch's (x - mean) * rstd * gamma + beta, rounding for rounding
              o0[i + u] = (v[0][c][i + u] - mean[0]) * rstd[0] * gm[u] + bt[u];
              o1[i + u] = (v[1][c][i + u] - mean[1]) * rstd[1] * gm[u] + bt[u];
            } else {                             // bf16 output: two FMAs per element (the bf16 rounding is 1e4 x any difference)
              o0[i + u] = fmaf(fmaf(v[0][c][i + u], rstd[0], shift[0]), gm[u], bt[u]);
              o1[i + u] = fmaf(fmaf(v[1][c][i + u], rstd[1], shift[1]), gm[u], bt[u]);
            }
          }
        }
        Vec16<T>::store(dst + col, o0);
        if (two) Vec16<T>::store(dst1 + col, o1);
      }
    }
    if (++slot == kRing) { slot = 0; parity ^= 1; }
  }
}

struct BwdParams {
  const void* dy;     // grad wrt LN output [rows, d]
  const void* x;      // pre-LN input [rows, d]
  const float* mean; const float* rstd; const float* gamma;
  void* dx;           // grad wrt pre-LN input [rows, d]            (nullable in embed mode)
  void* dz;           // optional: dropout-masked copy of dx (the branch gradient) [rows, d]
  void* dfeat;        // embed mode: grad wrt frame features [B, T, d]
  void* dcls;         // embed mode: grad wrt CLS rows [B, d]
  float* dgamma; float* dbeta;  // [d], accumulated with atomics (caller zero-fills)
  float* dbias;       // optional [d]: column sum of dz (or dx when dz is null)
  long long rows; int d; int S;
  float dropout_scale; unsigned dropout_thr16; unsigned long long dropout_seed; const unsigned long long* seed_src;
  long long dropout_ld;  // row pitch used for dropout element indices (N of the producing GEMM)
  const void* dres;      // optional residual-path gradient added to dx
};

constexpr int kBwdWarps = 4;

template <typename T, int kChunks>
__global__ void __launch_bounds__(kBwdWarps * 32, kChunks <= 4 ? 3 : 1) ln_bwd_kernel(const BwdParams p) {
  constexpr int V = Vec16<T>::kN;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const long long warp0 = static_cast<long long>(blockIdx.x) * kBwdWarps + warp;
  const long long nwarps = static_cast<long long>(gridDim.x) * kBwdWarps;
  // Register budget decides this kernel's speed (occupancy = bytes in flight): gamma is re-read from shared memory
  // and dy * gamma / xhat are recomputed in the second pass instead of being held; what must stay in registers are
  // the three column accumulators and the packed rows (current + prefetched).
  __shared__ __align__(16) float gam_s[kChunks * 32 * V];
  for (int t = threadIdx.x; t < kChunks * 32 * V; t += kBwdWarps * 32) gam_s[t] = t < p.d ? p.gamma[t] : 0.0f;
  __syncthreads();
  float acc_g[kChunks][V], acc_b[kChunks][V], acc_z[kChunks][V];
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
#pragma unroll
    for (int i = 0; i < V; ++i) acc_g[c][i] = acc_b[c][i] = acc_z[c][i] = 0.0f;
  }

  uint4 rdy[kChunks], rx[kChunks], ndy[kChunks], nx[kChunks];
  if (warp0 < p.rows) {
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int col = (c * 32 + lane) * V;
      rdy[c] = rx[c] = make_uint4(0, 0, 0, 0);
      if (col < p.d) {
        rdy[c] = ldg_stream(reinterpret_cast<const T*>(p.dy) + warp0 * p.d + col);
        rx[c] = ldg_stream(reinterpret_cast<const T*>(p.x) + warp0 * p.d + col);
      }
    }
  }
  for (long long row = warp0; row < p.rows; row += nwarps) {
    if (row + nwarps < p.rows) {   // next row's loads in flight while this row is processed
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        const int col = (c * 32 + lane) * V;
        ndy[c] = nx[c] = make_uint4(0, 0, 0, 0);
        if (col < p.d) {
          ndy[c] = ldg_stream(reinterpret_cast<const T*>(p.dy) + (row + nwarps) * p.d + col);
          nx[c] = ldg_stream(reinterpret_cast<const T*>(p.x) + (row + nwarps) * p.d + col);
        }
      }
    }
    const float mean = p.mean[row], rstd = p.rstd[row];
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int col = (c * 32 + lane) * V;
      if (col < p.d) {
        float dyv[V], xv[V], gm[V];
        Vec16<T>::unpack(rdy[c], dyv);
        Vec16<T>::unpack(rx[c], xv);
#pragma unroll
        for (int i = 0; i < V; i += 4) *reinterpret_cast<float4*>(gm + i) = *reinterpret_cast<const float4*>(gam_s + col + i);
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float xh = (xv[i] - mean) * rstd;
          acc_g[c][i] += dyv[i] * xh;
          acc_b[c][i] += dyv[i];
          const float g = dyv[i] * gm[i];
          s1 += g;
          s2 += g * xh;
        }
      }
    }
    s1 = warp_sum(s1) / p.d;
    s2 = warp_sum(s2) / p.d;
    // destination rows (embed mode scatters to the feature / CLS gradients)
    T* dx_row = p.dx ? reinterpret_cast<T*>(p.dx) + row * p.d : nullptr;
    T* dz_row = p.dz ? reinterpret_cast<T*>(p.dz) + row * p.d : nullptr;
    if (p.S > 0) {
      const long long b = row / p.S;
      const int s = static_cast<int>(row - b * p.S);
      dz_row = s == 0 ? (p.dcls ? reinterpret_cast<T*>(p.dcls) + b * p.d : nullptr)
                      : (p.dfeat ? reinterpret_cast<T*>(p.dfeat) + (b * (p.S - 1) + (s - 1)) * p.d : nullptr);
    }
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int col = (c * 32 + lane) * V;
      if (col < p.d) {
        float o[V], dyv[V], xv[V], gm[V];
        Vec16<T>::unpack(rdy[c], dyv);
        Vec16<T>::unpack(rx[c], xv);
#pragma unroll
        for (int i = 0; i < V; i += 4) *reinterpret_cast<float4*>(gm + i) = *reinterpret_cast<const float4*>(gam_s + col + i);
#pragma unroll
        for (int i = 0; i < V; ++i) o[i] = rstd * (dyv[i] * gm[i] - s1 - (xv[i] - mean) * rstd * s2);
        if (p.dres) {
          float rr[V];
          Vec16<T>::load(reinterpret_cast<const T*>(p.dres) + row * p.d + col, rr);
#pragma unroll
          for (int i = 0; i < V; ++i) o[i] += rr[i];
        }
        if (dx_row) Vec16<T>::store(dx_row + col, o);
        if (p.dropout_thr16) {   // dropout_ld % 4 == 0 and col % 4 == 0: one hash per 4 consecutive elements
          const unsigned long long e4 = (static_cast<unsigned long long>(row) * p.dropout_ld + col) >> 2;
#pragma unroll
          for (int g = 0; g < V / 4; ++g) {
            const uint64_t bits = dropout_bits4(mix_seed(p.dropout_seed, p.seed_src), e4 + g);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              o[4 * g + i] = dropout_keep_lane(bits, i, p.dropout_thr16) ? o[4 * g + i] * p.dropout_scale : 0.0f;
          }
        }
        if (dz_row) Vec16<T>::store(dz_row + col, o);
#pragma unroll
        for (int i = 0; i < V; ++i) acc_z[c][i] += o[i];
      }
    }
#pragma unroll
    for (int c = 0; c < kChunks; ++c) { rdy[c] = ndy[c]; rx[c] = nx[c]; }
  }
  // CTA-level reduction of the column partials through shared memory, then one atomic per column.
  __shared__ float red[kBwdWarps][32 * V + 1];
  for (int which = 0; which < 3; ++which) {
    float* out = which == 0 ? p.dgamma : (which == 1 ? p.dbeta : p.dbias);
    if (!out) continue;  // uniform
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < V; ++i)
        red[warp][lane * V + i] = which == 0 ? acc_g[c][i] : (which == 1 ? acc_b[c][i] : acc_z[c][i]);
      __syncthreads();
      for (int t = threadIdx.x; t < 32 * V; t += kBwdWarps * 32) {
        float s = 0.0f;
#pragma unroll
        for (int w = 0; w < kBwdWarps; ++w) s += red[w][t];
        const int col = c * 32 * V + t;
        if (col < p.d) atomicAdd(out + col, s);
      }
    }
  }
}

// ---- plain-mode backward on the same bulk-copy ring ------------------------------------------------------------------------
// ncu on ln_bwd_kernel: 793 instructions per row at 48 % issue utilisation, i.e. instruction-bound (both passes unpack dy and
// x, the second pass recomputes x_hat and dy * gamma, every column access is predicated, and the prefetched rows cost 48
// registers).  Here a slot of the warp's ring holds the dy row, the x row and (if given) the residual-gradient row, fetched
// by cp.async.bulk; x_hat and dy * gamma are computed once and kept (the registers the prefetch used to take), so the second
// pass is two FMAs per element, and the next row's mean / rstd are fetched one iteration ahead.
constexpr int kRingB = 3;

template <typename T, int kChunks, bool kExact>
__global__ void __launch_bounds__(kRingWarps * 32) ln_bwd_ring_kernel(const BwdParams p) {
  constexpr int V = Vec16<T>::kN;
  extern __shared__ __align__(128) uint8_t ring_raw[];
  __shared__ __align__(16) float gam_s[kChunks * 32 * V];
  __shared__ __align__(8) unsigned long long bars[kRingWarps][kRingB];
  __shared__ float red[kRingWarps][32 * V + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long warp0 = static_cast<long long>(blockIdx.x) * kRingWarps + warp;
  const long long nwarps = static_cast<long long>(gridDim.x) * kRingWarps;
  const uint32_t row_bytes = static_cast<uint32_t>(p.d) * sizeof(T);
  const int nsrc = p.dres ? 3 : 2;
  const uint32_t slot_bytes = nsrc * row_bytes;
  const uint32_t ring_s = static_cast<uint32_t>(__cvta_generic_to_shared(ring_raw)) + warp * kRingB * slot_bytes;
  const uint32_t bar_s = static_cast<uint32_t>(__cvta_generic_to_shared(&bars[warp][0]));
  auto fetch = [&](long long r, int slot) {
    const uint32_t dst = ring_s + slot * slot_bytes, bar = bar_s + 8 * slot;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(slot_bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(reinterpret_cast<const T*>(p.dy) + r * p.d), "r"(row_bytes), "r"(bar) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst + row_bytes), "l"(reinterpret_cast<const T*>(p.x) + r * p.d), "r"(row_bytes), "r"(bar) : "memory");
    if (p.dres)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dst + 2 * row_bytes), "l"(reinterpret_cast<const T*>(p.dres) + r * p.d), "r"(row_bytes), "r"(bar) : "memory");
  };
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kRingB; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s + 8 * s) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  pdl_wait();
  pdl_trigger();
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kRingB; ++s)
      if (warp0 + s * nwarps < p.rows) fetch(warp0 + s * nwarps, s);
  }
  for (int t = threadIdx.x; t < kChunks * 32 * V; t += kRingWarps * 32) gam_s[t] = t < p.d ? p.gamma[t] : 0.0f;
  __syncthreads();
  float acc_g[kChunks][V], acc_b[kChunks][V], acc_z[kChunks][V];
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
#pragma unroll
    for (int i = 0; i < V; ++i) acc_g[c][i] = acc_b[c][i] = acc_z[c][i] = 0.0f;
  }
  const float inv_d = 1.0f / p.d;
  float mean_n = 0.0f, rstd_n = 0.0f;
  if (warp0 < p.rows) { mean_n = p.mean[warp0]; rstd_n = p.rstd[warp0]; }
  int slot = 0;
  uint32_t parity = 0;
  for (long long row = warp0; row < p.rows; row += nwarps) {
    const float rstd = rstd_n, shift = -mean_n * rstd_n;
    if (row + nwarps < p.rows) { mean_n = p.mean[row + nwarps]; rstd_n = p.rstd[row + nwarps]; }
    bar_wait(bar_s + 8 * slot, parity);
    const uint32_t base = ring_s + slot * slot_bytes;
    float xh[kChunks][V], g[kChunks][V];
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int col = (c * 32 + lane) * V;
      uint4 rdy = make_uint4(0, 0, 0, 0), rx = make_uint4(0, 0, 0, 0);
      if (kExact || col < p.d) {
        const uint32_t a = base + col * static_cast<uint32_t>(sizeof(T));
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rdy.x), "=r"(rdy.y), "=r"(rdy.z), "=r"(rdy.w) : "r"(a) : "memory");
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rx.x), "=r"(rx.y), "=r"(rx.z), "=r"(rx.w) : "r"(a + row_bytes) : "memory");
      }
      float dyv[V], xv[V];
      Vec16<T>::unpack(rdy, dyv);
      Vec16<T>::unpack(rx, xv);
#pragma unroll
      for (int i = 0; i < V; i += 4) {
        const float4 g4 = *reinterpret_cast<const float4*>(gam_s + col + i);
        const float gm[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float h = (kExact || col < p.d) ? fmaf(xv[i + u], rstd, shift) : 0.0f;
          const float gg = dyv[i + u] * gm[u];
          xh[c][i + u] = h;
          g[c][i + u] = gg;
          acc_g[c][i + u] = fmaf(dyv[i + u], h, acc_g[c][i + u]);
          acc_b[c][i + u] += dyv[i + u];
          s1 += gg;
          s2 = fmaf(gg, h, s2);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const float A = -rstd * s2 * inv_d, B = -rstd * s1 * inv_d;
    T* dx_row = reinterpret_cast<T*>(p.dx) + row * p.d;
    T* dz_row = p.dz ? reinterpret_cast<T*>(p.dz) + row * p.d : nullptr;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int col = (c * 32 + lane) * V;
      if (kExact || col < p.d) {
        float o[V];
#pragma unroll
        for (int i = 0; i < V; ++i) o[i] = fmaf(xh[c][i], A, fmaf(g[c][i], rstd, B));
        if (p.dres) {
          uint4 rr;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rr.x), "=r"(rr.y), "=r"(rr.z), "=r"(rr.w)
                       : "r"(base + 2 * row_bytes + col * static_cast<uint32_t>(sizeof(T))) : "memory");
          float rv[V];
          Vec16<T>::unpack(rr, rv);
#pragma unroll
          for (int i = 0; i < V; ++i) o[i] += rv[i];
        }
        Vec16<T>::store(dx_row + col, o);
        if (p.dropout_thr16) {   // dropout_ld % 4 == 0 and col % 4 == 0: one hash per 4 consecutive elements
          const unsigned long long e4 = (static_cast<unsigned long long>(row) * p.dropout_ld + col) >> 2;
#pragma unroll
          for (int q = 0; q < V / 4; ++q) {
            const uint64_t bits = dropout_bits4(mix_seed(p.dropout_seed, p.seed_src), e4 + q);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              o[4 * q + i] = dropout_keep_lane(bits, i, p.dropout_thr16) ? o[4 * q + i] * p.dropout_scale : 0.0f;
          }
        }
        if (dz_row) Vec16<T>::store(dz_row + col, o);
#pragma unroll
        for (int i = 0; i < V; ++i) acc_z[c][i] += o[i];
      }
    }
    // the slot (incl. its residual-gradient row) has been consumed by every lane: refill it kRingB rows ahead
    __syncwarp();
    if (lane == 0 && row + kRingB * nwarps < p.rows) fetch(row + kRingB * nwarps, slot);
    if (++slot == kRingB) { slot = 0; parity ^= 1; }
  }
  // CTA-level reduction of the column partials through shared memory, then one atomic per column.
  for (int which = 0; which < 3; ++which) {
    float* out = which == 0 ? p.dgamma : (which == 1 ? p.dbeta : p.dbias);
    if (!out) continue;  // uniform
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < V; ++i)
        red[warp][lane * V + i] = which == 0 ? acc_g[c][i] : (which == 1 ? acc_b[c][i] : acc_z[c][i]);
      __syncthreads();
      for (int t = threadIdx.x; t < 32 * V; t += kRingWarps * 32) {
        float sacc = 0.0f;
#pragma unroll
        for (int w = 0; w < kRingWarps; ++w) sacc += red[w][t];
        const int col = c * 32 * V + t;
        if (col < p.d) atomicAdd(out + col, sacc);
      }
    }
  }
}

template <typename T, template <typename, int> class Launcher, typename P>
static int dispatch_chunks(const P& p, int d, cudaStream_t s) {
  constexpr int V = Vec16<T>::kN;
  const int chunks = (d + 32 * V - 1) / (32 * V);
  if (chunks <= 1) return Launcher<T, 1>::run(p, s);
  if (chunks <= 2) return Launcher<T, 2>::run(p, s);
  if (chunks <= 3) return Launcher<T, 3>::run(p, s);
  if (chunks <= 4) return Launcher<T, 4>::run(p, s);
  if (chunks <= 6) return Launcher<T, 6>::run(p, s);
  if (chunks <= 8) return Launcher<T, 8>::run(p, s);
  if (chunks <= 16) return Launcher<T, 16>::run(p, s);
  set_last_error("layernorm: d=%d too wide (max %d)", d, 16 * 32 * V);
  return TVT_EINVAL;
}

// One persistent wave: grid = SMs x resident CTAs per SM (queried per kernel), capped by the row count.
template <typename K>
static int grid_for(K kern, int threads, long long rows, int warps_per_cta) {
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, 0) != cudaSuccess || occ < 1) occ = 2;
  const long long want = (rows + warps_per_cta - 1) / warps_per_cta;
  const long long cap = static_cast<long long>(num_sms()) * occ;
  return static_cast<int>(want < cap ? (want > 0 ? want : 1) : cap);
}

template <typename T, int C>
struct FwdLauncher {
  static int run(const FwdParams& p, cudaStream_t s) {
    ln_fwd_kernel<T, C><<<grid_for(ln_fwd_kernel<T, C>, kWarps * 32, p.rows, kWarps), kWarps * 32, 0, s>>>(p);
    return check_launch("tvt_layernorm_fwd");
  }
};
template <typename T, int C>
struct FwdRingLauncher {
  template <bool kExact>
  static int launch(const FwdParams& p, cudaStream_t s) {
    auto kern = ln_fwd_ring_kernel<T, C, kExact>;
    const size_t bytes = static_cast<size_t>(kRingWarps) * kRing * 2 * p.d * sizeof(T);
    static bool attr_done = false;   // benign race: the attribute call is idempotent
    if (!attr_done) {
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024) != cudaSuccess) return check_launch("tvt_layernorm_fwd");
      attr_done = true;
    }
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kRingWarps * 32, bytes) != cudaSuccess || occ < 1) occ = 2;
    const long long want = ((p.rows + 1) / 2 + kRingWarps - 1) / kRingWarps;
    const long long cap = static_cast<long long>(num_sms()) * occ;
    launch_pdl(kern, static_cast<int>(want < cap ? want : cap), kRingWarps * 32, bytes, s, 1, p);
    return check_launch("tvt_layernorm_fwd");
  }
  static int run(const FwdParams& p, cudaStream_t s) {
    return p.d == C * 32 * Vec16<T>::kN ? launch<true>(p, s) : launch<false>(p, s);
  }
};
template <typename T, int C>
struct BwdRingLauncher {
  template <bool kExact>
  static int launch(const BwdParams& p, cudaStream_t s) {
    auto kern = ln_bwd_ring_kernel<T, C, kExact>;
    const size_t bytes = static_cast<size_t>(kRingWarps) * kRingB * (p.dres ? 3 : 2) * p.d * sizeof(T);
    static bool attr_done = false;   // benign race: the attribute call is idempotent
    if (!attr_done) {
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024) != cudaSuccess) return check_launch("tvt_layernorm_bwd");
      attr_done = true;
    }
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kRingWarps * 32, bytes) != cudaSuccess || occ < 1) occ = 2;
    // fewer, fatter CTAs than rows would allow: each ends with d atomics per output vector
    const long long want = ((p.rows + 3) / 4 + kRingWarps - 1) / kRingWarps;
    const long long cap = static_cast<long long>(num_sms()) * occ;
    launch_pdl(kern, static_cast<int>(want < cap ? (want > 0 ? want : 1) : cap), kRingWarps * 32, bytes, s, 1, p);
    return check_launch("tvt_layernorm_bwd");
  }
  static int run(const BwdParams& p, cudaStream_t s) {
    return p.d == C * 32 * Vec16<T>::kN ? launch<true>(p, s) : launch<false>(p, s);
  }
};
template <typename T, int C>
struct BwdLauncher {
  static int run(const BwdParams& p, cudaStream_t s) {
    // fewer, fatter CTAs: each ends with d atomics per output vector
    const int grid = grid_for(ln_bwd_kernel<T, C>, kBwdWarps * 32, (p.rows + 3) / 4, kBwdWarps);
    ln_bwd_kernel<T, C><<<grid, kBwdWarps * 32, 0, s>>>(p);
    return check_launch("tvt_layernorm_bwd");
  }
};

}  // namespace ln
}  // namespace tvt

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int tvt_layernorm_fwd(const tvt_layernorm_fwd_args* a, void* stream) {
  using namespace tvt;
  TVT_REQUIRE(a != nullptr, "tvt_layernorm_fwd: null args");
  TVT_REQUIRE(a->x && a->y && a->gamma && a->beta, "tvt_layernorm_fwd: null pointer");
  TVT_REQUIRE(a->rows >= 0 && a->d > 0 && a->d % 8 == 0, "tvt_layernorm_fwd: d must be a positive multiple of 8 (got %lld)", (long long)a->d);
  TVT_REQUIRE(a->dtype == TVT_BF16 || a->dtype == TVT_F32, "tvt_layernorm_fwd: bad dtype");
  TVT_REQUIRE(al16(a->x) && al16(a->y) && al16(a->cls) && al16(a->pre) && al16(a->gamma) && al16(a->beta), "tvt_layernorm_fwd: pointers must be 16-byte aligned");
  TVT_REQUIRE(a->seq_len >= 0, "tvt_layernorm_fwd: bad seq_len");
  if (a->seq_len > 0) {
    TVT_REQUIRE(a->cls && a->pe, "tvt_layernorm_fwd: embed mode needs cls and pe");
    TVT_REQUIRE(a->seq_len >= 2 && a->rows % a->seq_len == 0, "tvt_layernorm_fwd: rows must be a multiple of seq_len");
  }
  TVT_REQUIRE(a->dropout_p >= 0.0f && a->dropout_p < 1.0f, "tvt_layernorm_fwd: dropout_p must be in [0,1)");
  TVT_REQUIRE(a->dropout_p == 0.0f || a->seq_len > 0, "tvt_layernorm_fwd: dropout only applies in embed mode");
  if (a->y_seq != 0) {
    const int max_d = 4 * 32 * (a->dtype == TVT_F32 ? 4 : 8);
    TVT_REQUIRE(a->y_seq > 0 && a->rows % a->y_seq == 0 && a->y_pitch >= a->y_seq * a->d && a->y_pitch % 8 == 0,
                "tvt_layernorm_fwd: y_seq must divide rows and y_pitch (elements, multiple of 8) must hold y_seq rows");
    TVT_REQUIRE(a->seq_len == 0 && a->d <= max_d, "tvt_layernorm_fwd: the blocked output layout (y_seq) is built for plain rows of d <= %d", max_d);
  }
  if (a->rows == 0) return TVT_OK;
  int rc = require_sm100();
  if (rc != TVT_OK) return rc;
  ln::FwdParams p{};
  static const int ln_reverse = [] { const char* e = getenv("TVT_LN_REVERSE"); return e ? atoi(e) : 1; }();
  p.reverse = ln_reverse;
  p.x = a->x; p.cls = a->cls; p.pe = a->pe; p.gamma = a->gamma; p.beta = a->beta; p.y = a->y; p.pre = a->pre;
  p.mean = a->mean; p.rstd = a->rstd; p.rows = a->rows; p.d = (int)a->d; p.S = (int)a->seq_len; p.eps = a->eps;
  p.y_seq = (int)a->y_seq; p.y_pitch = a->y_pitch;
  if (a->dropout_p > 0.0f) {
    p.dropout_thr16 = (unsigned)(a->dropout_p * 65536.0f + 0.5f);
    p.dropout_scale = 65536.0f / (65536.0f - (float)p.dropout_thr16);
    p.dropout_seed = a->dropout_seed;
    p.seed_src = seed_source();
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // plain rows of up to 4 chunks (d <= 1024 bf16 / 512 fp32): bulk-copy ring kernel; embed mode and wider rows: register-prefetch kernel
  const int vec = a->dtype == TVT_F32 ? 4 : 8;
  if (a->seq_len == 0 && p.d <= 4 * 32 * vec)
    return a->dtype == TVT_F32 ? ln::dispatch_chunks<float, ln::FwdRingLauncher>(p, p.d, s)
                               : ln::dispatch_chunks<__nv_bfloat16, ln::FwdRingLauncher>(p, p.d, s);
  return a->dtype == TVT_F32 ? ln::dispatch_chunks<float, ln::FwdLauncher>(p, p.d, s)
                             : ln::dispatch_chunks<__nv_bfloat16, ln::FwdLauncher>(p, p.d, s);
}

extern "C" int tvt_layernorm_bwd(const tvt_layernorm_bwd_args* a, void* stream) {
  using namespace tvt;
  TVT_REQUIRE(a != nullptr, "tvt_layernorm_bwd: null args");
  TVT_REQUIRE(a->dy && a->x && a->mean && a->rstd && a->gamma, "tvt_layernorm_bwd: null pointer");
  TVT_REQUIRE(a->rows >= 0 && a->d > 0 && a->d % 8 == 0, "tvt_layernorm_bwd: d must be a positive multiple of 8");
  TVT_REQUIRE(a->dtype == TVT_BF16 || a->dtype == TVT_F32, "tvt_layernorm_bwd: bad dtype");
  TVT_REQUIRE(al16(a->dy) && al16(a->x) && al16(a->dx) && al16(a->dz) && al16(a->dfeat) && al16(a->dcls) && al16(a->gamma),
              "tvt_layernorm_bwd: pointers must be 16-byte aligned");
  if (a->seq_len > 0) {
    TVT_REQUIRE(a->seq_len >= 2 && a->rows % a->seq_len == 0, "tvt_layernorm_bwd: rows must be a multiple of seq_len");
    TVT_REQUIRE(!a->dz, "tvt_layernorm_bwd: dz is implied by dfeat/dcls in embed mode");
  } else {
    TVT_REQUIRE(a->dx, "tvt_layernorm_bwd: dx required");
    TVT_REQUIRE(!a->dfeat && !a->dcls, "tvt_layernorm_bwd: dfeat/dcls need seq_len > 0");
  }
  TVT_REQUIRE(a->dropout_p >= 0.0f && a->dropout_p < 1.0f, "tvt_layernorm_bwd: dropout_p must be in [0,1)");
  if (a->rows == 0) return TVT_OK;
  int rc = require_sm100();
  if (rc != TVT_OK) return rc;
  ln::BwdParams p{};
  p.dy = a->dy; p.x = a->x; p.mean = a->mean; p.rstd = a->rstd; p.gamma = a->gamma;
  p.dx = a->dx; p.dz = a->dz; p.dfeat = a->dfeat; p.dcls = a->dcls;
  p.dgamma = a->dgamma; p.dbeta = a->dbeta; p.dbias = a->dbias;
  p.rows = a->rows; p.d = (int)a->d; p.S = (int)a->seq_len;
  if (a->dropout_p > 0.0f) {
    p.dropout_thr16 = (unsigned)(a->dropout_p * 65536.0f + 0.5f);
    p.dropout_scale = 65536.0f / (65536.0f - (float)p.dropout_thr16);
    p.dropout_seed = a->dropout_seed;
    p.seed_src = seed_source();
  }
  p.dropout_ld = a->d;
  p.dres = a->dres;
  TVT_REQUIRE(!(a->dres && (a->dz || a->seq_len > 0 || a->dropout_p > 0.0f)), "tvt_layernorm_bwd: dres cannot be combined with dz / embed mode / dropout");
  TVT_REQUIRE(al16(a->dres), "tvt_layernorm_bwd: dres must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int vec = a->dtype == TVT_F32 ? 4 : 8;
  if (a->seq_len == 0 && p.d <= 4 * 32 * vec)
    return a->dtype == TVT_F32 ? ln::dispatch_chunks<float, ln::BwdRingLauncher>(p, p.d, s)
                               : ln::dispatch_chunks<__nv_bfloat16, ln::BwdRingLauncher>(p, p.d, s);
  return a->dtype == TVT_F32 ? ln::dispatch_chunks<float, ln::BwdLauncher>(p, p.d, s)
                             : ln::dispatch_chunks<__nv_bfloat16, ln::BwdLauncher>(p, p.d, s);
}
