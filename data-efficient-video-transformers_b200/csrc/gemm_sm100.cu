// tcgen05 / TMEM / TMA GEMM for sm_100a with a fused epilogue (see include/tvt.h: tvt_gemm).
//
// Persistent, warp-specialised kernel, one CTA per SM:
//   warp 0 (one lane)  TMA producer: global -> 128B-swizzled smem ring, mbarrier complete_tx
//   warp 1 (one lane)  MMA issuer:   tcgen05.mma 128 x BN x 16, fp32 accumulators in tensor memory,
//                                    tcgen05.commit releases smem stages / publishes accumulators
//   warp 2             TMEM allocator (alloc at start, dealloc at end)
//   warps 4..7         epilogue: tcgen05.ld 32 lanes x 32 columns -> registers -> fused math -> global
// Two accumulator stages in TMEM let the epilogue of tile i overlap the MMAs of tile i+1.
//
// Work item = (split, m_block, n_block); split-K partials are combined with red.global.add.f32.
// Operand storage is selected per operand (K-major or MN-major) through the UMMA descriptors, so the
// forward (X W^T), dgrad (dY W) and wgrad (dY^T X) products all read the tensors as torch stores them.
#include <cstdlib>
#include <cuda.h>
#include <mutex>

#include "tvt_common.cuh"
#include "tvt_ptx.cuh"

namespace tvt {
namespace gemm {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kEpiWarps = 8;                 // two warps per TMEM lane quadrant, each owning half the tile columns
constexpr int kThreads = 32 * kEpiWarps + 96;  // warps 0..7 epilogue (warp % 4 = TMEM lane quadrant), 8 TMA producer, 9 MMA issuer + TMEM owner, 10 side-operand producer
constexpr int kSideSlotBytes = 2 * BM * 128;  // one ring slot: a [128 rows x 64 bf16] box for each column half of the tile
constexpr int kStagePitch = 64;              // bytes per staged row: 32 bf16; 16 B chunks XOR-swizzled by (row >> 1) & 3
constexpr int kBiasSlab = 512;               // bytes: the bias of the (up to) 128 columns one epilogue warp owns
constexpr int kEpiStageBytes = kBiasSlab + 32 * kStagePitch;  // per epilogue warp: the bias slab, then the staging rows (none in the fast kernels)
constexpr int kSmemLimit = 227 * 1024;

struct Params {
  int M, N, K, splits;
  int act;
  float alpha;
  const float* bias;
  const void* residual; int residual_f32; long long ld_residual;
  const void* relu_mask; int mask_f32; long long ld_mask;
  const void* gelu_gate; int gate_f32; long long ld_gate;
  float dropout_scale; unsigned dropout_thr16; unsigned long long dropout_seed; const unsigned long long* seed_src;
  unsigned drop_rk[kDropoutRounds];   // per-round keys of the dropout hash (host-computed: they are launch constants)
  void* out_preact; int preact_f32; long long ld_preact;
  // LayerNorm folded into the epilogue (see tvt_gemm_args): raw (sum, sum of squares) rows, 1 / ln_dim, eps
  const float* ln_in_stats; const float* ln_in_c; const float* ln_res_stats; const float* ln_res_gamma; const float* ln_res_beta;
  float* stats_out; float ln_inv_d, ln_eps; int ln_slots;   // partial (sum, sum of squares) slots per row: 2 per 256-column block of the producer
  float* out_f32; long long ld_f32; int atomic_out;
  float* a_rowsum;          // kEpiAtomicSum: [M] += sum over k of A[m, k]
  __nv_bfloat16* out_bf16; __nv_bfloat16* out_bf16_lo; long long ld_bf16;
  unsigned mn_lbo, mn_sbo;  // MN-major descriptor strides (bring-up knob, see tvt_debug_set_mn_desc)
  int dbg;                  // bring-up knob: low 2 bits 1 = epilogue drains TMEM only, 2 = no global stores; 4 = no main loop; 8 = no epilogue
};

static unsigned g_mn_lbo = BK * 128, g_mn_sbo = 1024;
static int g_dbg = [] { const char* e = getenv("TVT_GEMM_DBG"); return e ? atoi(e) : 0; }();   // bring-up bits, see Params::dbg (32 = release.cluster accumulator hand-back)
static int g_pair = [] { const char* e = getenv("TVT_GEMM_PAIR"); return e ? atoi(e) : 1; }();   // bring-up knob: 0 = never use the CTA-pair kernels
static long long g_fast_fallbacks = 0;   // fast-path launches that had no exact-stage kernel (see tvt_gemm)

template <int BN, int kPlanes, bool kSide, bool kFastEpi, bool kCta2 = false, int kSlabs = 1, bool kRowsum = false>
struct Cfg {
  static constexpr int kAPlane = BM * BK * 2;
  static constexpr int kBRows = kCta2 ? BN / 2 : BN;   // a CTA pair splits the B tile between its two CTAs
  static constexpr int kBPlane = kBRows * BK * 2;
  static constexpr int kStageBytes = kPlanes * (kAPlane + kBPlane);
  static constexpr int kEpiWarpBytes = kFastEpi ? kSlabs * kBiasSlab : kEpiStageBytes;   // fast kernels: bias (+ c | gamma, beta) slabs
  static constexpr int kEpiBytes = kEpiWarps * kEpiWarpBytes;
  static constexpr int kSideSlots = BN / 128;  // one per two 32-column chunk steps of the epilogue warps
  static constexpr int kSideBytes = kSide ? kSideSlots * kSideSlotBytes : 0;
  static constexpr int kStages = (kSmemLimit - 2048 - kEpiBytes - kSideBytes) / kStageBytes;
  // kRowsum (kEpiAtomicSum): 16 extra accumulator columns per stage hold A . 1 (the bias gradient of a wgrad GEMM); a 256-wide
  // tile then keeps ONE accumulator stage (512 TMEM columns in all) - the wgrad tiles it serves run hundreds of k-blocks per
  // epilogue, so the lost overlap is noise
  static constexpr int kAccStages = (kRowsum && BN == 256) ? 1 : 2;
  static constexpr int kSumCol = kAccStages * BN;     // first column of the row-sum accumulators (16 per stage)
  static constexpr int kTmemCols = kRowsum ? 512 : kAccStages * BN;  // 512 or 256: powers of two
  static constexpr int kSmemBytes = kStages * kStageBytes + kSideBytes + kEpiBytes + 1024 /*align slack*/ + 512 /*barriers*/;
  static_assert(kStages >= 2, "need at least a double buffer");
  static_assert(!kSide || kPlanes == 2 || kStages >= 3, "side-ring kernels keep three operand stages");
  static_assert(kSmemBytes <= kSmemLimit, "smem budget");
};

__device__ __forceinline__ void load8(const void* base, int is_f32, long long off, float (&v)[8]) {
  if (is_f32) {
    const float* p = reinterpret_cast<const float*>(base) + off;
    float4 a = *reinterpret_cast<const float4*>(p);
    float4 b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    Vec16<__nv_bfloat16>::load(reinterpret_cast<const __nv_bfloat16*>(base) + off, v);
  }
}
__device__ __forceinline__ void store8(void* base, int is_f32, long long off, const float (&v)[8]) {
  if (is_f32) {
    float* p = reinterpret_cast<float*>(base) + off;
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    Vec16<__nv_bfloat16>::store(reinterpret_cast<__nv_bfloat16*>(base) + off, v);
  }
}

// Fused epilogue on one 32-column chunk of one output row (v = raw accumulators in, final values out).
// ng = number of valid 8-column groups in the chunk (4 except in the last column block of a ragged N).
// Every optional stage is ONE uniform branch around a straight-line block, which keeps the epilogue small
// enough for the instruction cache (the first version, branching per 8-column group, was I$-bound).
//
// kEpi selects how much of this is compiled in.  The loop around it is instruction-cache bound when everything
// is present (an 80 KB kernel against a 32 KB L1.5 I$), so the hot combinations get their own small kernels:
//   kEpiFast+st: alpha, bias and the stages in st (relu, bf16 relu mask, dropout, bf16 residual) -> bf16 output
//   kEpiAtomic : split-K fp32 red.add only
//   kEpiGeneric: every stage (fp32 operands, pre-activation store, gelu, hi/lo planes, fp32 output ...)
// kEpiFast + a stage mask (kStRelu ...) is a kernel with exactly those stages compiled in, unconditionally.
//   kEpiAtomicSum: kEpiAtomic + the row sums of the A operand over the contraction (tvt_gemm_args.a_rowsum): one extra N = 16
//                MMA per k-step against a constant tile of ones, on every num_n-th k-block of a tile
enum { kEpiGeneric = 0, kEpiAtomic = 2, kEpiAtomicSum = 3, kEpiFast = 16 };
__host__ __device__ constexpr bool is_atomic(int kEpi) { return kEpi == kEpiAtomic || kEpi == kEpiAtomicSum; }
enum { kStRelu = 1, kStMask = 2, kStDrop = 4, kStRes = 8 };
// LayerNorm-folded inference epilogues (tvt_gemm_args.ln_*): kStLnIn = the A operand is a PRE-norm tensor (per-row rstd / mean
// and the per-column c vector rebuild LN(y) W^T); kStLnRes = the residual is LN(side operand), recomputed; kStStats = accumulate
// the output row's (sum, sum of squares) for the next LayerNorm.  Bits above kSideLdg.
enum { kStLnIn = 64, kStLnRes = 128, kStStats = 256 };
__host__ __device__ constexpr int epi_slabs(int kEpi) {
  return kEpi >= 16 ? ((((kEpi - 16) & kStLnRes) != 0) ? 3 : ((((kEpi - 16) & kStLnIn) != 0) ? 2 : 1)) : 1;
}
__host__ __device__ constexpr bool is_fast(int kEpi) { return kEpi >= kEpiFast; }
__host__ __device__ constexpr bool has_stage(int kEpi, int st) { return kEpi >= kEpiFast && ((kEpi - kEpiFast) & st) != 0; }
// Fast kernels with a bf16 side operand (relu mask or residual) fetch it one of two ways:
//   ring (default): TMA streams it through a shared-memory ring about one tile-epilogue ahead of its use.  The
//     ring costs one of the four operand stages, which a short-K (epilogue-bound) problem does not miss;
//   kSideLdg: each lane loads its row's 64 B one chunk ahead with 256-bit loads.  Slow per byte, but free of shared
//     memory: for long-K problems, where the epilogue has slack and the main loop wants all four stages.
enum { kSideLdg = 32 };
__host__ __device__ constexpr bool side_ldg(int kEpi) { return kEpi >= kEpiFast && ((kEpi - kEpiFast) & kSideLdg) != 0 && has_stage(kEpi, kStMask | kStRes); }
__host__ __device__ constexpr bool has_side(int kEpi) { return has_stage(kEpi, kStMask | kStRes) && !side_ldg(kEpi); }

__device__ __forceinline__ void epilogue_atomic(const Params& p, long long row, int col0, int ng, const float (&v)[32]) {
  float* dst = p.out_f32 + row * p.ld_f32 + col0;
#pragma unroll
  for (int q = 0; q < 8; ++q)
    if (q < 2 * ng)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * q), "f"(v[4 * q] * p.alpha),
                   "f"(v[4 * q + 1] * p.alpha), "f"(v[4 * q + 2] * p.alpha), "f"(v[4 * q + 3] * p.alpha)
                   : "memory");
}

template <int kEpi>
__device__ __forceinline__ void epilogue_stages(const Params& p, long long row, int col0, int ng, float (&v)[32],
                                                const uint32_t (&mask_pk)[16], const uint32_t (&res_pk)[16], bool pre, uint32_t bias_s) {
  if (kEpi == kEpiGeneric && p.atomic_out) {
    epilogue_atomic(p, row, col0, ng, v);
    return;
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] *= p.alpha;
  if (p.bias) {   // broadcast reads of the warp's bias slab (global loads here serialised on the L2 latency)
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const uint4 b = lds128(bias_s + 16 * q);
      v[4 * q] += __uint_as_float(b.x); v[4 * q + 1] += __uint_as_float(b.y);
      v[4 * q + 2] += __uint_as_float(b.z); v[4 * q + 3] += __uint_as_float(b.w);
    }
  }
  if constexpr (kEpi == kEpiGeneric) {
    if (p.out_preact) {
#pragma unroll
      for (int g = 0; g < 4; ++g)
        if (g < ng) store8(p.out_preact, p.preact_f32, row * p.ld_preact + col0 + 8 * g, *reinterpret_cast<float(*)[8]>(&v[8 * g]));
    }
  }
  if (p.act == TVT_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
  } else if (kEpi == kEpiGeneric && p.act == TVT_ACT_GELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gelu_f(v[i]);
  }
  if (p.relu_mask) {
    if (pre && !p.mask_f32) {   // bf16 mask row fetched coalesced by the caller: sign/zero test on the raw halves
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float2 m = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&mask_pk[i]));
        v[2 * i] = m.x > 0.0f ? v[2 * i] : 0.0f;
        v[2 * i + 1] = m.y > 0.0f ? v[2 * i + 1] : 0.0f;
      }
    } else if constexpr (kEpi == kEpiGeneric) {
#pragma unroll
      for (int g = 0; g < 4; ++g)
        if (g < ng) {
          float m[8];
          load8(p.relu_mask, p.mask_f32, row * p.ld_mask + col0 + 8 * g, m);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[8 * g + i] = m[i] > 0.0f ? v[8 * g + i] : 0.0f;
        }
    }
  }
  if constexpr (kEpi == kEpiGeneric) {
    if (p.gelu_gate) {
#pragma unroll
      for (int g = 0; g < 4; ++g)
        if (g < ng) {
          float m[8];
          load8(p.gelu_gate, p.gate_f32, row * p.ld_gate + col0 + 8 * g, m);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[8 * g + i] *= gelu_grad_f(m[i]);
        }
    }
  }
  if (p.dropout_thr16) {
    const unsigned long long e4 = (static_cast<unsigned long long>(row) * p.N + col0) >> 2;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const uint64_t bits = dropout_bits4(mix_seed(p.dropout_seed, p.seed_src), e4 + q);
#pragma unroll
      for (int i = 0; i < 4; ++i) v[4 * q + i] = dropout_keep_lane(bits, i, p.dropout_thr16) ? v[4 * q + i] * p.dropout_scale : 0.0f;
    }
  }
  if (p.residual) {
    if (pre && !p.residual_f32) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float2 m = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&res_pk[i]));
        v[2 * i] += m.x;
        v[2 * i + 1] += m.y;
      }
    } else if constexpr (kEpi == kEpiGeneric) {
#pragma unroll
      for (int g = 0; g < 4; ++g)
        if (g < ng) {
          float m[8];
          load8(p.residual, p.residual_f32, row * p.ld_residual + col0 + 8 * g, m);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[8 * g + i] += m[i];
        }
    }
  }
  if constexpr (kEpi != kEpiGeneric) return;
  if (p.out_f32) {
#pragma unroll
    for (int g = 0; g < 4; ++g)
      if (g < ng) store8(p.out_f32, 1, row * p.ld_f32 + col0 + 8 * g, *reinterpret_cast<float(*)[8]>(&v[8 * g]));
  }
  if (p.out_bf16 && p.out_bf16_lo) {  // fp32-parity mode: hi / lo planes, written directly
#pragma unroll
    for (int g = 0; g < 4; ++g)
      if (g < ng) {
        float hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { hi[i] = __bfloat162float(__float2bfloat16_rn(v[8 * g + i])); lo[i] = v[8 * g + i] - hi[i]; }
        Vec16<__nv_bfloat16>::store(p.out_bf16 + row * p.ld_bf16 + col0 + 8 * g, hi);
        Vec16<__nv_bfloat16>::store(p.out_bf16_lo + row * p.ld_bf16 + col0 + 8 * g, lo);
      }
  }
  // the plain bf16 output (the hot path) is staged through shared memory by the caller
}

// 256-bit global accesses: one 32 B sector per lane, so "thread == row" epilogue traffic needs no transposition
// through shared memory to be sector-efficient.
__device__ __forceinline__ void ldg256(const void* ptr, uint32_t* r) {
  asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(ptr));
}
__device__ __forceinline__ void stg256(void* ptr, const uint32_t* r) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
               "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}

// The hot-path epilogue (kEpiFast) on one 32-column chunk of one row: v = fma(acc, scale, bias'), relu, bf16
// relu mask, dropout, bf16 residual, packed to bf16.  scale = alpha * dropout_scale and bias' = bias *
// dropout_scale are folded by the caller (every stage before the residual is positively homogeneous), side[]
// holds the row's 32 mask or residual values.  About 320 instructions with every stage on.
// Per-row scalars of the LayerNorm-folded stages (thread == row): rs / rb = rstd and -mean * rstd of the A operand's row
// (kStLnIn), ls / lb the same for the residual's row (kStLnRes); sum / sq accumulate the output row (kStStats).
struct RowLn { float rs, rb, ls, lb, sum, sq; };

template <int kEpi>
__device__ __forceinline__ void epilogue_fast(const Params& p, float scale, uint32_t bias_s, uint32_t e4_lo, uint32_t e4_hi,
                                              const uint32_t (&r)[32], const uint32_t (&side)[16], uint32_t (&out)[16], RowLn& ln) {
  float v[32];
  if constexpr (has_stage(kEpi, kStLnIn)) {
    const float sc = scale * ln.rs;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const uint4 b = lds128(bias_s + 16 * q);               // b'
      const uint4 c = lds128(bias_s + kBiasSlab + 16 * q);   // c = rowsum(W diag(gamma))
      v[4 * q] = fmaf(__uint_as_float(r[4 * q]), sc, fmaf(ln.rb, __uint_as_float(c.x), __uint_as_float(b.x)));
      v[4 * q + 1] = fmaf(__uint_as_float(r[4 * q + 1]), sc, fmaf(ln.rb, __uint_as_float(c.y), __uint_as_float(b.y)));
      v[4 * q + 2] = fmaf(__uint_as_float(r[4 * q + 2]), sc, fmaf(ln.rb, __uint_as_float(c.z), __uint_as_float(b.z)));
      v[4 * q + 3] = fmaf(__uint_as_float(r[4 * q + 3]), sc, fmaf(ln.rb, __uint_as_float(c.w), __uint_as_float(b.w)));
    }
  } else {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const uint4 b = lds128(bias_s + 16 * q);   // broadcast read of the warp's bias slab
      v[4 * q] = fmaf(__uint_as_float(r[4 * q]), scale, __uint_as_float(b.x));
      v[4 * q + 1] = fmaf(__uint_as_float(r[4 * q + 1]), scale, __uint_as_float(b.y));
      v[4 * q + 2] = fmaf(__uint_as_float(r[4 * q + 2]), scale, __uint_as_float(b.z));
      v[4 * q + 3] = fmaf(__uint_as_float(r[4 * q + 3]), scale, __uint_as_float(b.w));
    }
  }
  if constexpr (has_stage(kEpi, kStRelu)) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
  }
  if constexpr (has_stage(kEpi, kStMask)) {   // keep where the bf16 mask value is > 0: integer tests on the packed halves
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      v[2 * i] = static_cast<int>(side[i] << 16) > 0 ? v[2 * i] : 0.0f;
      v[2 * i + 1] = static_cast<int>(side[i]) >= 0x10000 ? v[2 * i + 1] : 0.0f;
    }
  }
  if constexpr (has_stage(kEpi, kStDrop)) {
    const uint32_t thr_hi = p.dropout_thr16 << 16;   // (w >> 16) >= thr  <=>  w >= thr << 16
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      uint32_t lo, hi;
      dropout_words(p.drop_rk, e4_lo + q, e4_hi, lo, hi);
      v[4 * q] = (lo << 16) >= thr_hi ? v[4 * q] : 0.0f;
      v[4 * q + 1] = lo >= thr_hi ? v[4 * q + 1] : 0.0f;
      v[4 * q + 2] = (hi << 16) >= thr_hi ? v[4 * q + 2] : 0.0f;
      v[4 * q + 3] = hi >= thr_hi ? v[4 * q + 3] : 0.0f;
    }
  }
  if constexpr (has_stage(kEpi, kStRes)) {
    if constexpr (has_stage(kEpi, kStLnRes)) {   // residual = LN(side row): ((y - mean) * rstd) * gamma + beta
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint4 g = lds128(bias_s + kBiasSlab + 16 * q), b = lds128(bias_s + 2 * kBiasSlab + 16 * q);
        const float y0 = __uint_as_float(side[2 * q] << 16), y1 = __uint_as_float(side[2 * q] & 0xFFFF0000u);
        const float y2 = __uint_as_float(side[2 * q + 1] << 16), y3 = __uint_as_float(side[2 * q + 1] & 0xFFFF0000u);
        v[4 * q] += fmaf(fmaf(y0, ln.ls, ln.lb), __uint_as_float(g.x), __uint_as_float(b.x));
        v[4 * q + 1] += fmaf(fmaf(y1, ln.ls, ln.lb), __uint_as_float(g.y), __uint_as_float(b.y));
        v[4 * q + 2] += fmaf(fmaf(y2, ln.ls, ln.lb), __uint_as_float(g.z), __uint_as_float(b.z));
        v[4 * q + 3] += fmaf(fmaf(y3, ln.ls, ln.lb), __uint_as_float(g.w), __uint_as_float(b.w));
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        v[2 * i] += __uint_as_float(side[i] << 16);
        v[2 * i + 1] += __uint_as_float(side[i] & 0xFFFF0000u);
      }
    }
  }
  if constexpr (has_stage(kEpi, kStStats)) {   // row statistics from the fp32 values (the bf16 rounding of the stored row is zero-mean noise)
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      ln.sum += v[i] + v[i + 1];
      ln.sq = fmaf(v[i], v[i], fmaf(v[i + 1], v[i + 1], ln.sq));
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) out[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
}

template <int kEpi>
__device__ __forceinline__ void epilogue32(const Params& p, long long row, int col0, int ng, float (&v)[32],
                                           const uint32_t (&mask_pk)[16], const uint32_t (&res_pk)[16], bool pre, uint32_t bias_s) {
  if constexpr (is_atomic(kEpi)) epilogue_atomic(p, row, col0, ng, v);
  else epilogue_stages<kEpi>(p, row, col0, ng, v, mask_pk, res_pk, pre, bias_s);
}

// Warp-cooperative fetch of a [32 rows x 32 bf16] block of a row-major matrix, in two halves so the global
// loads of the NEXT chunk can be in flight while the current chunk is processed:
//   fetch_issue: coalesced global reads (4 lanes x 16 B per row, 8 rows per pass) into registers;
//   fetch_land:  transpose to "one row per lane" through the warp's swizzled staging buffer.
// Rows >= M / columns >= N read as zero.
__device__ __forceinline__ void fetch_issue(const void* base, long long ld, long long row0, int col0, int M, int N, int lane,
                                            uint4 (&t)[4]) {
  const int seg = lane & 3;
  const int gcol = col0 + seg * 8;
  const __nv_bfloat16* gptr = reinterpret_cast<const __nv_bfloat16*>(base) + (row0 + (lane >> 2)) * ld + gcol;
#pragma unroll
  for (int ps = 0; ps < 4; ++ps) {
    const int rr = ps * 8 + (lane >> 2);
    t[ps] = make_uint4(0, 0, 0, 0);
    if (row0 + rr < M && gcol < N) t[ps] = __ldg(reinterpret_cast<const uint4*>(gptr + static_cast<long long>(ps) * 8 * ld));
  }
}
__device__ __forceinline__ void fetch_land(const uint4 (&t)[4], uint32_t stage_addr, int lane, uint32_t (&out)[16]) {
  const int seg = lane & 3;
#pragma unroll
  for (int ps = 0; ps < 4; ++ps) {
    const int rr = ps * 8 + (lane >> 2);
    sts128(stage_addr + rr * kStagePitch + ((seg ^ ((rr >> 1) & 3)) << 4), t[ps].x, t[ps].y, t[ps].z, t[ps].w);
  }
  __syncwarp();
  const int sw = (lane >> 1) & 3;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const uint4 q = lds128(stage_addr + lane * kStagePitch + ((g ^ sw) << 4));
    out[4 * g] = q.x; out[4 * g + 1] = q.y; out[4 * g + 2] = q.z; out[4 * g + 3] = q.w;
  }
  __syncwarp();
}

// kCta2: the CTA pair of a 2-cluster computes one [256 x BN] tile with cta_group::2 MMAs issued by the leader (rank
// 0).  Each CTA loads its own 128 rows of A and BN/2 rows of B (a third less shared-memory fill and L2 traffic per
// flop, so more stages fit), accumulates its 128 rows in its own TMEM and runs its own epilogue.  Cross-CTA
// signalling: both CTAs' TMA loads count on the leader's full barrier; the leader's commits multicast to both
// CTAs' empty / accumulator-full barriers; the peer's epilogue warps arrive remotely on the leader's
// accumulator-empty barrier.
template <int BN, bool kAMN, bool kBMN, int kPlanes, int kEpi, bool kCta2>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAlo,
            const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBlo,
            const __grid_constant__ CUtensorMap tmSide, const Params p) {
  constexpr bool kSide = has_side(kEpi);
  constexpr bool kRowsum = kEpi == kEpiAtomicSum;
  using C = Cfg<BN, kPlanes, kSide, is_fast(kEpi), kCta2, epi_slabs(kEpi), kRowsum>;
  constexpr int kCtas = kCta2 ? 2 : 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* side_ring = smem + C::kStages * C::kStageBytes;   // 1024-aligned: SWIZZLE_64B boxes need 512
  uint8_t* epi_stage = side_ring + C::kSideBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_stage + C::kEpiBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::kStages;
  uint64_t* tfull_bar = bars + 2 * C::kStages;
  uint64_t* tempty_bar = tfull_bar + C::kAccStages;
  uint64_t* sfull_bar = tempty_bar + C::kAccStages;          // side ring: slot filled by TMA
  uint64_t* sempty_bar = sfull_bar + C::kSideSlots;          // slot read by all epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sempty_bar + C::kSideSlots);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = kCta2 ? cluster_ctarank() : 0;
  const int w0 = kCta2 ? blockIdx.x >> 1 : blockIdx.x;        // first tile and tile stride of this CTA (pair)
  const int wstep = kCta2 ? gridDim.x >> 1 : gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int s = 0; s < C::kAccStages; ++s) {
      mbar_init(smem_u32(&tfull_bar[s]), 1);
      mbar_init(smem_u32(&tempty_bar[s]), kEpiWarps * kCtas);  // one arrive per epilogue warp (of both CTAs of a pair)
    }
    for (int s = 0; s < C::kSideSlots; ++s) {
      mbar_init(smem_u32(&sfull_bar[s]), 1);
      mbar_init(smem_u32(&sempty_bar[s]), kEpiWarps);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if constexpr (kRowsum) {   // the B operand of the row-sum MMAs: 16 K-major rows of bf16 ones (8 per CTA of a pair), in the unused epilogue staging
    static_assert(C::kEpiBytes >= 2048 && kPlanes == 1, "row sums: single-plane kernels, 2 KB of ones");
    for (int i = threadIdx.x; i < 512; i += kThreads) reinterpret_cast<uint32_t*>(epi_stage)[i] = 0x3F803F80u;
    fence_proxy_async_smem();
  }
  if (warp == kEpiWarps + 1) {
    if constexpr (kCta2) {
      tmem_alloc_cta2(smem_u32(tmem_slot), C::kTmemCols);
      tmem_relinquish_cta2();
    } else {
      tmem_alloc(smem_u32(tmem_slot), C::kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (kCta2) cluster_sync_all(); else __syncthreads();   // barriers of both CTAs initialised before any remote use
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // the prologue above overlapped the previous kernel's tail; its results are visible from here on
  pdl_trigger();

  const int num_m = (p.M + BM * kCtas - 1) / (BM * kCtas);
  const int num_n = (p.N + BN - 1) / BN;
  const int total = num_m * num_n * p.splits;
  const int kb_total = (p.K + BK - 1) / BK;

  if (warp == kEpiWarps) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = w0; w < total; w += wstep) {
        const int n_blk = w % num_n;
        const int t = w / num_n;
        const int m_row = ((t % num_m) * kCtas + cta_rank) * BM;            // first A row of this CTA
        const int n_row = n_blk * BN + cta_rank * C::kBRows;                // first B row of this CTA
        const int split = t / num_m;
        const int kb0 = static_cast<int>(static_cast<long long>(split) * kb_total / p.splits);
        const int kb1 = static_cast<int>(static_cast<long long>(split + 1) * kb_total / p.splits);
        for (int kb = kb0; kb < kb1; ++kb) {
          if (p.dbg & 4) break;
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          // pair: the leader's barrier counts the bytes of both CTAs' loads
          const uint32_t fb = kCta2 ? mapa_u32(smem_u32(&full_bar[stage]), 0) : smem_u32(&full_bar[stage]);
          if (cta_rank == 0) mbar_arrive_expect_tx(smem_u32(&full_bar[stage]), C::kStageBytes * kCtas);
          auto tma = [&](uint32_t dst, const CUtensorMap* m, int c0, int c1) {
            if constexpr (kCta2) tma_load_2d_cta2(dst, m, fb, c0, c1); else tma_load_2d(dst, m, fb, c0, c1);
          };
          uint8_t* st = smem + stage * C::kStageBytes;
#pragma unroll
          for (int pl = 0; pl < kPlanes; ++pl) {
            const CUtensorMap* ma = pl == 0 ? &tmA : &tmAlo;
            const CUtensorMap* mb = pl == 0 ? &tmB : &tmBlo;
            const uint32_t sa = smem_u32(st + pl * C::kAPlane);
            const uint32_t sb = smem_u32(st + kPlanes * C::kAPlane + pl * C::kBPlane);
            if constexpr (!kAMN) {
              tma(sa, ma, kb * BK, m_row);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j) tma(sa + j * (BK * 128), ma, m_row + j * 64, kb * BK);
            }
            if constexpr (!kBMN) {
              tma(sb, mb, kb * BK, n_row);
            } else {
#pragma unroll
              for (int j = 0; j < C::kBRows / 64; ++j) tma(sb + j * (BK * 128), mb, n_row + j * 64, kb * BK);
            }
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == kEpiWarps + 1) {
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM * kCtas, BN, kAMN, kBMN);
      auto commit = [&](uint64_t* bar) {
        if constexpr (kCta2) tc_commit_cta2(smem_u32(bar), 3); else tc_commit(smem_u32(bar));
      };
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int w = w0; w < total; w += wstep) {
        const int split = (w / num_n) / num_m;
        const int kb0 = static_cast<int>(static_cast<long long>(split) * kb_total / p.splits);
        const int kb1 = static_cast<int>(static_cast<long long>(split + 1) * kb_total / p.splits);
        mbar_wait(smem_u32(&tempty_bar[as]), aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        uint32_t accumulate = 0;
        // the row sums are shared out over the column blocks of a row block: the tile of column block j adds the k-blocks with
        // kb % num_n == j (an N = 16 MMA costs far more than 1/16 of the N = 256 one, and a launch is as slow as its slowest tile)
        const int sum_blk = w % num_n;
        uint32_t acc_sums = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          if (p.dbg & 4) break;
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tc_fence_after();
          uint8_t* st = smem + stage * C::kStageBytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // plane pairs: (hi,hi) always; (hi,lo) and (lo,hi) in the 3-pass fp32-accurate mode
#pragma unroll
            for (int pr = 0; pr < (kPlanes == 1 ? 1 : 3); ++pr) {
              const int pa = (pr == 2) ? 1 : 0;
              const int pb = (pr == 1) ? 1 : 0;
              const uint32_t sa = smem_u32(st + pa * C::kAPlane);
              const uint32_t sb = smem_u32(st + kPlanes * C::kAPlane + pb * C::kBPlane);
              const uint64_t adesc = kAMN ? make_smem_desc_sw128(sa + k * 16 * 128, p.mn_lbo, p.mn_sbo)
                                          : make_smem_desc_sw128(sa + k * 32, 16, 1024);
              const uint64_t bdesc = kBMN ? make_smem_desc_sw128(sb + k * 16 * 128, p.mn_lbo, p.mn_sbo)
                                          : make_smem_desc_sw128(sb + k * 32, 16, 1024);
              if constexpr (kCta2) tc_mma_f16_ss_cta2(d_tmem, adesc, bdesc, idesc, accumulate);
              else tc_mma_f16_ss(d_tmem, adesc, bdesc, idesc, accumulate);
              accumulate = 1;
            }
          }
          if constexpr (kRowsum) {
            if (kb % num_n == sum_blk) {   // [BM x 16] += A . ones^T, after the block's main MMAs (switching shape per k-step stalls the pipe)
              constexpr uint32_t idesc1 = make_idesc_bf16(BM * kCtas, 16, kAMN, false);
              const uint32_t sa = smem_u32(st);
              const uint32_t d_sums = tmem_base + C::kSumCol + as * 16;
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {
                const uint64_t adesc = kAMN ? make_smem_desc_sw128(sa + k * 16 * 128, p.mn_lbo, p.mn_sbo)
                                            : make_smem_desc_sw128(sa + k * 32, 16, 1024);
                const uint64_t odesc = make_smem_desc_sw128(smem_u32(epi_stage) + k * 32, 16, 1024);
                if constexpr (kCta2) tc_mma_f16_ss_cta2(d_sums, adesc, odesc, idesc1, acc_sums);
                else tc_mma_f16_ss(d_sums, adesc, odesc, idesc1, acc_sums);
                acc_sums = 1;
              }
            }
          }
          commit(&empty_bar[stage]);
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        commit(&tfull_bar[as]);
        if (++as == C::kAccStages) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp == kEpiWarps + 2) {
    if constexpr (kSide) {
      if (lane == 0) {
        // slot c of the ring holds, for chunk steps 2c and 2c+1 of the epilogue, the [128 x 64] block each column
        // half's warps read; it is refilled for this CTA's next tile as soon as all eight warps have copied it to registers, so the
        // side operand runs about one tile-epilogue ahead of its use.  Out-of-range rows / columns arrive as zeros.
        tma_prefetch_desc(&tmSide);
        uint32_t sphase = 0;
        for (int w = w0; w < total; w += wstep) {
          const int n_blk = w % num_n;
          const int m_row = (((w / num_n) % num_m) * kCtas + cta_rank) * BM;
#pragma unroll 1
          for (int c = 0; c < C::kSideSlots; ++c) {
            if (p.dbg & 16) break;
            mbar_wait(smem_u32(&sempty_bar[c]), sphase ^ 1);
            const uint32_t fb = smem_u32(&sfull_bar[c]);
            mbar_arrive_expect_tx(fb, kSideSlotBytes);
            const uint32_t dst = smem_u32(side_ring + c * kSideSlotBytes);
            tma_load_2d(dst, &tmSide, fb, n_blk * BN + c * 64, m_row);
            tma_load_2d(dst + kSideSlotBytes / 2, &tmSide, fb, n_blk * BN + BN / 2 + c * 64, m_row);
          }
          sphase ^= 1;
        }
      }
    }
  } else if (warp < kEpiWarps) {
    const int ew = warp;
    const int quad = ew & 3;     // == warp % 4: the TMEM lane quadrant this warp may read
    const int half = ew >> 2;    // which half of the tile's columns this warp drains
    const uint32_t bias_addr = smem_u32(epi_stage + ew * C::kEpiWarpBytes);   // this warp's bias slab
    const uint32_t stage_addr = bias_addr + kBiasSlab;                       // and its staging rows
    const bool staged = kEpi == kEpiGeneric && p.out_bf16 != nullptr && p.out_bf16_lo == nullptr && !p.atomic_out;
    int as = 0;
    uint32_t aphase = 0;
    // fast-kernel state that lives across tiles (dead code in the other kernels)
    const float dscale = has_stage(kEpi, kStDrop) ? p.dropout_scale : 1.0f;
    const float scale = p.alpha * dscale;
    bool bias_valid = false;
    float4 bias_nx = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t sphase = 0;
    auto load_bias = [&](int cb) {
      const int c = cb + 4 * lane;
      return (p.bias && c < p.N) ? __ldg(reinterpret_cast<const float4*>(p.bias + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    for (int w = w0; w < total; w += wstep) {
      const int n_blk = w % num_n;
      const int m_blk = (w / num_n) % num_m;
      const int split = (w / num_n) / num_m;
      const int kb0 = static_cast<int>(static_cast<long long>(split) * kb_total / p.splits);
      const int kb1 = static_cast<int>(static_cast<long long>(split + 1) * kb_total / p.splits);
      const long long row0 = (static_cast<long long>(m_blk) * kCtas + cta_rank) * BM + quad * 32;
      const long long row = row0 + lane;
      const bool row_ok = row < p.M && kb1 > kb0;
      // operands of the fused stages that live in global memory are fetched coalesced (one warp = 32 rows) and
      // one chunk ahead: the first fetch is issued before waiting for the accumulator, the next one while the
      // current chunk is processed
      if constexpr (is_fast(kEpi)) {
        // thread == row throughout: the row's 32 mask / residual values of a chunk come from the TMA-fed ring, the
        // bf16 result leaves by 256-bit stores.  The bias of the next tile is fetched a tile ahead.
        const int colbase = n_blk * BN + half * (BN / 2);
        const int wn = w + wstep;
        const bool has_next = wn < total;
        if (4 * lane < BN / 2) {   // this warp's bias columns (pre-scaled; zeros without a bias) -> shared memory
          if (!bias_valid) bias_nx = load_bias(colbase);
          sts128(bias_addr + 16 * lane, __float_as_uint(bias_nx.x * dscale), __float_as_uint(bias_nx.y * dscale),
                 __float_as_uint(bias_nx.z * dscale), __float_as_uint(bias_nx.w * dscale));
          if (has_next) bias_nx = load_bias((wn % num_n) * BN + half * (BN / 2));
        }
        bias_valid = has_next;
        RowLn ln{1.0f, 0.0f, 1.0f, 0.0f, 0.0f, 0.0f};
        if constexpr (has_stage(kEpi, kStLnIn) || has_stage(kEpi, kStLnRes)) {
          // the extra per-column vectors of the folded LayerNorm (c | gamma, beta) next to the bias slab, and this row's
          // statistics turned into (rstd, -mean * rstd)
          if (4 * lane < BN / 2) {
            const int c = colbase + 4 * lane;
            const float* v1 = has_stage(kEpi, kStLnIn) ? p.ln_in_c : p.ln_res_gamma;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 a = c < p.N ? __ldg(reinterpret_cast<const float4*>(v1 + c)) : z;
            sts128(bias_addr + kBiasSlab + 16 * lane, __float_as_uint(a.x), __float_as_uint(a.y), __float_as_uint(a.z), __float_as_uint(a.w));
            if constexpr (has_stage(kEpi, kStLnRes)) {
              const float4 b = c < p.N ? __ldg(reinterpret_cast<const float4*>(p.ln_res_beta + c)) : z;
              sts128(bias_addr + 2 * kBiasSlab + 16 * lane, __float_as_uint(b.x), __float_as_uint(b.y), __float_as_uint(b.z), __float_as_uint(b.w));
            }
          }
          if (row < p.M) {
            // the producer's column-block partials, summed in a fixed order (deterministic; no atomics, no zero-fill)
            const float2* sp = reinterpret_cast<const float2*>(has_stage(kEpi, kStLnIn) ? p.ln_in_stats : p.ln_res_stats) + row * p.ln_slots;
            float2 st = __ldg(sp);
            for (int k = 1; k < p.ln_slots; ++k) { const float2 t = __ldg(sp + k); st.x += t.x; st.y += t.y; }
            const float mean = st.x * p.ln_inv_d;
            const float rstd = rsqrtf(fmaxf(fmaf(-mean, mean, st.y * p.ln_inv_d), 0.0f) + p.ln_eps);
            if constexpr (has_stage(kEpi, kStLnIn)) { ln.rs = rstd; ln.rb = -mean * rstd; }
            else { ln.ls = rstd; ln.lb = -mean * rstd; }
          }
        }
        __syncwarp();
        uint32_t side[16];
        const __nv_bfloat16* side_row = nullptr;
        if constexpr (side_ldg(kEpi)) {   // first chunk's mask / residual values, in flight while the main loop finishes
          side_row = reinterpret_cast<const __nv_bfloat16*>(has_stage(kEpi, kStRes) ? p.residual : p.relu_mask) +
                     row * (has_stage(kEpi, kStRes) ? p.ld_residual : p.ld_mask);
#pragma unroll
          for (int i = 0; i < 16; ++i) side[i] = 0;
          if (row < p.M && colbase < p.N) { ldg256(side_row + colbase, side); ldg256(side_row + colbase + 16, side + 8); }
        }
        mbar_wait(smem_u32(&tfull_bar[as]), aphase);
        tc_fence_after();
        const int trow = quad * 32 + lane;   // row inside the tile
#pragma unroll 1
        for (int c = 0; c < BN / 64; ++c) {
          const int tcol = half * (BN / 2) + c * 32;   // column inside the tile
          const int col0 = n_blk * BN + tcol;
          const bool live = col0 < p.N && !(p.dbg & 8);   // warp-uniform
          uint32_t r[32];
          if (live) tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN + tcol, r);
          if constexpr (kSide) {   // copy this row's 64 B out of the slot (SWIZZLE_128B: 16 B chunk ^ (row & 7)); free it after its second half
            const int slot = c >> 1;
            if ((c & 1) == 0 && !(p.dbg & 16)) mbar_wait(smem_u32(&sfull_bar[slot]), sphase);
            const uint32_t srow = smem_u32(side_ring + slot * kSideSlotBytes) + half * (kSideSlotBytes / 2) + trow * 128;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint4 q = lds128(srow + ((((c & 1) * 4 + g) ^ (trow & 7)) << 4));
              side[4 * g] = q.x; side[4 * g + 1] = q.y; side[4 * g + 2] = q.z; side[4 * g + 3] = q.w;
            }
          }
          // the slot is handed back only after its values have been consumed from registers (so the shared-memory
          // reads are certainly complete before TMA may overwrite the slot)
          auto release_slot = [&]() {
            if constexpr (kSide) {
              if (c & 1) {
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&sempty_bar[c >> 1]));
              }
            }
          };
          if (!live) { release_slot(); continue; }
          tmem_ld_wait_dep(r);
          if ((p.dbg & 3) == 1) { release_slot(); continue; }
          const unsigned long long e4 = (static_cast<unsigned long long>(row) * p.N + col0) >> 2;   // multiple of 8: + q never carries
          uint32_t out[16];
          epilogue_fast<kEpi>(p, scale, bias_addr + c * 128, static_cast<uint32_t>(e4),
                              static_cast<uint32_t>(e4 >> 32) ^ static_cast<uint32_t>(mix_seed(p.dropout_seed, p.seed_src) >> 32), r, side, out, ln);
          release_slot();
          if constexpr (side_ldg(kEpi)) {
            if (c + 1 < BN / 64 && col0 + 32 < p.N && row < p.M) { ldg256(side_row + col0 + 32, side); ldg256(side_row + col0 + 48, side + 8); }
          }
          if (row_ok && (p.dbg & 3) != 2) {
            __nv_bfloat16* orow = p.out_bf16 + row * p.ld_bf16 + col0;
            stg256(orow, out);
            stg256(orow + 16, out + 8);
          }
        }
        if constexpr (has_stage(kEpi, kStStats)) {
          if (row_ok)   // this warp's half of this column block: its own slot of the row (written exactly once)
            *reinterpret_cast<float2*>(p.stats_out + 2 * (row * (2 * num_n) + 2 * n_blk + half)) = make_float2(ln.sum, ln.sq);
        }
        sphase ^= 1;
        __syncwarp();   // every lane is done with the bias slab before the next tile overwrites it
      } else {
        const bool pre = kEpi == kEpiGeneric && !p.atomic_out;
        const bool pre_res = pre && p.residual && !p.residual_f32;
        const bool pre_mask = pre && p.relu_mask && !p.mask_f32;
        const int colbase = n_blk * BN + half * (BN / 2);
        uint4 pf[4];
        if (pre_res) fetch_issue(p.residual, p.ld_residual, row0, colbase, p.M, p.N, lane, pf);
        else if (pre_mask) fetch_issue(p.relu_mask, p.ld_mask, row0, colbase, p.M, p.N, lane, pf);
        if constexpr (!is_atomic(kEpi)) {
          // while the main loop of this tile runs: pull the rest of the residual / mask slab into L2 (the register
          // prefetch is only one chunk deep) and copy this warp's bias columns to shared memory
          if (pre_res || pre_mask) {
            const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(pre_res ? p.residual : p.relu_mask);
            const long long ld = pre_res ? p.ld_residual : p.ld_mask;
  #pragma unroll
            for (int j = 0; j < BN / 128; ++j) {
              const int line = lane + 32 * j;                       // 128 B lines of the [32 x BN/2] slab
              const long long r = row0 + line / (BN / 128);
              const int c = colbase + (line % (BN / 128)) * 64;
              if (r < p.M && c < p.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + r * ld + c));
            }
          }
          if (p.bias && !(kEpi == kEpiGeneric && p.atomic_out)) {
            const int c = colbase + 4 * lane;
            if (4 * lane < BN / 2) {
              float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
              if (c < p.N) b = __ldg(reinterpret_cast<const float4*>(p.bias + c));
              sts128(bias_addr + 16 * lane, __float_as_uint(b.x), __float_as_uint(b.y), __float_as_uint(b.z), __float_as_uint(b.w));
            }
            __syncwarp();
          }
        }
        mbar_wait(smem_u32(&tfull_bar[as]), aphase);
        tc_fence_after();
  #pragma unroll 1
        for (int c = 0; c < BN / 64; ++c) {
          if (p.dbg & 8) break;
          const int tcol = half * (BN / 2) + c * 32;   // column inside the tile
          const int col0 = n_blk * BN + tcol;
          if (col0 >= p.N) break;  // warp-uniform
          uint32_t mask_pk[16], res_pk[16];
          if (pre_res) {
            fetch_land(pf, stage_addr, lane, res_pk);
            if (c + 1 < BN / 64) fetch_issue(p.residual, p.ld_residual, row0, col0 + 32, p.M, p.N, lane, pf);
            if (pre_mask) {   // both (not on the hot path): the mask is fetched in place
              uint4 t[4];
              fetch_issue(p.relu_mask, p.ld_mask, row0, col0, p.M, p.N, lane, t);
              fetch_land(t, stage_addr, lane, mask_pk);
            }
          } else if (pre_mask) {
            fetch_land(pf, stage_addr, lane, mask_pk);
            if (c + 1 < BN / 64) fetch_issue(p.relu_mask, p.ld_mask, row0, col0 + 32, p.M, p.N, lane, pf);
          }
          uint32_t r[32];
          tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN + tcol, r);
          tmem_ld_wait_dep(r);
          if ((p.dbg & 3) == 1) continue;
          float v[32];
  #pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          const int ng = (p.N - col0) >= 32 ? 4 : (p.N - col0) >> 3;
          if (row_ok) epilogue32<kEpi>(p, row, col0, ng, v, mask_pk, res_pk, pre, bias_addr + c * 128);
          if (staged) {
            // thread == row: four 16 B chunks, XOR-swizzled so neither this write nor the read-back below conflicts
            const uint32_t wbase = stage_addr + lane * kStagePitch;
            const int sw = (lane >> 1) & 3;
  #pragma unroll
            for (int g = 0; g < 4; ++g)
              sts128(wbase + ((g ^ sw) << 4), pack_bf16x2(v[8 * g], v[8 * g + 1]), pack_bf16x2(v[8 * g + 2], v[8 * g + 3]),
                     pack_bf16x2(v[8 * g + 4], v[8 * g + 5]), pack_bf16x2(v[8 * g + 6], v[8 * g + 7]));
            __syncwarp();
            // coalesced write-out: 4 lanes x 16 B cover one staged row (64 B), 8 rows per pass
            const int seg = lane & 3;
            const int gcol = col0 + seg * 8;
            __nv_bfloat16* gptr = p.out_bf16 + (row0 + (lane >> 2)) * p.ld_bf16 + gcol;
  #pragma unroll
            for (int ps = 0; ps < 4; ++ps) {
              const int rr = ps * 8 + (lane >> 2);
              const uint4 val = lds128(stage_addr + rr * kStagePitch + ((seg ^ ((rr >> 1) & 3)) << 4));
              if (row0 + rr < p.M && gcol < p.N && kb1 > kb0 && (p.dbg & 3) != 2) *reinterpret_cast<uint4*>(gptr + static_cast<long long>(ps) * 8 * p.ld_bf16) = val;
            }
            __syncwarp();
          }
        }
        if constexpr (kRowsum) {   // this row's sum of A over the split's k range (16 identical columns: take the first)
          const int kb_first = kb0 + (n_blk - kb0 % num_n + num_n) % num_n;   // this tile's first row-sum k-block
          if (kb_first < kb1 && half == 0 && !(p.dbg & 8)) {
            uint32_t r1[16];
            tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + C::kSumCol + as * 16, r1);
            tmem_ld_wait_dep(r1);
            if (row_ok) atomicAdd(p.a_rowsum + row, __uint_as_float(r1[0]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        // the leader's MMA warp waits on it.  Default-semantics arrive (mbar_arrive_remote): the hand-back publishes nothing, and the
        // release.cluster form made every epilogue warp wait for its tile's global stores to drain first (ncu: `membar` among the
        // top three stalls of every pair kernel) - 0.8 ms of a 44 ms C5 step.  TVT_GEMM_DBG=32 keeps the old form for A/B runs.
        if constexpr (kCta2) {
          if (p.dbg & 32) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[as]), 0));
          else mbar_arrive_remote(mapa_u32(smem_u32(&tempty_bar[as]), 0));
        }
        else mbar_arrive(smem_u32(&tempty_bar[as]));
      }
      if (++as == C::kAccStages) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  if constexpr (kCta2) cluster_sync_all(); else __syncthreads();   // pair: neither CTA may exit while the other can still reach it
  tc_fence_after();
  if (warp == kEpiWarps + 1) {
    if constexpr (kCta2) tmem_dealloc_cta2(tmem_base, C::kTmemCols); else tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 2-D bf16 tensor map with 128B swizzle: dims {inner, outer}, row pitch ld elements.
static int make_map(CUtensorMap* m, const void* ptr, long long inner, long long outer, long long ld,
                    int box_inner, int box_outer, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return TVT_ECUDA;
  }
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(outer)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_inner), static_cast<cuuint32_t>(box_outer)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed with CUresult %d (inner=%lld outer=%lld ld=%lld)",
                   static_cast<int>(r), inner, outer, ld);
    return TVT_ECUDA;
  }
  return TVT_OK;
}

template <int BN, bool kAMN, bool kBMN, int kPlanes, int kEpi, bool kCta2 = false>
static int launch(const tvt_gemm_args* a, const Params& p, cudaStream_t stream) {
  using C = Cfg<BN, kPlanes, has_side(kEpi), is_fast(kEpi), kCta2, epi_slabs(kEpi), kEpi == kEpiAtomicSum>;
  CUtensorMap tmA, tmAlo, tmB, tmBlo, tmSide;
  int rc;
  auto mapA = [&](CUtensorMap* m, const void* ptr) {
    return kAMN ? make_map(m, ptr, a->m, a->k, a->lda, 64, BK) : make_map(m, ptr, a->k, a->m, a->lda, BK, BM);
  };
  auto mapB = [&](CUtensorMap* m, const void* ptr) {
    return kBMN ? make_map(m, ptr, a->n, a->k, a->ldb, 64, BK) : make_map(m, ptr, a->k, a->n, a->ldb, BK, C::kBRows);
  };
  if ((rc = mapA(&tmA, a->a)) != TVT_OK) return rc;
  if ((rc = mapB(&tmB, a->b)) != TVT_OK) return rc;
  if (kPlanes == 2) {
    if ((rc = mapA(&tmAlo, a->a_lo)) != TVT_OK) return rc;
    if ((rc = mapB(&tmBlo, a->b_lo)) != TVT_OK) return rc;
  } else {
    tmAlo = tmA;
    tmBlo = tmB;
  }
  tmSide = tmA;
  if constexpr (has_side(kEpi)) {   // the row-major [m, n] bf16 mask / residual, in [128 x 64] boxes
    const bool res = has_stage(kEpi, kStRes);
    if ((rc = make_map(&tmSide, res ? a->residual : a->relu_mask, a->n, a->m, res ? a->ld_residual : a->ld_mask, 64, BM)) != TVT_OK) return rc;
  }
  auto kern = gemm_kernel<BN, kAMN, kBMN, kPlanes, kEpi, kCta2>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [&] {
    attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
  });
  if (attr_err != cudaSuccess) {
    set_last_error("cudaFuncSetAttribute(gemm): %s", cudaGetErrorString(attr_err));
    return TVT_ECUDA;
  }
  constexpr int kCtas = kCta2 ? 2 : 1;
  const long long num_m = (a->m + BM * kCtas - 1) / (BM * kCtas), num_n = (a->n + BN - 1) / BN;
  const long long total = num_m * num_n * p.splits;   // tiles (of a CTA pair when kCta2)
  const long long slots = num_sms() / kCtas;
  const int grid = static_cast<int>(total < slots ? total : slots) * kCtas;
  const cudaError_t e = launch_pdl(kern, grid, kThreads, C::kSmemBytes, stream, kCta2 ? 2 : 1, tmA, tmAlo, tmB, tmBlo, tmSide, p);
  if (e != cudaSuccess) {
    set_last_error("tvt_gemm: launch failed: %s", cudaGetErrorString(e));
    return TVT_ECUDA;
  }
  return check_launch("tvt_gemm");
}

template <int BN, int kPlanes, int kEpi>
static int dispatch_major(const tvt_gemm_args* a, const Params& p, cudaStream_t s) {
  if (!a->a_mn_major && !a->b_mn_major) return launch<BN, false, false, kPlanes, kEpi>(a, p, s);
  if (!a->a_mn_major && a->b_mn_major) return launch<BN, false, true, kPlanes, kEpi>(a, p, s);
  if (a->a_mn_major && a->b_mn_major) return launch<BN, true, true, kPlanes, kEpi>(a, p, s);
  return launch<BN, true, false, kPlanes, kEpi>(a, p, s);
}

template <int kPlanes, int kEpi>
static int dispatch_width(bool narrow, const tvt_gemm_args* a, const Params& p, cudaStream_t s) {
  return narrow ? dispatch_major<128, kPlanes, kEpi>(a, p, s) : dispatch_major<256, kPlanes, kEpi>(a, p, s);
}

}  // namespace gemm
}  // namespace tvt

// Bring-up hook (not part of include/tvt.h): override the MN-major descriptor byte offsets.
extern "C" void tvt_debug_set_mn_desc(unsigned lbo, unsigned sbo) {
  tvt::gemm::g_mn_lbo = lbo;
  tvt::gemm::g_mn_sbo = sbo;
}

extern "C" void tvt_debug_set_epilogue(int mode) { tvt::gemm::g_dbg = mode; }
extern "C" void tvt_debug_set_pair(int on) { tvt::gemm::g_pair = on; }
extern "C" long long tvt_debug_gemm_fast_fallbacks() { return tvt::gemm::g_fast_fallbacks; }

// 1 when tvt_gemm would run an [m, n, k] bf16 forward GEMM (K-major operands, no split-K) on the CTA-pair [256 x 256] tiles, i.e.
// when the LayerNorm-folded epilogues are available for it (the same selection as in tvt_gemm below).
extern "C" int tvt_gemm_ln_fold_supported(int64_t m, int64_t n, int64_t k) {
  if (m <= 0 || n <= 0 || k <= 0 || n % 32 != 0 || k % 8 != 0 || !tvt::gemm::g_pair) return 0;
  const long long nsm = tvt::num_sms(), kb = (k + tvt::gemm::BK - 1) / tvt::gemm::BK, m_tiles = (m + 127) / 128;
  const long long w256 = m_tiles * ((n + 255) / 256), w128 = m_tiles * ((n + 127) / 128);
  const long long c256 = ((w256 + nsm - 1) / nsm) * (kb * 512 + 3000), c128 = ((w128 + nsm - 1) / nsm) * (kb * 400 + 1800);
  if (n <= 128 || c128 < c256) return 0;
  const long long wpair = ((m + 255) / 256) * ((n + 255) / 256), npair = nsm / 2;
  return (wpair + npair - 1) / npair <= (w256 + nsm - 1) / nsm ? 1 : 0;
}

// 1 when tvt_gemm would run this split-K / accumulating fp32 GEMM with MN-major operands (the wgrad orientation) on the CTA-pair
// [256 x 256] tiles, i.e. when tvt_gemm_args.a_rowsum is available for it (the same selection as in tvt_gemm below).
extern "C" int tvt_gemm_rowsum_supported(int64_t m, int64_t n, int64_t k, int32_t splits) {
  if (m <= 0 || n <= 0 || k <= 0 || splits < 1 || m % 8 || n % 8 || !tvt::gemm::g_pair) return 0;
  const long long nsm = tvt::num_sms(), kb = (k + tvt::gemm::BK - 1) / tvt::gemm::BK, kb_per = (kb + splits - 1) / splits;
  if (splits > kb) return 0;
  const long long m_tiles = (m + 127) / 128;
  const long long w256 = m_tiles * ((n + 255) / 256) * splits, w128 = m_tiles * ((n + 127) / 128) * splits;
  const long long c256 = ((w256 + nsm - 1) / nsm) * (kb_per * 512 + 3000), c128 = ((w128 + nsm - 1) / nsm) * (kb_per * 400 + 1800);
  if (n <= 128 || c128 < c256) return 0;
  const long long wpair = ((m + 255) / 256) * ((n + 255) / 256) * splits, npair = nsm / 2;
  return (wpair + npair - 1) / npair <= (w256 + nsm - 1) / nsm ? 1 : 0;
}

extern "C" int tvt_gemm(const tvt_gemm_args* a, void* stream) {
  using namespace tvt;
  TVT_REQUIRE(a != nullptr, "tvt_gemm: null args");
  TVT_REQUIRE(a->a && a->b, "tvt_gemm: null operand");
  TVT_REQUIRE(a->m > 0 && a->n > 0 && a->k > 0, "tvt_gemm: m, n, k must be positive (got %lld %lld %lld)",
              (long long)a->m, (long long)a->n, (long long)a->k);
  TVT_REQUIRE(a->m < (1ll << 31) && a->n < (1ll << 31) && a->k < (1ll << 31), "tvt_gemm: dims exceed int32");
  TVT_REQUIRE((a->a_lo == nullptr) == (a->b_lo == nullptr), "tvt_gemm: a_lo and b_lo must be given together");
  TVT_REQUIRE(a->n % 8 == 0, "tvt_gemm: n must be a multiple of 8 (got %lld)", (long long)a->n);
  TVT_REQUIRE(a->lda % 8 == 0 && a->ldb % 8 == 0, "tvt_gemm: lda/ldb must be multiples of 8 elements");
  TVT_REQUIRE(a->lda >= (a->a_mn_major ? a->m : a->k), "tvt_gemm: lda too small");
  TVT_REQUIRE(a->ldb >= (a->b_mn_major ? a->n : a->k), "tvt_gemm: ldb too small");
  TVT_REQUIRE(!(a->a_mn_major && a->m % 8) && !(!a->a_mn_major && a->k % 8) && !(!a->b_mn_major && a->k % 8),
              "tvt_gemm: contiguous operand extent must be a multiple of 8 elements");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  TVT_REQUIRE(al16(a->a) && al16(a->b) && al16(a->a_lo) && al16(a->b_lo), "tvt_gemm: operands must be 16-byte aligned");
  TVT_REQUIRE(a->splits >= 1, "tvt_gemm: splits must be >= 1");
  TVT_REQUIRE(a->splits == 1 || a->atomic_out, "tvt_gemm: split-K requires atomic_out");
  TVT_REQUIRE(a->out_f32 || a->out_bf16, "tvt_gemm: no output given");
  if (a->atomic_out) {
    TVT_REQUIRE(a->out_f32 && !a->out_bf16 && !a->bias && !a->residual && !a->relu_mask && !a->gelu_gate &&
                    a->act == TVT_ACT_NONE && a->dropout_p == 0.0f && !a->out_preact,
                "tvt_gemm: atomic_out only supports a plain fp32 accumulate epilogue");
  }
  TVT_REQUIRE(a->act >= TVT_ACT_NONE && a->act <= TVT_ACT_GELU, "tvt_gemm: bad act");
  TVT_REQUIRE(a->dropout_p >= 0.0f && a->dropout_p < 1.0f, "tvt_gemm: dropout_p must be in [0, 1)");
  TVT_REQUIRE(!a->out_f32 || (a->ld_f32 >= a->n && a->ld_f32 % 4 == 0 && al16(a->out_f32)), "tvt_gemm: bad out_f32 / ld_f32");
  TVT_REQUIRE(!a->out_bf16 || (a->ld_bf16 >= a->n && a->ld_bf16 % 8 == 0 && al16(a->out_bf16) && al16(a->out_bf16_lo)),
              "tvt_gemm: bad out_bf16 / ld_bf16");
  TVT_REQUIRE(!a->out_bf16_lo || a->out_bf16, "tvt_gemm: out_bf16_lo without out_bf16");
  TVT_REQUIRE(!a->residual || (a->ld_residual >= a->n && a->ld_residual % 8 == 0 && al16(a->residual)), "tvt_gemm: bad residual");
  TVT_REQUIRE(!a->relu_mask || (a->ld_mask >= a->n && a->ld_mask % 8 == 0 && al16(a->relu_mask)), "tvt_gemm: bad relu_mask");
  TVT_REQUIRE(!a->gelu_gate || (a->ld_gate >= a->n && a->ld_gate % 8 == 0 && al16(a->gelu_gate)), "tvt_gemm: bad gelu_gate");
  TVT_REQUIRE(!a->out_preact || (a->ld_preact >= a->n && a->ld_preact % 8 == 0 && al16(a->out_preact)), "tvt_gemm: bad out_preact");
  TVT_REQUIRE(!a->bias || al16(a->bias), "tvt_gemm: bias must be 16-byte aligned");
  const long long kb_total = (a->k + gemm::BK - 1) / gemm::BK;
  TVT_REQUIRE(a->splits <= kb_total, "tvt_gemm: more splits (%d) than k-blocks (%lld)", a->splits, kb_total);
  TVT_REQUIRE(!a->a_rowsum || (a->atomic_out && !a->a_lo && a->a_mn_major && a->b_mn_major && (reinterpret_cast<uintptr_t>(a->a_rowsum) & 3) == 0),
              "tvt_gemm: a_rowsum needs atomic_out, single bf16 planes and MN-major operands (the wgrad orientation)");

  int rc = require_sm100();
  if (rc != TVT_OK) return rc;

  gemm::Params p{};
  p.M = (int)a->m; p.N = (int)a->n; p.K = (int)a->k; p.splits = a->splits;
  p.act = a->act; p.alpha = a->alpha == 0.0f ? 1.0f : a->alpha;
  p.bias = a->bias;
  p.residual = a->residual; p.residual_f32 = a->residual_dtype == TVT_F32; p.ld_residual = a->ld_residual;
  p.relu_mask = a->relu_mask; p.mask_f32 = a->mask_dtype == TVT_F32; p.ld_mask = a->ld_mask;
  p.gelu_gate = a->gelu_gate; p.gate_f32 = a->gate_dtype == TVT_F32; p.ld_gate = a->ld_gate;
  if (a->dropout_p > 0.0f) {
    p.dropout_thr16 = (unsigned)(a->dropout_p * 65536.0f + 0.5f);
    p.dropout_scale = 65536.0f / (65536.0f - (float)p.dropout_thr16);
    p.dropout_seed = a->dropout_seed;
    p.seed_src = seed_source();
    for (int r = 0; r < kDropoutRounds; ++r) p.drop_rk[r] = static_cast<unsigned>(a->dropout_seed) + r * kDropoutWeyl;
  }
  p.out_preact = a->out_preact; p.preact_f32 = a->preact_dtype == TVT_F32; p.ld_preact = a->ld_preact;
  p.out_f32 = a->out_f32; p.ld_f32 = a->ld_f32; p.atomic_out = a->atomic_out; p.a_rowsum = a->a_rowsum;
  p.ln_in_stats = a->ln_in_stats; p.ln_in_c = a->ln_in_c; p.ln_res_stats = a->ln_res_stats;
  p.ln_res_gamma = a->ln_res_gamma; p.ln_res_beta = a->ln_res_beta; p.stats_out = a->stats_out;
  p.ln_inv_d = a->ln_dim > 0 ? 1.0f / static_cast<float>(a->ln_dim) : 0.0f; p.ln_eps = a->ln_eps;
  p.ln_slots = a->ln_dim > 0 ? 2 * static_cast<int>((a->ln_dim + 255) / 256) : 0;
  p.out_bf16 = (__nv_bfloat16*)a->out_bf16; p.out_bf16_lo = (__nv_bfloat16*)a->out_bf16_lo; p.ld_bf16 = a->ld_bf16;

  p.mn_lbo = gemm::g_mn_lbo; p.mn_sbo = gemm::g_mn_sbo; p.dbg = gemm::g_dbg;

  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // Tile width by a small cost model (cycles): waves x (k-blocks x MMA time per k-block + epilogue).  128-wide
  // tiles double the CTA count but run the main loop at the shared-memory bandwidth limit (A and B tiles are
  // the same size), so they only win when 256-wide tiles would leave SMs idle on a short-K problem.
  const long long nsm = tvt::num_sms();
  const long long m_tiles = (a->m + 127) / 128;
  const long long kb_per = (kb_total + a->splits - 1) / a->splits;
  const long long w256 = m_tiles * ((a->n + 255) / 256) * a->splits, w128 = m_tiles * ((a->n + 127) / 128) * a->splits;
  const long long c256 = ((w256 + nsm - 1) / nsm) * (kb_per * 512 + 3000);
  const long long c128 = ((w128 + nsm - 1) / nsm) * (kb_per * 400 + 1800);
  const bool narrow = a->n <= 128 || c128 < c256;
  // Epilogue kind: the small specialised kernels whenever the request fits them (see epilogue32); the fast ones
  // store whole 32 B sectors per lane, so the bf16 output must be 32-byte aligned row by row
  auto al32 = [](const void* ptr, long long ld) { return ptr == nullptr || ((reinterpret_cast<uintptr_t>(ptr) & 31) == 0 && ld % 16 == 0); };
  const bool fast = !a->atomic_out && a->out_bf16 && !a->out_bf16_lo && !a->out_f32 && !a->out_preact && !a->gelu_gate &&
                    (a->act == TVT_ACT_NONE || a->act == TVT_ACT_RELU) && !(a->residual && a->residual_dtype == TVT_F32) &&
                    !(a->relu_mask && a->mask_dtype == TVT_F32) && !(a->residual && a->relu_mask) && a->n % 32 == 0 &&
                    al32(a->out_bf16, a->ld_bf16);
  if (a->a_rowsum) {
    TVT_REQUIRE(tvt_gemm_rowsum_supported(a->m, a->n, a->k, a->splits),
                "tvt_gemm: a_rowsum is built for the CTA-pair [256 x 256] tiles (ask tvt_gemm_rowsum_supported first)");
    return gemm::launch<256, true, true, 1, gemm::kEpiAtomicSum, true>(a, p, s);
  }
  if (a->a_lo) {
    if (a->atomic_out) return gemm::dispatch_width<2, gemm::kEpiAtomic>(narrow, a, p, s);
    return gemm::dispatch_width<2, gemm::kEpiGeneric>(narrow, a, p, s);
  }
  // CTA pairs ([256 x 256] tiles) whenever 256-wide tiles are chosen and pairing does not add a wave
  const long long wpair = ((a->m + 255) / 256) * ((a->n + 255) / 256) * a->splits, npair = nsm / 2;
  const bool pair = !narrow && gemm::g_pair && (wpair + npair - 1) / npair <= (w256 + nsm - 1) / nsm;   // no extra wave
  if (a->atomic_out) {
    if (pair) {
      if (!a->a_mn_major && !a->b_mn_major) return gemm::launch<256, false, false, 1, gemm::kEpiAtomic, true>(a, p, s);
      if (!a->a_mn_major && a->b_mn_major) return gemm::launch<256, false, true, 1, gemm::kEpiAtomic, true>(a, p, s);
      if (a->a_mn_major && a->b_mn_major) return gemm::launch<256, true, true, 1, gemm::kEpiAtomic, true>(a, p, s);
      return gemm::launch<256, true, false, 1, gemm::kEpiAtomic, true>(a, p, s);
    }
    return gemm::dispatch_width<1, gemm::kEpiAtomic>(narrow, a, p, s);
  }
  const bool ln_any = a->ln_in_stats || a->ln_in_c || a->ln_res_stats || a->ln_res_gamma || a->ln_res_beta || a->stats_out;
  if (ln_any) {
    // LayerNorm-folded inference epilogues: CTA-pair fast kernels only (what the encoder layers' forward GEMMs use)
    using namespace gemm;
    TVT_REQUIRE(fast && !a->a_mn_major && !a->b_mn_major && a->dropout_p == 0.0f && !a->relu_mask && a->bias,
                "tvt_gemm: the LayerNorm-folded epilogues need the bf16 fast path (K-major operands, bias, no dropout / mask)");
    TVT_REQUIRE((a->ln_in_stats != nullptr) == (a->ln_in_c != nullptr), "tvt_gemm: ln_in_stats and ln_in_c go together");
    TVT_REQUIRE((a->ln_res_stats != nullptr) == (a->ln_res_gamma != nullptr) && (a->ln_res_stats != nullptr) == (a->ln_res_beta != nullptr),
                "tvt_gemm: ln_res_stats, ln_res_gamma and ln_res_beta go together");
    TVT_REQUIRE(!(a->ln_in_stats && a->ln_res_stats), "tvt_gemm: ln_in and ln_res cannot be combined in one call");
    TVT_REQUIRE(!a->ln_res_stats || a->residual, "tvt_gemm: ln_res needs the pre-norm residual tensor");
    TVT_REQUIRE((!a->ln_in_stats && !a->ln_res_stats) || (a->ln_dim > 0 && a->ln_eps > 0.0f), "tvt_gemm: ln_dim / ln_eps missing");
    TVT_REQUIRE(!(a->ln_in_stats && (a->residual || a->stats_out)), "tvt_gemm: ln_in supports bias (+ relu) only");
    TVT_REQUIRE(!a->stats_out || a->residual, "tvt_gemm: stats_out is implemented for the residual-adding GEMMs");
    TVT_REQUIRE(!a->stats_out || (reinterpret_cast<uintptr_t>(a->stats_out) & 7) == 0, "tvt_gemm: stats_out must be 8-byte aligned");
    TVT_REQUIRE(al16(a->ln_in_c) && al16(a->ln_res_gamma) && al16(a->ln_res_beta) && (reinterpret_cast<uintptr_t>(a->ln_in_stats) & 7) == 0 &&
                    (reinterpret_cast<uintptr_t>(a->ln_res_stats) & 7) == 0 && a->n % 4 == 0,
                "tvt_gemm: LayerNorm vectors must be 16-byte aligned, statistics 8-byte aligned");
    TVT_REQUIRE(pair && !narrow, "tvt_gemm: the LayerNorm-folded epilogues are built for the CTA-pair [256 x 256] tiles (m, n too small)");
    const bool side_ok = al32(a->residual, a->ld_residual);
    const bool ldg = a->residual && kb_per > 16 && side_ok;
    if (a->ln_in_stats) {
      if (a->act == TVT_ACT_RELU) return launch<256, false, false, 1, kEpiFast + kStLnIn + kStRelu, true>(a, p, s);
      return launch<256, false, false, 1, kEpiFast + kStLnIn, true>(a, p, s);
    }
    TVT_REQUIRE(a->act == TVT_ACT_NONE && a->stats_out, "tvt_gemm: residual GEMMs of the folded path produce statistics and have no activation");
    if (a->ln_res_stats) {
      if (ldg) return launch<256, false, false, 1, kEpiFast + kSideLdg + kStRes + kStLnRes + kStStats, true>(a, p, s);
      return launch<256, false, false, 1, kEpiFast + kStRes + kStLnRes + kStStats, true>(a, p, s);
    }
    if (ldg) return launch<256, false, false, 1, kEpiFast + kSideLdg + kStRes + kStStats, true>(a, p, s);
    return launch<256, false, false, 1, kEpiFast + kStRes + kStStats, true>(a, p, s);
  }
  if (fast) {
    // exact-stage kernels for the combinations the encoder layers launch (forward: both operands K-major; dgrad:
    // B MN-major); anything else takes the fast kernel that checks its stages at run time
    using namespace gemm;
    const int st = (a->act == TVT_ACT_RELU ? kStRelu : 0) | (a->relu_mask ? kStMask : 0) | (a->dropout_p > 0.0f ? kStDrop : 0) | (a->residual ? kStRes : 0);
    const bool side_al32 = al32(a->residual, a->ld_residual) && al32(a->relu_mask, a->ld_mask);
    const bool ldg = (st & (kStMask | kStRes)) && kb_per > 16 && side_al32;   // long K: keep four operand stages (see kSideLdg)
#define TVT_FAST_KIND(AMN, BMN, KIND)                                                                                 \
  return narrow ? launch<128, AMN, BMN, 1, KIND>(a, p, s)                                                             \
                : (pair ? launch<256, AMN, BMN, 1, KIND, true>(a, p, s) : launch<256, AMN, BMN, 1, KIND>(a, p, s));
#define TVT_FAST_CASE(AMN, BMN, ST)                                                                                   \
  if (a->a_mn_major == AMN && a->b_mn_major == BMN && st == (ST)) {                                                   \
    if (((ST) & (kStMask | kStRes)) && ldg) { TVT_FAST_KIND(AMN, BMN, kEpiFast + kSideLdg + (ST)) }                   \
    TVT_FAST_KIND(AMN, BMN, kEpiFast + (ST))                                                                          \
  }
    TVT_FAST_CASE(false, false, 0)
    TVT_FAST_CASE(false, false, kStRelu)
    TVT_FAST_CASE(false, false, kStRelu | kStDrop)
    TVT_FAST_CASE(false, false, kStRes)
    TVT_FAST_CASE(false, false, kStDrop | kStRes)
    TVT_FAST_CASE(false, false, kStDrop)
    TVT_FAST_CASE(false, true, 0)
    TVT_FAST_CASE(false, true, kStRes)
    TVT_FAST_CASE(false, true, kStMask)
    TVT_FAST_CASE(false, true, kStMask | kStDrop)
    TVT_FAST_CASE(false, true, kStDrop)
#undef TVT_FAST_CASE
#undef TVT_FAST_KIND
    ++g_fast_fallbacks;   // no exact kernel for this combination: the generic one below
  }
  return gemm::dispatch_width<1, gemm::kEpiGeneric>(narrow, a, p, s);
}
