/*
 * tvt.h — C-ABI of libtvt_b200.so: hand-written sm_100a kernels for the temporal video transformer
 * hot path of ed-fish/data-efficient-video-transformers.
 *
 * The reference has no FFI of its own: its arithmetic is torch.nn library calls made
 * from src/models/*.py. Each entry point below names the reference call site(s) whose arithmetic it
 * replaces.  Conventions shared by every entry point:
 *   - plain pointers and sizes only; every buffer (inputs, outputs, workspaces) is owned by the caller
 *     and lives in device memory; `stream` is a cudaStream_t passed as void*;
 *   - returns TVT_OK (0) or a negative tvt_status; never throws, never allocates device memory,
 *     never synchronises the device or the stream (all entry points are CUDA-graph capturable);
 *   - argument validation happens before any CUDA call, so TVT_EINVAL is reported even on a machine
 *     without a GPU;
 *   - tvt_last_error() returns a thread-local, human readable description of the last failure.
 */
#ifndef TVT_H_
#define TVT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  TVT_OK = 0,
  TVT_EINVAL = -1,     /* bad shape / alignment / null pointer */
  TVT_EARCH = -2,      /* device is not sm_100 */
  TVT_EWORKSPACE = -3, /* workspace too small */
  TVT_ECUDA = -4       /* CUDA runtime / driver error, see tvt_last_error() */
} tvt_status;

typedef enum { TVT_BF16 = 0, TVT_F32 = 1 } tvt_dtype;
typedef enum { TVT_ACT_NONE = 0, TVT_ACT_RELU = 1, TVT_ACT_GELU = 2 } tvt_act;

const char* tvt_last_error(void);
int tvt_version(void);
/* TVT_OK when the current CUDA device is a B200-class part (compute capability 10.x). */
int tvt_device_check(void);

/* ------------------------------------------------------------------------------------------------
 * GEMM on tcgen05 tensor cores (TMA-fed, TMEM accumulators, fused epilogue).
 * Replaces every aten::addmm / aten::mm the reference reaches through
 *   nn.MultiheadAttention in/out projections  (torch/nn/functional.py:6478, :6690 via
 *                                              src/models/transformer.py:39-47,116)
 *   TransformerEncoderLayer FFN               (torch/nn/modules/transformer.py:981)
 *   Reasoning.relation Linear layers          (src/models/TPN.py:88-99)
 * and their autograd dgrad / wgrad (AddmmBackward0).
 *
 *   acc[M,N]  = sum_k A(m,k) * B(n,k)                 (fp32 accumulate in tensor memory)
 *   v         = alpha * acc + bias[n]
 *   preact    = v                       (optional store, for GELU backward)
 *   v         = act(v)
 *   v         = v * (mask[m,n] > 0)     (optional: ReLU backward fused into a dgrad GEMM)
 *   v         = v * gelu'(gate[m,n])    (optional: GELU backward fused into a dgrad GEMM)
 *   v         = dropout(v)              (optional; mask is a pure function of seed and element index)
 *   v         = v + residual[m,n]       (optional)
 *   out       = v                       (any of: fp32, bf16, bf16 hi/lo split planes, fp32 atomic add)
 *
 * Operands are bf16.  With a_lo / b_lo non-NULL each operand is the sum of two bf16 planes
 * (x = hi + lo, lo = bf16(x - hi)) and the kernel issues hi*hi + hi*lo + lo*hi: this is the
 * "fp32-accumulate" parity mode (relative error ~1e-5 per GEMM instead of ~2e-3).
 * Storage: a_mn_major = 0 -> A stored [M, lda] with K contiguous; 1 -> A stored [K, lda] with M contiguous.
 *          b_mn_major = 0 -> B stored [N, ldb] with K contiguous; 1 -> B stored [K, ldb] with N contiguous.
 * Alignment: operand base pointers 16 B, leading dimensions multiples of 8 elements.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const void* a;
  const void* a_lo;
  const void* b;
  const void* b_lo;
  int64_t m, n, k;
  int64_t lda, ldb;
  int32_t a_mn_major;
  int32_t b_mn_major;
  int32_t splits; /* split-K factor; > 1 requires atomic_out */
  int32_t act;    /* tvt_act */
  float alpha;
  const float* bias;      /* [N] fp32 or NULL */
  const void* residual;   /* [M, ld_residual] or NULL */
  int32_t residual_dtype; /* tvt_dtype */
  int64_t ld_residual;
  const void* relu_mask; /* [M, ld_mask]; output multiplied by (mask > 0) */
  int32_t mask_dtype;
  int64_t ld_mask;
  const void* gelu_gate; /* [M, ld_gate]; output multiplied by gelu'(gate) */
  int32_t gate_dtype;
  int64_t ld_gate;
  float dropout_p; /* 0 disables */
  uint64_t dropout_seed;
  void* out_preact; /* optional, dtype preact_dtype, [M, ld_preact] */
  int32_t preact_dtype;
  int64_t ld_preact;
  float* out_f32; /* optional [M, ld_f32] */
  int64_t ld_f32;
  int32_t atomic_out;       /* 1: red.add into out_f32 (bias/act/etc. must be unset) */
  void* out_bf16;           /* optional [M, ld_bf16] */
  void* out_bf16_lo;        /* optional lo plane (same ld) */
  int64_t ld_bf16;
} tvt_gemm_args;

int tvt_gemm(const tvt_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TVT_H_ */
