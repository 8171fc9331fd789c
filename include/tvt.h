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

#if defined(__GNUC__)
#define TVT_API __attribute__((visibility("default")))
#else
#define TVT_API
#endif

typedef enum {
  TVT_OK = 0,
  TVT_EINVAL = -1,     /* bad shape / alignment / null pointer */
  TVT_EARCH = -2,      /* device is not sm_100 */
  TVT_EWORKSPACE = -3, /* workspace too small */
  TVT_ECUDA = -4       /* CUDA runtime / driver error, see tvt_last_error() */
} tvt_status;

typedef enum { TVT_BF16 = 0, TVT_F32 = 1, TVT_F64 = 2 /* evaluation labels only */ } tvt_dtype;
typedef enum { TVT_ACT_NONE = 0, TVT_ACT_RELU = 1, TVT_ACT_GELU = 2 } tvt_act;

/* Note on SURVEY.md section 8(b2), which sketched `(args, workspace, workspace_bytes, stream)` signatures: no entry point
 * takes a workspace.  Every kernel either needs none or accumulates into caller-owned output buffers (split-K partial
 * sums meet in fp32 atomics on `out_f32`), so the caller (PyTorch's caching allocator) still owns every byte. */
TVT_API const char* tvt_last_error(void);
TVT_API int tvt_version(void);
/* TVT_OK when the current CUDA device is a B200-class part (compute capability 10.x). */
TVT_API int tvt_device_check(void);

/* Whole-step CUDA graphs (the training step of the launch-bound BASELINE configs replayed with one launch): a captured
 * graph bakes every kernel argument, so what must change from step to step lives in DEVICE memory instead.
 *   tvt_set_seed_source(ptr)        registers a device uint64 step counter (NULL = none; the default).  While one is registered,
 *                                   every kernel that draws a dropout mask folds a hash of *ptr into its seed's high word,
 *                                   forward and backward alike (the backward recomputes the forward's mask).
 *   tvt_step_counter_advance(p,n,s) one-thread kernel: p[0..n) += 1 (the seed counter and, next to it, the optimizer's step
 *                                   count read through tvt_optim_step_args.step_dev); the first node of a captured step.
 * Process-wide setting, not thread-safe against concurrent launches (the reference drives its model from one Python thread). */
TVT_API int tvt_set_seed_source(const void* device_counter);
TVT_API int tvt_step_counter_advance(void* device_counters, int count, void* stream);

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
  /* ---- LayerNorm folded into the GEMMs around it (inference: no LayerNorm kernel, its output is never materialised).
   * A post-norm layer's LN(y) = (y - mean) * rstd * gamma + beta feeds a Linear:  LN(y) W^T + b =
   *     rstd_m * (y (W diag(gamma))^T)[m, n]  -  rstd_m * mean_m * c_n  +  b'_n,   c = rowsum(W diag(gamma)),  b' = b + W beta,
   * so the consumer GEMM runs on the PRE-norm tensor y with B = W diag(gamma) (caller-prepared), `bias` = b', and
   *   ln_in_stats [M, slots, 2] fp32 (partial sums and sums of squares of y's row, slots = 2 * ceil(ln_dim / 256): one per half of
   *   every 256-column block of the GEMM that wrote y, summed in slot order by the consumer) + ln_in_c [N]:
   *       v = rstd_m * alpha * acc + (-mean_m rstd_m) * c_n + bias_n           (replaces v = alpha * acc + bias)
   * and LN(y) as the RESIDUAL of the next GEMM is recomputed in its epilogue from the same y:
   *   residual = y (pre-norm, bf16), ln_res_stats [M, slots, 2], ln_res_gamma / ln_res_beta [N]:
   *       v += ((residual - mean_m) * rstd_m) * gamma_n + beta_n               (replaces v += residual)
   * The row statistics are produced by the epilogue of the GEMM that writes y:
   *   stats_out [M, 2 * ceil(N / 256), 2] fp32: every epilogue warp stores the (sum, sum of squares) of its 128 columns of the
   *   row (fp32, before the bf16 rounding of the stored value) in its own slot - no atomics, no zero-fill, bit-reproducible.
   * bf16 fast path only (plain bf16 output, both operands K-major, n % 32 == 0); ln_dim = the normalised width, ln_eps. */
  const float* ln_in_stats;
  const float* ln_in_c;
  const float* ln_res_stats;
  const float* ln_res_gamma;
  const float* ln_res_beta;
  float* stats_out;
  int64_t ln_dim;
  float ln_eps;
  int32_t reserved;
  /* ---- bias gradient inside the weight-gradient GEMM: dW = dY^T X is this GEMM with A = dY (MN-major), so the bias gradient
   * colsum(dY) = A . 1 is the row sum of the A operand over the contraction.  a_rowsum [m] fp32 receives += sum_k A[m, k] (atomics,
   * like out_f32): one extra N = 16 tensor-core MMA per k-step against a tile of ones, no second pass over dY.  Needs atomic_out,
   * single bf16 planes, both operands MN-major and the CTA-pair tiles: ask tvt_gemm_rowsum_supported(m, n, k, splits) first. */
  float* a_rowsum;
} tvt_gemm_args;

TVT_API int tvt_gemm(const tvt_gemm_args* args, void* stream);
/* 1 when tvt_gemm would run this bf16 forward GEMM on the tiles that carry the LayerNorm-folded epilogues (ln_* fields), else 0:
 * callers choose between the folded inference path and LayerNorm as a kernel of its own with it (host-only, no launch). */
TVT_API int tvt_gemm_ln_fold_supported(int64_t m, int64_t n, int64_t k);
/* 1 when tvt_gemm_args.a_rowsum is available for this [m, n, k] accumulating (atomic_out) GEMM with `splits` k-splits. */
TVT_API int tvt_gemm_rowsum_supported(int64_t m, int64_t n, int64_t k, int32_t splits);


/* ------------------------------------------------------------------------------------------------
 * LayerNorm forward / backward over the last dimension (eps inside the sqrt, affine), fp32 statistics.
 * Replaces aten::native_layer_norm(+backward) reached from TransformerEncoderLayer.norm1/norm2
 * (torch/nn/modules/transformer.py:952-956), SimpleTransformer.norm (src/models/transformer.py:49,80)
 * and mlp_head[0] (src/models/transformer.py:54).
 *
 * Embed mode (seq_len = S > 0) fuses SimpleTransformer.add_pos_cls (src/models/transformer.py:74-82):
 * output row (b, s) = LN(dropout((s == 0 ? cls[b] : x[b, s-1]) + pe[s])), rows = B * S, tokens laid out
 * batch-major ([B, S, d]); `pre` optionally receives the pre-LN rows (needed by the backward pass).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const void* x;      /* [rows, d]; embed mode: frame features [B, S-1, d] */
  const void* cls;    /* embed mode: [B, d] */
  const float* pe;    /* embed mode: [S, d] fp32 */
  const float* gamma; /* [d] fp32 */
  const float* beta;  /* [d] fp32 */
  void* y;            /* [rows, d] */
  void* pre;          /* embed mode, optional: [rows, d] pre-LN rows */
  float* mean;        /* optional [rows] */
  float* rstd;        /* optional [rows] */
  int64_t rows, d, seq_len;
  int32_t dtype;      /* tvt_dtype of x / cls / y / pre */
  float eps;
  float dropout_p;    /* embed mode only: dropout after the positional encoding */
  uint64_t dropout_seed;
  /* optional blocked output (plain mode, d <= 1024 bf16 / 512 fp32): row r is written to y + (r / y_seq) * y_pitch + (r % y_seq) * d
   * elements, i.e. the y_seq rows of one clip land inside a wider per-clip block.  Builds the cross-attention memory of
   * src/models/transformer.py:110-121 (the other experts' tokens side by side) without a concatenation pass.  0 = dense [rows, d]. */
  int64_t y_seq, y_pitch;
} tvt_layernorm_fwd_args;
TVT_API int tvt_layernorm_fwd(const tvt_layernorm_fwd_args* args, void* stream);

/* dx = LN'(dy); dgamma += sum_rows dy * xhat; dbeta += sum_rows dy (atomics: caller zero-fills).
 * dz (optional) = dropout-masked dx with the mask of the GEMM epilogue that produced the branch
 * (element index row * d + col), dbias (optional) += column sums of dz (of dx when dz is NULL).
 * Embed mode scatters the (dropout-masked) dx rows to dfeat [B, S-1, d] and dcls [B, d].
 * dres (optional) is added to dx: in a pre-norm block (src/models/vit.py:71-75) x' = x + f(LN(x)), so
 * dL/dx = dL/dx' + LN'(dL/dLN). */
typedef struct {
  const void* dy;
  const void* x;      /* pre-LN input saved by the forward pass */
  const float* mean;
  const float* rstd;
  const float* gamma;
  void* dx;
  void* dz;
  void* dfeat;
  void* dcls;
  float* dgamma;      /* optional */
  float* dbeta;       /* optional */
  float* dbias;       /* optional */
  int64_t rows, d, seq_len;
  int32_t dtype;
  float dropout_p;
  uint64_t dropout_seed;
  const void* dres;   /* optional [rows, d]: added to dx (the residual-path gradient of a pre-norm block) */
} tvt_layernorm_bwd_args;
TVT_API int tvt_layernorm_bwd(const tvt_layernorm_bwd_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Scaled-dot-product attention, softmax(Q K^T * scale) V, forward and backward, any Sq / Sk (self- and
 * cross-modal attention share the kernel).  Replaces aten::scaled_dot_product_attention reached from
 * F.multi_head_attention_forward (torch/nn/functional.py:6682 via src/models/transformer.py:116) and
 * the explicit einsum/softmax/einsum of src/models/vit.py:51-55.
 * Layout: q/k/v/o are 2-D row-major token matrices; token (b, s) is row b * seq + s; head h occupies
 * columns [h * head_dim, (h+1) * head_dim) starting at the given pointer; ld* are row pitches in
 * elements (so q/k/v may alias one packed [n, 3d] in-projection output).  lse is [B, H, Sq] fp32.
 * impl: 0 = auto, 1 = fp32 CUDA-core kernel (any dtype; the fp32 parity mode), 2 = tcgen05 kernel (bf16).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const void* q; const void* k; const void* v;
  void* o;
  float* lse;
  int64_t batch, heads, sq, sk, head_dim;
  int64_t ldq, ldk, ldv, ldo;
  float scale;
  int32_t dtype;
  int32_t impl;
  float dropout_p;        /* attention-probability dropout */
  uint64_t dropout_seed;
} tvt_attention_fwd_args;
TVT_API int tvt_attention_fwd(const tvt_attention_fwd_args* args, void* stream);

typedef struct {
  const void* q; const void* k; const void* v; const void* o; const void* d_o;
  const float* lse;
  void* dq; void* dk; void* dv;
  int64_t batch, heads, sq, sk, head_dim;
  int64_t ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv;
  float scale;
  int32_t dtype;
  int32_t impl;
  float dropout_p;
  uint64_t dropout_seed;
} tvt_attention_bwd_args;
TVT_API int tvt_attention_bwd(const tvt_attention_bwd_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Temporal pyramid pooling: all scales of sum_group (src/models/TPN.py:64-72) in one pass, with the
 * leading nn.ReLU() of Reasoning.relation (src/models/TPN.py:89) fused:
 *   out_g[b, k*d + j] = relu(sum_{t in [g*k, g*(k+1))} x[b, t, j]),  k < floor(T/g)  (remainder dropped)
 * x rows are addressed as x + (b * x_batch_stride + t * x_frame_stride) (elements), so the frame
 * tokens of a [B, S, d] token tensor can be pooled in place (skip the CLS row via the base pointer).
 * Backward: dx[b,t,:] = sum_g [t < g*floor(T/g)] dout_g[b, t/g, :] * (out_g[b, t/g, :] > 0).
 * ---------------------------------------------------------------------------------------------- */
#define TVT_MAX_POOL_SCALES 8
typedef struct {
  const void* x;
  int64_t batch, frames, d;
  int64_t x_batch_stride, x_frame_stride;
  int32_t num_scales;
  int32_t groups[TVT_MAX_POOL_SCALES];
  void* out[TVT_MAX_POOL_SCALES];  /* out[i]: [B, floor(T/groups[i]) * d] */
  int32_t dtype;
  int32_t relu;
} tvt_pyramid_pool_fwd_args;
TVT_API int tvt_pyramid_pool_fwd(const tvt_pyramid_pool_fwd_args* args, void* stream);

typedef struct {
  const void* dout[TVT_MAX_POOL_SCALES];
  const void* out[TVT_MAX_POOL_SCALES]; /* forward outputs (ReLU mask); ignored when relu = 0 */
  void* dx;
  int64_t batch, frames, d;
  int64_t dx_batch_stride, dx_frame_stride;
  int32_t num_scales;
  int32_t groups[TVT_MAX_POOL_SCALES];
  int32_t dtype;
  int32_t relu;
  int32_t accumulate; /* 1: dx += ... (dx already holds another gradient) */
} tvt_pyramid_pool_bwd_args;
TVT_API int tvt_pyramid_pool_bwd(const tvt_pyramid_pool_bwd_args* args, void* stream);

/* Spatial pyramid pooling (src/models/TPN.py:2-40): global average of a [frames, C, HW] feature map
 * over HW -> [frames, C] written at column offset `col_offset` of a [frames, ld_out] matrix, so the
 * three levels land in the reference's concat order (high, mid, low) (src/models/TPN.py:58). */
typedef struct {
  const void* x;
  void* out;
  int64_t frames, channels, hw;
  int64_t ld_out, col_offset;
  int32_t dtype;     /* of x */
  int32_t out_dtype;
} tvt_spatial_pool_args;
TVT_API int tvt_spatial_pool_fwd(const tvt_spatial_pool_args* args, void* stream);

/* Gradient of the spatial average pool (the feature maps come from a trainable CNN in the reference's TPN,
 * src/models/TPN.py:46-53): dx[f, c, :] = dpooled[f, col_offset + c] / hw.  dpooled fp32 [frames, ld]; dx has
 * the dtype of the forward input and must be 16-byte aligned. */
typedef struct {
  const void* dpooled;
  void* dx;
  int64_t frames, channels, hw;
  int64_t ld, col_offset;
  int32_t dtype;     /* of dx */
  int32_t reserved;
} tvt_spatial_pool_bwd_args;
TVT_API int tvt_spatial_pool_bwd(const tvt_spatial_pool_bwd_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Classification / distillation loss, forward + gradient in one launch.
 * Replaces nn.BCEWithLogitsLoss (src/models/transformer.py:35,142; frame_transformer.py:89,251),
 * nn.CrossEntropyLoss on argmax(teacher) (frame_transformer.py:90,250), nn.CosineSimilarity monitor
 * (frame_transformer.py:121,257), plus the north-star KL term T^2 * kl_div(log_softmax(s/T),
 * softmax(t/T), 'batchmean') (torch-pinned extension).
 *   loss = w_bce * BCE + w_ce * CE + w_kl * KL
 * losses[0..4] = {total, bce, ce, kl, cos(student, teacher)[0]} (accumulated: caller zero-fills);
 * dlogits = d total / d student * grad_scale.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const float* student;  /* [B, C] fp32 logits */
  const float* teacher;  /* [B, C] or NULL (then w_ce, w_kl must be 0) */
  const float* target;   /* [B, C] multi-hot */
  float* losses;         /* [5] */
  float* dlogits;        /* [B, C] or NULL */
  int64_t batch, classes;
  float w_bce, w_ce, w_kl, temperature, grad_scale;
} tvt_distill_loss_args;
TVT_API int tvt_distill_loss(const tvt_distill_loss_args* args, void* stream);

/* Reasoning output stage (src/models/TPN.py:98,112): p = mean_g sigmoid(z_g); optional BCE(p, target)
 * (mean over B*C) accumulated into loss[0], and dz_g = dL/dz_g * grad_scale. z: [G, B, C] fp32. */
typedef struct {
  const float* z;
  const float* target; /* optional */
  float* prob;         /* [B, C] */
  float* loss;         /* optional [1], accumulated */
  float* dz;           /* optional [G, B, C] */
  int64_t scales, batch, classes;
  float grad_scale;
} tvt_pyramid_head_args;
TVT_API int tvt_pyramid_head(const tvt_pyramid_head_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Small helpers on the path.
 * ---------------------------------------------------------------------------------------------- */
/* out[c] += sum_rows x[r, c] (fp32 atomics; caller zero-fills). Bias gradients of Linear layers. */
typedef struct {
  const void* x;
  float* out;
  int64_t rows, cols, ld;
  int32_t dtype;
} tvt_colsum_args;
TVT_API int tvt_colsum(const tvt_colsum_args* args, void* stream);

/* Convert fp32 -> bf16 (hi) and optionally the bf16 residual (lo = bf16(x - hi)); n elements.
 * Used for weights every step and for activations in the fp32 parity mode. */
typedef struct {
  const float* x;
  void* hi;
  void* lo; /* optional */
  int64_t n;
} tvt_split_args;
TVT_API int tvt_split_f32(const tvt_split_args* args, void* stream);

/* Three-plane split for the exact forward GEMMs of the fp32 parity mode: x [rows, cols] (row pitch ld) = x0 + x1 + x2
 * (bf16 planes, all 24 mantissa bits) laid out so that ONE 3-pass tvt_gemm over K' = 4 * cols computes every product of
 * order <= 2:  operand 0 (A):  hi4 = [x0 | x0 | x2 | x1], lo4 = [x1 | 0 | 0 | 0];
 *              operand 1 (B):  hi4 = [w0 | w2 | w0 | w1], lo4 = [w1 | 0 | 0 | 0];  both [rows, 4 * cols] bf16. */
typedef struct {
  const float* x;
  void* hi4;
  void* lo4;
  int64_t rows, cols, ld;
  int32_t operand;
  int32_t reserved;
} tvt_split3_args;
TVT_API int tvt_split_f32x3(const tvt_split3_args* args, void* stream);

/* y = dropout(act(x + bias)) (+ residual) elementwise over [rows, cols] fp32 input -> T output (after a split-K
 * GEMM whose epilogue cannot apply them), and its backward dx = dy * act'(y) (* dropout mask). */
typedef struct {
  const float* x;
  const float* bias; /* [cols] or NULL */
  void* y;
  int64_t rows, cols;
  int32_t out_dtype;
  int32_t act;
  float dropout_p;
  uint64_t dropout_seed;
  const void* residual; /* optional [rows, cols] of out_dtype, added after the dropout */
  void* preact;         /* optional [rows, cols] of out_dtype: x + bias before the activation (GELU backward) */
} tvt_bias_act_args;
TVT_API int tvt_bias_act_fwd(const tvt_bias_act_args* args, void* stream);

/* PositionalEncoding.forward (src/models/transformer.py:23-25): y[b, s, :] = dropout(x[b, s, :] + pe[s, :])
 * on batch-major tokens [B*S, d]; pe fp32 [S, d].  Backward is tvt_act_bwd with act = NONE. */
typedef struct {
  const void* x;
  const float* pe;
  void* y;
  int64_t rows, d, seq_len;
  int32_t dtype;
  float dropout_p;
  uint64_t dropout_seed;
} tvt_posenc_args;
TVT_API int tvt_posenc_fwd(const tvt_posenc_args* args, void* stream);

/* dx = dy * act'(.) * dropout-mask: backward of y = dropout(act(z)) when no GEMM epilogue can absorb it.
 * ReLU uses the forward output y (y > 0 covers both the ReLU and the dropped elements); GELU uses the
 * saved pre-activation z.  Elementwise over [rows, cols], all tensors of dtype `dtype`. */
typedef struct {
  const void* dy;
  const void* y_or_z;
  void* dx;
  int64_t rows, cols;
  int32_t dtype;
  int32_t act;
  float dropout_p;
  uint64_t dropout_seed;
} tvt_act_bwd_args;
TVT_API int tvt_act_bwd(const tvt_act_bwd_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Optimizer step on flat parameter / gradient buckets (the step right after
 * backward).  Replaces torch.optim.AdamW / SGD of configure_optimizers (src/models/frame_transformer.py:123-134,
 * src/models/transformer.py:58-64) with one pass over the bucket that also applies the data-parallel
 * gradient averaging (grad_scale = 1 / world) and refreshes the bf16 operand planes of the weights, so the
 * next step's GEMMs need no separate conversion pass.
 *   kind 0 (AdamW): p *= 1 - lr*wd;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
 *                   p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
 *   kind 1 (SGD):   g += wd * p;  buf = momentum * buf + g  (buf = g on the first step);  p -= lr * buf
 *   kind 2 (Adagrad, src/models/frame_transformer.py:131-133): g += wd * p;  v += g^2;  p -= lr * g / (sqrt(v) + eps)
 * All arrays fp32 of n elements; `m` is the first-moment / momentum buffer, `v` unused for SGD. */
typedef struct {
  float* p;
  const float* g;
  float* m;
  float* v;
  void* p_hi;   /* optional bf16 copy of the updated parameters */
  void* p_lo;   /* optional bf16 residual plane (fp32 parity mode) */
  int64_t n;
  int32_t kind;
  int32_t step; /* 1-based */
  float lr, beta1, beta2, eps, weight_decay, momentum, grad_scale;
  const void* step_dev; /* optional device int64: the 1-based step count read by the kernel (overrides `step`), for whole-step CUDA graphs */
} tvt_optim_step_args;
TVT_API int tvt_optim_step(const tvt_optim_step_args* args, void* stream);

/* Skinny linear layer for class heads (N = classes <= 64, not tensor-core shaped):
 *   y[m, c] = sum_k x[m, k] * w[c, k] + b[c]      (mlp_head[1], src/models/transformer.py:54;
 *                                                   Reasoning's last Linear, src/models/TPN.py:97)
 * fwd: x [M, K] (T), w [C, K] fp32, b [C] fp32 -> y [M, C] fp32.
 * bwd: dy [M, C] fp32 -> dx [M, K] (T), dw [C, K] fp32 (+=), db [C] fp32 (+=) (atomics). */
typedef struct {
  const void* x; const float* w; const float* b;
  float* y;
  int64_t m, k, classes;
  int32_t dtype;
} tvt_head_linear_fwd_args;
TVT_API int tvt_head_linear_fwd(const tvt_head_linear_fwd_args* args, void* stream);
typedef struct {
  const void* x; const float* w; const float* dy;
  void* dx; float* dw; float* db;
  int64_t m, k, classes;
  int32_t dtype;
} tvt_head_linear_bwd_args;
TVT_API int tvt_head_linear_bwd(const tvt_head_linear_bwd_args* args, void* stream);

/* CLS gather + expert sum (src/models/transformer.py:123,127-130): out[b, :] = sum_e tok_e[b*S, :];
 * and the scatter of its gradient back into zero-filled token-gradient tensors. */
#define TVT_MAX_EXPERTS 8
typedef struct {
  const void* tokens[TVT_MAX_EXPERTS]; /* each [B*S, d], CLS = row b*S */
  void* out;                           /* [B, d] */
  int64_t batch, seq_len, d;
  int32_t num_experts;
  int32_t dtype;
} tvt_cls_sum_args;
TVT_API int tvt_cls_sum_fwd(const tvt_cls_sum_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Collaborative gating fusion (src/models/collabgating.py:17-56,59-87): the
 * elementwise / row-wise work between the three stacked projection GEMMs (tvt_gemm), forward and backward.
 *   tvt_stretch_cast     F.interpolate(x, 2048) nearest-neighbour stretch of a narrow expert (collabgating.py:12-16),
 *                        y[r, i] = x[r, floor(i * d_in / d_out)], cast to the activation dtype, written at row pitch ld_out
 *   tvt_collab_mix_fwd   T_i = (E-1) C_i + sum_{j>i} C_j + sum_{j<i} PC_j  (the pairwise sums t_i = c_i + c_j of :35-43 with
 *                        the re-projected experts of :49);  _bwd:  dC_j = (E-1) dT_j + sum_{i<j} dT_i,  dPC_j = sum_{i>j} dT_i
 *   tvt_collab_gate_fwd  g = sum_i C_i * sigmoid(C_i + A_i)  (ContextGating's GLU :83-85, summed over experts :50);
 *                        _bwd: dC_i, dA_i from dg
 *   tvt_l2norm_fwd/_bwd  F.normalize of GatedEmbeddingUnit (:66-69): y = x / max(||x||, eps) per row (y, dy fp32)
 * c / a / out / dc / da: [experts, rows, d] (pc / dpc: [experts - 1, rows, d]; gate out / dout: [rows, d]), contiguous,
 * of `dtype`. */
typedef struct {
  const float* x;     /* [rows, d_in] fp32 */
  void* y;            /* [rows, ld_out] of out_dtype; columns [0, d_out) written */
  int64_t rows, d_in, d_out, ld_out;
  int32_t out_dtype;
  int32_t reserved;
} tvt_stretch_cast_args;
TVT_API int tvt_stretch_cast(const tvt_stretch_cast_args* args, void* stream);

typedef struct {
  const void* c;      /* C  */
  const void* pc;     /* PC (mix) */
  const void* a;      /* A  (gate) */
  void* out;          /* T (mix fwd) or g (gate fwd) */
  const void* dout;   /* dT (mix bwd) or dg (gate bwd) */
  void* dc;
  void* dpc;          /* mix bwd */
  void* da;           /* gate bwd */
  int64_t rows, d;
  int32_t experts;
  int32_t dtype;
} tvt_collab_args;
TVT_API int tvt_collab_mix_fwd(const tvt_collab_args* args, void* stream);
TVT_API int tvt_collab_mix_bwd(const tvt_collab_args* args, void* stream);
TVT_API int tvt_collab_gate_fwd(const tvt_collab_args* args, void* stream);
TVT_API int tvt_collab_gate_bwd(const tvt_collab_args* args, void* stream);

typedef struct {
  const void* x;      /* fwd: [rows, d] of dtype */
  float* y;           /* [rows, d] fp32: written by fwd, read by bwd */
  float* inv_norm;    /* [rows] fp32: 1 / max(||x||, eps), written by fwd, read by bwd */
  const float* dy;    /* bwd */
  void* dx;           /* bwd: [rows, d] of dtype */
  int64_t rows, d;
  float eps;
  int32_t dtype;
} tvt_l2norm_args;
TVT_API int tvt_l2norm_fwd(const tvt_l2norm_args* args, void* stream);
TVT_API int tvt_l2norm_bwd(const tvt_l2norm_args* args, void* stream);

/* Loader-side feature path on the GPU (src/dataloaders/MMX_Temporal_dl.py:167-181), one pass per expert tensor:
 *   zero-pad each [1, D_in] feature vector to D_out columns (ConstantPad1d to 2048, :167-169; d_out == d_in skips it),
 *   train-time add_transforms (:176-181): with probability p_drop the vector becomes zeros, then with probability
 *   p_noise Gaussian noise N(0, noise_std^2) is added to ALL d_out columns (the padded ones too, as in the reference),
 *   and the result is written in the activation dtype (fusing the fp32 -> bf16 cast of the embed prologue).
 * Randomness is counter-based (same hash as dropout): row decisions from hash(seed, row) words, element noise by
 * Box-Muller on hash(seed ^ stream, row * d_out / 2 + j), so the oracle reproduces every decision bit for bit. */
typedef struct {
  const float* x;        /* [rows, d_in] */
  void* y;               /* [rows, d_out], out_dtype */
  int64_t rows, d_in, d_out;   /* d_in <= d_out, both multiples of 4 */
  float p_drop, p_noise, noise_std;   /* 0.3, 0.3, sqrt(0.1) in the reference; both p = 0 at evaluation time */
  uint64_t seed;
  int32_t out_dtype;
} tvt_feature_augment_args;
TVT_API int tvt_feature_augment(const tvt_feature_augment_args* args, void* stream);

/* Evaluation read-out (src/models/transformer.py:146-158 validation_step + src/callbacks/callbacks.py:34-45): one pass
 * over a batch of logits that appends, at row `row_offset` of caller-owned running buffers of `capacity` rows,
 *   probs = sigmoid(logits)            (what the reference appends to running_logits),
 *   labels = (int)target               (target.int(), appended to running_labels),
 *   pred_bits: bit k set iff probs > thresholds[k]   (the callback's (running_logits > t) for its list of thresholds),
 *   top1 = argmax_c logits             (first maximum, torch.argmax tie rule).
 * Replaces the reference's Python lists of per-batch tensors + torch.cat + one elementwise pass per threshold. */
#define TVT_MAX_THRESHOLDS 16
typedef struct {
  const float* logits;       /* [batch, classes] */
  const void* target;        /* [batch, classes], target_dtype; may be NULL (labels untouched) */
  float* probs;              /* [capacity, classes] */
  int32_t* labels;           /* [capacity, classes] or NULL */
  uint16_t* pred_bits;       /* [capacity, classes] or NULL */
  int32_t* top1;             /* [capacity] or NULL */
  int64_t batch, classes, row_offset, capacity;
  float thresholds[TVT_MAX_THRESHOLDS];
  int32_t num_thresholds;
  int32_t target_dtype;      /* TVT_F32 or TVT_F64 */
} tvt_eval_readout_args;
TVT_API int tvt_eval_readout(const tvt_eval_readout_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TVT_H_ */
