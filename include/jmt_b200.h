/*
 * jmt_b200.h -- C-ABI of the B200-native (sm_100a) engine for the Joint Multimodal Transformer
 * data-parallel hot path: fusion forward/backward, TCN, CCC loss/metric.
 *
 * The reference (PoloWlg/Joint-Multimodal-Transformer-6th-ABAW) has no FFI: its boundary is the
 * Python nn.Module API.  The drop-in modules in
 * joint-multimodal-transformer-6th-abaw_b200/ keep that API and drive this library through
 * ctypes; every entry point below states which reference computation (file:line) it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers on the current device unless
 *     stated otherwise; `stream` is a cudaStream_t passed as void*.
 *   - functions never allocate, never synchronise, never throw; they return JMT_OK (0) or a
 *     negative jmt_status.  jmt_last_error() gives a thread-local message for the last failure.
 *   - thread-safe / re-entrant: no global mutable state except an immutable per-process table of
 *     driver entry points and per-function attribute flags.
 *   - dtype codes: JMT_F32 = 0, JMT_BF16 = 1.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns
 *     JMT_ERR_CUDA.
 */
#ifndef JMT_B200_H_
#define JMT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JMT_ABI_VERSION 9

typedef enum {
  JMT_OK = 0,
  JMT_ERR_INVALID = -1,   /* bad argument (shape, alignment, enum) */
  JMT_ERR_CUDA = -2,      /* CUDA runtime / driver error (launch, tensor-map encode, no device) */
  JMT_ERR_UNSUPPORTED = -3
} jmt_status;

enum { JMT_F32 = 0, JMT_BF16 = 1 };
enum { JMT_ACT_NONE = 0, JMT_ACT_RELU = 1, JMT_ACT_LEAKY_RELU = 2 };
enum { JMT_MAJOR_K = 0, JMT_MAJOR_MN = 1 };
enum { JMT_STORE = 0, JMT_ACCUMULATE = 1, JMT_ATOMIC_ADD = 2 };
/* CCC closed forms */
enum {
  JMT_CCC_METRIC = 0,      /* EvaluationMetrics/cccmetric.py:4-21  (population std)            */
  JMT_CCC_LOSS_LIVE = 1,   /* losses/loss.py:18-32, digitize_num==1 (unbiased std, eps in rho)  */
  JMT_CCC_LOSS_MASKED = 2, /* losses/CCCLoss.py:12-43 (y != ignore mask, pre-mask N divisor)    */
  JMT_CCC_NUMPY = 3        /* cccmetric.py:41-56 (np.cov N-1 over np.var N, +1e-8)              */
};

int jmt_abi_version(void);
const char* jmt_last_error(void);
/* number of kernels launched by this library in the calling process since load (all threads) */
int64_t jmt_launch_count(void);

/* ------------------------------------------------------------------------------------------ *
 * GEMM family.  One descriptor drives both implementations:
 *   jmt_gemm_bf16 : TMA-fed tcgen05.mma (kind::f16, bf16 operands, fp32 accumulate in TMEM),
 *                   persistent warp-specialised kernel.  a/b are bf16.
 *   jmt_gemm_f32  : fp32 FFMA tiled kernel (the high-precision parity mode).  a/b are fp32.
 *
 *   D[b](m, n) (op)= act( alpha * sum_{tap} sum_{k} A[b](m + ash(tap), k) * B[b](n, tap*K + k + ...) + bias[n] )
 *
 * Operand addressing (elements):  K-major operand X: X(r, k) = x[batch_off + r*ld + k]
 *                                 MN-major operand X: X(r, k) = x[batch_off + k*ld + r]
 *   batch b = b1*nb0 + b0, batch_off = b0*bs0 + b1*bs1.  A stride of 0 on a batch dimension (A or B) = that operand is
 *   shared by every entry of the dimension.  Batch strides may be smaller than the operand (overlapping entries): the k taps
 *   of a conv weight gradient are ONE launch with b_bs1 = dilation rows and d_bs1 = one column block of dW.
 *   `*_rows` is the number of valid rows of the 2-D matrix (m/n extent for K-major, k extent for
 *   MN-major); rows outside [0, rows) read as zero (TMA out-of-bounds fill) -- this is how the
 *   causal left padding of the TCN (temporal_convolutional_model.py:12-18,24-27) and ragged tile
 *   tails are realised without materialising padding.
 *   Taps (implicit-GEMM dilated conv): for tap j the A row index is shifted by
 *   a_shift0 + j*a_shift_step (K-major A: shifts m; MN-major A: shifts k), the B row index by
 *   b_shift0 + j*b_shift_step when B is MN-major, and for K-major B the k offset advances by
 *   j*K (weights laid out (n, tap*K + k)).
 *   reduce_batch != 0: the batch index becomes an extra reduction dimension (D is not batched).
 *   split_k > 1 requires store_mode == JMT_ATOMIC_ADD and d_dtype == JMT_F32.
 *   colmask: see the field comment (channel dropout fused into the epilogue).
 *
 * Replaces: every nn.Linear / in_proj / out_proj / bmm of the fusion path
 * (mm_multi_transformers.py:52-56,120-124,142-167,203-209; two_transformers.py:104-128;
 * torch F.multi_head_attention_forward's addmm/bmm) and nn.Conv1d of the TCN
 * (temporal_convolutional_model.py:24-41), forward, dgrad and wgrad.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const void* a; const void* b; void* d;
  const float* bias;            /* length N, nullable */
  int32_t a_major, b_major;     /* JMT_MAJOR_* */
  int32_t M, N, K;              /* K = reduction extent per tap (per batch) */
  int32_t a_rows, b_rows;       /* valid rows of each operand's 2-D matrix (see above) */
  int64_t a_ld, a_bs0, a_bs1;
  int64_t b_ld, b_bs0, b_bs1;
  int64_t d_ld, d_bs0, d_bs1;   /* D(m, n) = d[batch_off + m*d_ld + n] */
  int32_t nb0, nb1;             /* batch extents (>= 1) */
  int32_t d_dtype;              /* JMT_F32 / JMT_BF16 */
  int32_t act;                  /* JMT_ACT_* applied after bias */
  float alpha, slope;
  int32_t store_mode;           /* JMT_STORE / JMT_ACCUMULATE / JMT_ATOMIC_ADD */
  int32_t ntaps;                /* >= 1 */
  int32_t a_shift0, a_shift_step, b_shift0, b_shift_step;
  int32_t reduce_batch;
  int32_t split_k;              /* >= 1 */
  /* optional per-(batch, column) keep-mask applied AFTER the activation: D(m, n) *= colmask[b*N + n] ? colmask_scale : 0
   * (nn.Dropout2d of the TCN fused into the conv epilogue, temporal_convolutional_model.py:29,36); NULL = none.
   * Needs reduce_batch == 0 and split_k == 1. */
  const uint8_t* colmask;
  float colmask_scale;
  /* Flat TCN layout (all sequences stacked in ONE (N*(pad+L), C) matrix, `pad` zero rows in front of every sequence =
   * the causal left padding of the next conv and the right padding of the previous sequence's dgrad, so a 128-row tile
   * is never 56 % empty as with M = L = 300 per batch entry):
   *   colmask_row_period > 0: the keep-mask row is  m / colmask_row_period  instead of the batch index (nb0 = nb1 = 1);
   *                           must be >= 127 (a 128-row tile stages the keep-flags of at most two samples; shorter sequences:
   *                           run the GEMM without colmask and apply the mask with jmt_apply_mask);
   *   zero_row_period > 0:    output rows with (m % zero_row_period) < zero_row_count contribute zeros (stored as 0 /
   *                           nothing added): the padding rows must stay zero for the next layer. */
  int32_t colmask_row_period;
  int32_t zero_row_period, zero_row_count;
  /* Backward epilogue extensions (jmt_gemm_bf16 only; bf16 D with 16-byte aligned geometry, act == JMT_ACT_NONE, split_k == 1,
   * reduce_batch == 0), so that the activation-gradient pass and the bias-gradient column sums need no kernel of their own:
   *   epi_aux  (nullable, bf16, SAME element offsets as D): the stored / added value is multiplied by act'(aux) =
   *            (aux > 0 ? 1 : aux_slope) -- aux is the saved forward output y = act(...) of the layer whose input gradient this
   *            GEMM produces (ReLU: slope 0, mm_multi_transformers.py:52-56; LeakyReLU 0.01, temporal_convolutional_model.py:28,35);
   *   d_colsum (nullable, fp32): d_colsum[b0 * colsum_bs0 + n] += sum over rows m (and b1) of the values this launch stores / adds
   *            to D(m, n) (fp32 atomics; the caller zeroes it) = this launch's share of the bias gradient of the Linear / conv
   *            that produced D's forward counterpart. */
  const void* epi_aux;
  float aux_slope;
  float* d_colsum;
  int64_t colsum_bs0;
} jmt_gemm_desc;

int jmt_gemm_bf16(const jmt_gemm_desc* g, void* stream);
int jmt_gemm_f32(const jmt_gemm_desc* g, void* stream);
/* "bf16x3" high-precision tensor-core mode (SURVEY 7 hard part 4; the 1e-3 / 1e-4 parity gate of north_star): both
 * operands are fp32 values split into two bf16 matrices of IDENTICAL geometry, x = hi + lo (jmt_split_bf16x2); g->a / g->b
 * point at the hi parts, a_lo / b_lo at the lo parts.  Every k-step issues three tcgen05.mma into the same fp32 TMEM
 * accumulator: A_hi B_hi + A_hi B_lo + A_lo B_hi (16 operand mantissa bits; the dropped lo x lo term is 2^-16 relative).
 * Everything else (taps, batches, epilogue, store modes) as jmt_gemm_bf16. */
int jmt_gemm_bf16x3(const jmt_gemm_desc* g, const void* a_lo, const void* b_lo, void* stream);
/* hi[i] = bf16(x[i]), lo[i] = bf16(x[i] - hi[i]) over n contiguous fp32 values (all pointers 16-byte aligned) */
int jmt_split_bf16x2(const float* x, void* hi, void* lo, int64_t n, void* stream);
/* Debug aid (no reference counterpart): when dev_buf != NULL (device buffer of 148*16 uint64) every later
 * jmt_gemm_bf16 launch overwrites per-CTA cycle counters of its TMA / MMA / epilogue roles; NULL disables. */
int jmt_gemm_set_profile_buffer(void* dev_buf);
/* B-stationary scheduling of jmt_gemm_bf16 for short-K linears (K <= 512, N % 256 == 0, one tap, no batch): every CTA pair keeps
 * its 256-column slice of B resident in shared memory and streams only A (csrc/gemm_tc.cu, TcParams::bres).  Results are
 * independent of the mode.  mode: -1 = environment (JMT_GEMM_BRES, default automatic), 0 = off, 1 = automatic (large M only),
 * 2 = whenever the geometry allows (tests).  Returns the previous mode. */
int jmt_gemm_set_bres_mode(int mode);

/* ------------------------------------------------------------------------------------------ *
 * Fused attention core (bf16, tcgen05): two chained GEMMs with the row-wise softmax algebra between them.
 *   mode 0 (forward):  X = P  = softmax(scale * A1 B1^T) (rows of A1 = queries, rows of B1 = keys);  D = X B2
 *                      A1 = Q, B1 = K, B2 = V;  X (the probabilities) is also written to `x` for backward.
 *   mode 1 (backward): X = dS = scale * P o (A1 B1^T - delta), delta = rowsum(dO o O) = rowsum(P o dP), P from p_in;  D (+)= X B2
 *                      A1 = dO, B1 = V, B2 = K  ->  D = dQ;  X (= dS) is written to `x` (dK = dS^T Q is a plain GEMM).
 * Operand (r, k) of (head h, batch b) = ptr[b*bs + h*hs + r*ld + k], k < dh contiguous; X / p_in are
 * (NB, heads, Lq, x_ld) contiguous bf16 with x_ld % 8 == 0.  Supported: dh in {64,128,256,512}, S <= 512 subject to
 * shared memory (S <= 320 at dh = 512); jmt_attn_chain_supported() says whether a geometry is (the caller
 * otherwise composes the same attention from jmt_gemm_bf16 + jmt_softmax_*).
 * Replaces torch's MHA math path (bmm, softmax, bmm) at the nn.MultiheadAttention call sites
 * mm_multi_transformers.py:62,142-167 and mm_transformers.py:76,125-134, forward and backward.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const void* a1; const void* b1; const void* b2;   /* bf16 */
  const void* p_in;                                  /* mode 1: saved probabilities; else NULL */
  const void* o_in;                                  /* mode 1: saved forward output O, same geometry as a1 (= dO); else NULL */
  const float* delta_in;                             /* mode 1: precomputed delta (NB, heads, Lq) fp32 (jmt_rowdot_bf16), or NULL: from o_in;
                                                        both NULL: delta = rowsum(P o dP) inside the kernel (no separate pass over dO / O) */
  void* x;                                           /* out: P (mode 0) / dS (mode 1), bf16; mode 0 with D: may be NULL (forward-only callers: the
                                                        probabilities are then never written to memory) */
  void* d;                                           /* out: O (mode 0) / dQ (mode 1), bf16; NULL: stop after X (P / dS) -- GEMM2 is then a plain jmt_gemm_bf16 */
  int32_t mode;
  int32_t Lq, S, dh, heads, NB;
  int64_t a1_ld, a1_hs, a1_bs;
  int64_t b1_ld, b1_hs, b1_bs;
  int64_t b2_ld, b2_hs, b2_bs;
  int64_t d_ld, d_hs, d_bs;
  int64_t x_ld;
  float scale;
  int32_t store_mode;                                /* JMT_STORE / JMT_ACCUMULATE for D */
  float* lse_out;                                    /* mode 0, nullable: logsumexp_j(scale * q_i . k_j) per query row, (NB, heads, Lq) fp32 */
} jmt_attn_desc;
int jmt_attn_chain_supported(const jmt_attn_desc* g);   /* 1 / 0, no launch */
int jmt_attn_chain_bf16(const jmt_attn_desc* g, void* stream);
/* out[(b*heads + h)*rows + r] = sum_d a[b,h,r,d] * b[b,h,r,d] (element (r, d) of (h, b) at b*bs + h*hs + r*ld + d): the
 * delta_i = dO_i . O_i = sum_j P_ij dP_ij term of the softmax backward, one pass over dO and O. */
int jmt_rowdot_bf16(const void* a, const void* b, int64_t ld, int64_t hs, int64_t bs, int NB, int heads, int rows, int dh,
                    float* out, void* stream);
/* Attention over long key sequences (S beyond what one jmt_attn_chain_bf16 tile holds on chip; SURVEY 7.6a, the batch-dimension
 * attention of MultimodalTransformer_wo_JR at large batch, mm_transformers.py:120-122): the caller runs the forward chain kernel
 * over key CHUNKS (x = NULL, lse_out set) and merges with the running log-sum-exp -- flash-attention's algebra across launches,
 * no (L, S) score or probability tensor ever exists in memory:
 *   lse' = logaddexp(lse_acc, lse_chunk);  acc' = acc * exp(lse_acc - lse') + o_chunk * exp(lse_chunk - lse')
 * o_chunk bf16, element (r, d) of (head h, batch b) at b*bs + h*hs + r*ld + d; acc fp32 with the same offsets; lse_* (NB, heads, rows)
 * fp32.  first != 0: acc / lse_acc are initialised from the chunk.  out != NULL (last chunk): the bf16 result goes to `out` (same
 * offsets) instead of acc. */
int jmt_attn_merge(const void* o_chunk, int64_t ld, int64_t hs, int64_t bs, const float* lse_chunk, float* acc, float* lse_acc,
                   void* out, int first, int NB, int heads, int rows, int dh, void* stream);
/* Debug aid: per-CTA cycle counters of the TMA / MMA / row-warp roles (148*16 uint64 device buffer; NULL disables). */
int jmt_attn_set_profile_buffer(void* dev_buf);

/* ------------------------------------------------------------------------------------------ *
 * Attention backward, the three GEMMs behind dS / P in ONE launch (bf16, tcgen05, CTA pairs):
 *   part 0 (x_trans = 0):  dQ(q, c) (op)= alpha * sum_s dS(q, s) * K(s, c)
 *   part 1 (x_trans = 1):  dK(s, c) (op)= alpha * sum_q dS(q, s) * Q(q, c)
 *   part 2 (x_trans = 1):  dV(s, c) (op)= alpha * sum_q P(q, s)  * dO(q, c)
 * i.e. every part is  d(n, c) (op)= alpha * sum_k X(n, k) * a(k, c)  with X(n, k) = x[n*x_ld + k] (x_trans = 0: n < Lq, k < S) or
 * x[k*x_ld + n] (x_trans = 1: n < S, k < Lq); x = dS or P, (NB, heads, Lq, x_ld) contiguous bf16 as written by
 * jmt_attn_chain_bf16; a / d element (r, c) of (head h, batch b) at ptr[b*bs + h*hs + r*ld + c], c < dh.
 * A part with d == NULL is skipped.  store_mode: JMT_STORE / JMT_ACCUMULATE.  colsum (nullable, fp32, heads*dh entries,
 * caller-zeroed or running): colsum[h*dh + c] += sum over (b, n) of the values this launch stores / adds to d(n, c) -- the bias
 * gradient of the projection that produced Q / K / V.
 * Computed transposed on the tensor cores (head dimension on the MMA's M axis, the sequence extent on N: see
 * csrc/attn_bwd_tc.cu).  Supported: dh in {256, 512}, Lq, S <= 320, heads*dh <= 1024, x_ld % 8 == 0, 16-byte aligned a / d
 * geometry; jmt_attn_bwd_dqkv_supported() says whether a geometry is (the caller otherwise issues three jmt_gemm_bf16).
 * Replaces the two bmm backward pairs of torch's MHA math path at the nn.MultiheadAttention call sites
 * mm_multi_transformers.py:62,142-167.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const void* a;                  /* bf16 (k extent, dh) matrix contracted over its rows: K (dQ), Q (dK), dO (dV) */
  int64_t a_ld, a_hs, a_bs;
  const void* x;                  /* bf16 dS or P */
  int32_t x_trans;
  void* d;                        /* bf16 output (n extent, dh); NULL: part skipped */
  int64_t d_ld, d_hs, d_bs;
  int32_t store_mode;
  float alpha;
  float* colsum;
} jmt_attn_bwd_part;
typedef struct {
  jmt_attn_bwd_part part[3];
  int32_t Lq, S, dh, heads, NB;
  int64_t x_ld;
} jmt_attn_bwd_desc;
int jmt_attn_bwd_dqkv_supported(const jmt_attn_bwd_desc* g);   /* 1 / 0, no launch */
int jmt_attn_bwd_dqkv_bf16(const jmt_attn_bwd_desc* g, void* stream);
/* Debug aid: per-CTA cycle counters (148*16 uint64 device buffer; NULL disables): [0] MMA wait-A [1] wait-X [2] wait-TMEM
 * [3] MMA total [4] epilogue wait-accumulator [5] epilogue total */
int jmt_attn_bwd_set_profile_buffer(void* dev_buf);

/* ------------------------------------------------------------------------------------------ *
 * Memory-bound fused row kernels (one warp per row, warp-shuffle reductions, 16-byte accesses).
 * `dtype` is the activation dtype of x/out tensors; statistics and parameters are fp32.
 * ------------------------------------------------------------------------------------------ */
/* F.normalize(x, dim=-1): two_transformers.py:118-119.  in: (rows, D) in_dtype (row pitch in_ld);
 * out: (rows, D) out_dtype; inv_norm (rows) fp32 saved for backward (nullable). */
int jmt_l2norm_fwd(const void* x, int in_dtype, int64_t in_ld, void* out, int out_dtype, int64_t rows,
                   int D, float eps, float* inv_norm, void* stream);
/* dx = r*(dy - y*(y.dy)) (r = inv_norm; rows with ||x|| < eps: dx = r*dy).  y = saved output. */
int jmt_l2norm_bwd(const void* dy, const void* y, int dtype, const float* inv_norm, float eps,
                   void* dx, int dx_dtype, int64_t rows, int D, void* stream);
/* The same on the TCN's flat padded layout (I3DWSDDA.py:44 `temporal(x).transpose(1, 2)` followed by two_transformers.py:118): the
 * input of the forward / the dx of the backward has seq_rows = row0 + seq_len physical rows per sequence, the first row0 of them
 * padding; out / dy / y / inv_norm are compact (nseq*seq_len rows).  The backward writes the padding rows of dx as zeros.  Saves the
 * un-padding copy and, in the backward, the zero fill + scatter copy of the padded gradient. */
int jmt_l2norm_fwd_seq(const void* x, int in_dtype, int64_t in_ld, void* out, int out_dtype, int64_t nseq, int seq_len,
                       int seq_rows, int row0, int D, float eps, float* inv_norm, void* stream);
int jmt_l2norm_bwd_seq(const void* dy, const void* y, int dtype, const float* inv_norm, float eps, void* dx, int dx_dtype,
                       int64_t nseq, int seq_len, int seq_rows, int row0, int D, void* stream);

/* y = LayerNorm(x + res) * gamma + beta  (post-LN residual: mm_multi_transformers.py:62-69).
 * res nullable.  Saves mean/rstd (rows) fp32. */
int jmt_add_layernorm_fwd(const void* x, const void* res, const float* gamma, const float* beta, float eps,
                          void* y, float* mean, float* rstd, int64_t rows, int D, int dtype, void* stream);
/* dz = d(x+res); dgamma/dbeta (D) fp32 are ACCUMULATED with atomics (caller zeroes).
 * dz_accumulate != 0: dz += result.  dz_colsum (D, nullable): += column sums of dz = the bias gradient of the
 * Linear that produced the residual branch (out_proj / feed_forward.2), saving a separate pass over dz. */
int jmt_add_layernorm_bwd(const void* dy, const void* x, const void* res, const float* gamma,
                          const float* mean, const float* rstd, void* dz, int dz_accumulate, float* dgamma,
                          float* dbeta, float* dz_colsum, int64_t rows, int D, int dtype, void* stream);

/* P = softmax(S) over the last dim (torch MHA math path, keys axis).  S fp32 (rows, s_ld),
 * P p_dtype (rows, p_ld); columns >= cols of P are written as zero up to p_ld. */
int jmt_softmax_fwd(const float* s, int64_t s_ld, void* p, int p_dtype, int64_t p_ld, int64_t rows, int cols,
                    void* stream);
/* dS = P o (dP - rowsum(dP o P)); dP fp32 (rows, dp_ld); dS ds_dtype (rows, ds_ld), pad zeroed. */
int jmt_softmax_bwd(const void* p, int p_dtype, int64_t p_ld, const float* dp, int64_t dp_ld, void* ds,
                    int ds_dtype, int64_t ds_ld, int64_t rows, int cols, void* stream);

/* Tiny-sequence attention (L = S <= 8): Intra_modal_transformer_fusion (L=2,
 * intra_modal_transformer_fusion.py:93-108) and the SELF_ATTEN head (L=6,
 * mm_multi_transformers.py:173-193).  qkv: (L, N, 3E) packed projections (dtype), out: (L, N, E);
 * one warp per (n, head), everything in registers.  probs (N*h, L, L) fp32 saved for backward. */
int jmt_attn_small_fwd(const void* qkv, void* out, float* probs, int L, int64_t N, int E, int heads,
                       float scale, int dtype, void* stream);
int jmt_attn_small_bwd(const void* qkv, const void* dout, const float* probs, void* dqkv, int L, int64_t N,
                       int E, int heads, float scale, int dtype, void* stream);

/* Regressor tail: out[g][b*sb + t*st] = h[g][m,:128] . w[g] + b[g], row m = b*T + t, G <= 4 groups
 * (two_transformers.py:104-114: Linear(128,1) of vregressor/aregressor; :146-149: Linear(128,2)),
 * fused with squeeze and the (B,T) / (T,B) output layout.  h: hidden after ReLU (and
 * dropout), dtype `dtype`, row pitch h_ld.  Pointer arrays are HOST arrays of device pointers. */
int jmt_regressor_tail_fwd(int G, const void* const* h, int64_t h_ld, int dtype, const float* const* w,
                           const float* const* b, float* const* out, int64_t M, int64_t T, int64_t sb, int64_t st,
                           void* stream);
/* dh[g][m,j] (+)= (h>0) * dout[g][m] * w[g][j] * scale[g]; dw[g] += sum_m dout*h; db[g] += sum_m dout
 * (fp32 atomics; caller zeroes dw/db).  accumulate[g] != 0: add into dh[g] (hidden shared by groups). */
int jmt_regressor_tail_bwd(int G, const void* const* h, int64_t h_ld, int dtype, const float* const* w,
                           const float* const* dout, void* const* dh, const int* accumulate, const float* scale,
                           float* const* dw, float* const* db, int64_t M, int64_t T, int64_t sb, int64_t st,
                           void* stream);

/* ------------------------------------------------------------------------------------------ *
 * Elementwise / column reductions
 * ------------------------------------------------------------------------------------------ */
/* dx = dy * (y > 0 ? 1 : slope)   (ReLU: slope 0; LeakyReLU 0.01), in place allowed. */
int jmt_act_bwd(const void* dy, const void* y, void* dx, int64_t n, float slope, int dtype, void* stream);
/* out[c] += sum_r x[r*ld + c]  (bias gradients), fp32 atomics; caller zeroes. */
int jmt_colsum(const void* x, int dtype, int64_t ld, int64_t rows, int cols, float* out, void* stream);
/* One pass for the backward of Linear/Conv1d -> (Leaky)ReLU -> channel dropout (two_transformers.py:104-114 ReLU heads,
 * mm_multi_transformers.py:52-56 FFN, temporal_convolutional_model.py:24-36):
 *   dx[r,c] = (mask ? (mask[r / L, c] ? dy*scale : 0) : dy) * (y[r,c] > 0 ? 1 : slope);  colsum[c] += sum_r dx[r,c]
 * dy/y/dx contiguous (rows, cols), cols % 8 == 0; mask (rows / L, cols) uint8 nullable; colsum fp32 nullable. */
int jmt_act_bwd_fused(const void* dy, const void* y, const uint8_t* mask, void* dx, int64_t rows, int cols, int L,
                      float scale, float slope, float* colsum, int dtype, void* stream);
/* Backward of out = act(a + b) fused with the backward of a = act2(pre) [* channel keep-mask * scale] (the TemporalBlock residual
 * LeakyReLU together with conv2's LeakyReLU + Dropout2d, temporal_convolutional_model.py:54-57, 35-36):
 *   dz = dy * act'(out)  (gradient of both a and b);  dz2 = dz * [mask(r / L, c) ? scale : 0] * act2'(a);  colsum[c] += sum_r dz2(r, c).
 * One pass instead of jmt_act_bwd + jmt_act_bwd_fused; results are bit-identical to that sequence.  cols % 8 == 0, 16-byte aligned. */
int jmt_add_act_bwd_fused(const void* dy, const void* out, const void* a, const uint8_t* mask, void* dz, void* dz2,
                          int64_t rows, int cols, int L, float scale, float slope, float slope2, float* colsum, int dtype,
                          void* stream);
/* out = cast(in) */
int jmt_cast(const void* in, int in_dtype, void* out, int out_dtype, int64_t n, void* stream);
/* dst[e][i] = bf16(src[e][i]), i < n[e], for `count` tensors in one launch (HOST arrays of device pointers, 16-byte aligned):
 * the bf16 operand copies of all weight matrices of a module, refreshed once per optimizer step */
int jmt_cast_multi(int count, const float* const* src, void* const* dst, const int64_t* n, void* stream);
/* y += a*x (same dtype) */
int jmt_axpy(const void* x, void* y, float a, int64_t n, int dtype, void* stream);
/* strided 2-D copy with cast: out[r*out_ld + c] = in[r*in_ld + c] */
int jmt_copy2d(const void* in, int in_dtype, int64_t in_ld, void* out, int out_dtype, int64_t out_ld,
               int64_t rows, int cols, void* stream);
/* batched transpose of the two inner dims with cast: in (nb, R, C) -> out (nb, C, R)
 * ((N,C,L) <-> channels-last (N,L,C) at the TCN boundary, I3DWSDDA.py:44) */
int jmt_transpose(const void* in, int in_dtype, void* out, int out_dtype, int64_t nb, int R, int C, void* stream);
/* same with explicit batch strides in elements (0 = dense R*C): scatters into / gathers from the flat padded TCN layout */
int jmt_transpose_strided(const void* in, int in_dtype, int64_t in_bs, void* out, int out_dtype, int64_t out_bs, int64_t nb,
                          int R, int C, void* stream);
/* out[b*out_bs + r*cols + c] = cast(in[b*in_bs + r*cols + c])  (nb, rows, cols) blocks: pad / unpad of the flat layout */
int jmt_copy_rows3d(const void* in, int in_dtype, int64_t in_bs, void* out, int out_dtype, int64_t out_bs, int64_t nb,
                    int64_t rows, int cols, void* stream);
/* out[i0*os0 + i1*os1 + c] = cast(in[i0*is0 + i1*is1 + c]), i0 < n0, i1 < n1, c < cols: row permutation with cast in one
 * launch -- the (B, T, D) -> (T, B, D) layout MultimodalTransformer_w_JR's FC head returns (mm_multi_transformers.py:201-211,
 * SURVEY Q1) and its gradient's way back. */
int jmt_copy3d(const void* in, int in_dtype, int64_t is0, int64_t is1, void* out, int out_dtype, int64_t os0, int64_t os1,
               int64_t n0, int64_t n1, int cols, void* stream);
/* max over time of channels-last sequences without materialising `transpose(1,2)` (I3DWSDDA.py:44 then
 * tsav.py:216 `torch.max(ft, 1)`): x element (n, t, c) at x[n*batch_stride + t*C + c], t < L; out (nb, C) in `dtype`,
 * arg (nb, C) = first t attaining the maximum; backward scatters dout to those positions of a pre-zeroed dx. */
int jmt_time_max_fwd(const void* x, int64_t batch_stride, int64_t nb, int L, int C, void* out, int32_t* arg, int dtype, void* stream);
int jmt_time_max_bwd(const void* dout, const int32_t* arg, int64_t batch_stride, int64_t nb, int C, void* dx, int dtype, void* stream);
/* out = LeakyReLU(a + b)  (TemporalBlock.forward, temporal_convolutional_model.py:54-57) */
int jmt_add_act(const void* a, const void* b, void* out, int64_t n, int act, float slope, int dtype, void* stream);
/* channel dropout (nn.Dropout2d on (N,C,L)) / element dropout with an explicit
 * keep-mask: x (nb, L, C) channels-last; mask (nb, C) or (nb*L*C) uint8; scale = 1/(1-p). */
int jmt_apply_mask(const void* x, const uint8_t* mask, void* out, int64_t nb, int L, int C, int per_channel,
                   float scale, int dtype, void* stream);
/* Philox-4x32-10 keep-mask generator: mask[i] = uniform(seed, offset+i) >= p */
int jmt_dropout_mask(uint8_t* mask, int64_t n, float p, uint64_t seed, uint64_t offset, const uint64_t* dev_state, void* stream);
/* dev_state (nullable): device uint64[2] = (seed increment, Philox counter offset) added to seed / offset inside the kernel,
 * so that a captured CUDA graph (which bakes the host arguments in) draws fresh masks on every replay;
 * jmt_rng_advance bumps the counter offset by `inc` in stream order (once per forward, after its last mask). */
int jmt_rng_advance(uint64_t* dev_state, uint64_t inc, void* stream);

/* weight_norm (legacy, dim=0): w[co,:] = g[co] * v[co,:] / ||v[co,:]||  (Cout rows of `inner`
 * = Cin*k elements; temporal_convolutional_model.py:24-33).  Writes w re-laid out for the
 * implicit GEMM as (Cout, k*Cin) [tap-major] in out_dtype plus its transpose-per-tap for dgrad
 * (k*... see DESIGN.md); norm (Cout) fp32 saved. */
int jmt_weight_norm_fwd(const float* g, const float* v, void* w_fwd, void* w_dgrad, int out_dtype,
                        float* norm, int cout, int cin, int k, void* stream);
/* The same for n <= 16 convolutions in ONE launch each (every conv of a TemporalConvNet per pass: the per-conv kernels are
 * latency-bound).  Arrays are HOST arrays of length n; w_dgrad (or single entries of it) may be NULL. */
int jmt_weight_norm_fwd_batched(int n, const float* const* g, const float* const* v, void* const* w_fwd, void* const* w_dgrad,
                                int out_dtype, float* const* norm, const int* cout, const int* cin, int k, void* stream);
/* dw_fwd[e] == NULL: conv e received no gradient, its dg / dv are left untouched */
int jmt_weight_norm_bwd_batched(int n, const float* const* dw_fwd, const float* const* g, const float* const* v,
                                const float* const* norm, float* const* dg, float* const* dv, const int* cout, const int* cin,
                                int k, void* stream);
/* dv, dg from dw_fwd ((Cout, k*Cin) fp32, tap-major): dg = (dw.v)/||v||, dv = g/||v|| (dw - v (dw.v)/||v||^2) */
int jmt_weight_norm_bwd(const float* dw_fwd, const float* g, const float* v, const float* norm, float* dg,
                        float* dv, int cout, int cin, int k, void* stream);

/* ------------------------------------------------------------------------------------------ *
 * CCC statistics: single-pass six-sum reduction (N, Sx, Sy, Sxy, Sxx, Syy) in fp64.
 * Replaces losses/loss.py:18-32, losses/CCCLoss.py:12-43, EvaluationMetrics/cccmetric.py:4-56.
 * ------------------------------------------------------------------------------------------ */
/* x, y: npairs rows of n fp32 values (row pitch `stride`); sums (npairs, 6) fp64, ACCUMULATED
 * (caller zeroes) so partial shards / ranks can be combined; use_ignore: skip y == ignore. */
int jmt_ccc_sums(const float* x, const float* y, int64_t n, int npairs, int64_t stride, int use_ignore,
                 float ignore, double* sums, void* stream);
/* device-side finaliser: value[p] (fp32) = the CCC variant `kind`; coef (npairs, 4) fp64 =
 * (c0, cx, cy, valid) such that d value / d x_i = c0 + cx*x_i + cy*y_i on non-ignored elements.
 * n_all = pre-mask length (JMT_CCC_LOSS_MASKED). */
int jmt_ccc_finalize(const double* sums, int npairs, int kind, double n_all, double eps, float* value,
                     double* coef, void* stream);
/* dx[p][i] = gout[p or 0] * (c0 + cx*x + cy*y), zero where ignored. gout: device fp32. */
int jmt_ccc_bwd(const float* x, const float* y, int64_t n, int npairs, int64_t stride, const double* coef,
                const float* gout, int gout_per_pair, int use_ignore, float ignore, float* dx, void* stream);
/* the reference's only padding mask: mask[i] = (y[i] != ignore)  (losses/CCCLoss.py:19) */
int jmt_label_mask(const float* y, int64_t n, float ignore, uint8_t* mask, void* stream);
/* zero-left-pad copy of the collate (padSequence.py:14-21): out (rows, out_w) zero-filled, then
 * in (rows, in_w) copied right-aligned. fp32. */
int jmt_pad_right_align(const float* in, int64_t rows, int in_w, float* out, int out_w, void* stream);

/* ------------------------------------------------------------------------------------------ *
 * Validation post-processing on the device (val.py:313-382).
 * jmt_valpost_scatter: one batch of per-frame predictions (element e = b*T + t, fp32) into the per-video
 *   arrays (all videos concatenated, `offsets` = prefix sums of the lengths): skipped when either label is
 *   `ignore` (-5) or the 1-based frame id is outside [1, length]; within a batch the element that comes last in
 *   the reference's loop order wins, across batches the larger `seq` (1, 2, ...) wins.  `stamp` (total frames,
 *   zero-initialised) is the arbitration state.  Untouched frames keep the caller's initial (0, 0).
 * jmt_valpost_finalize: per video clip to [-1,1] + uniform_filter1d(size_v / size_a, mode='constant'), optional
 *   smoothed outputs, and sums (2, 6) fp64 += (N, Sx, Sy, Sxy, Sxx, Syy) for valence / arousal (finalise with
 *   jmt_ccc_finalize(JMT_CCC_METRIC)).
 * ------------------------------------------------------------------------------------------ */
int jmt_valpost_scatter(const float* v, const float* a, const float* lab_v, const float* lab_a, const int32_t* frame_id,
                        const int32_t* video, int64_t n, const int64_t* offsets, float ignore, uint64_t seq,
                        uint64_t* stamp, float* pred_v, float* pred_a, float* label_v, float* label_a, void* stream);
int jmt_valpost_finalize(const float* pred_v, const float* pred_a, const float* label_v, const float* label_a,
                         const int64_t* offsets, int videos, int64_t total_frames, int size_v, int size_a,
                         float* smooth_v, float* smooth_a, double* sums, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* JMT_B200_H_ */
