/* onebit.h - C ABI of libonebit.so: the B200 (sm_100a) kernels behind the quantised linear layer.
 *
 * This is the drop-in boundary for ONE path of y00njaekim/CMU-11785-IDL-1.58bit-ASR: the layer
 * `QuantizedLinear` of onebit_asr/quant.py (the reference's "BitLinear") and nothing else.  The
 * reference is pure Python/PyTorch and has no FFI of its own; the entry points below are what a
 * ctypes binding of that file would call, one per step of quant.py (citations on each entry,
 * relative to /root/reference).  INTEGRATION.md shows the binding.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer (cudaMalloc'ed, 16-byte aligned) owned by the caller; the
 *    library never allocates, frees or keeps a pointer after the call returns;
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*); no host sync is made;
 *  - return value: OB_OK or an OB_ERR_* code; ob_last_error_string() describes the last failure of
 *    the calling thread.  No C++ exception crosses this boundary;
 *  - `alpha` is a device scalar.  alpha_mode OB_ALPHA_RAW: the layer parameter, the kernels use
 *    alpha_eff = |alpha| + 1e-8 (quant.py:124) and chain sign(alpha) into its gradient;
 *    OB_ALPHA_EFF: the value is used as is (the free function quantize_weight, quant.py:95);
 *  - bitwidth is 1 (binary {-1,+1}) or 2 (ternary {-1,0,+1}); 32 never reaches the library
 *    (quant.py:121-122 bypasses the quantiser);
 *  - shapes: W is [N, K] row-major (out_features x in_features), x is [M, K], y is [M, N].
 *    The tensor-core entry points need K % 64 == 0 and N % 64 == 0; M is free.
 *
 * Packed 2-bit weight format (ours; the reference keeps fp32 W_hat): field 0b00 = 0, 0b10 = -1,
 * 0b11 = +1; 16 consecutive codes along the contraction axis share one little-endian 32-bit word;
 * code t of a group sits at bit
 *      OB_ORDER_I8   : 4*(t&3) + 2*((t>>2)&1) + 16*(t>>3)     (forward GEMM operand,  [N, K/4] bytes)
 *      OB_ORDER_BF16 : 8*(t&1) + 2*((t>>1)&3) + 16*(t>>3)     (grad_x GEMM operand,   [K, N/4] bytes)
 * so that one PRMT (byte permute) per output word expands a word in shared memory.
 */
#ifndef ONEBIT_H_
#define ONEBIT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OB_VERSION 100

#define OB_OK 0
#define OB_ERR_ARG 1       /* bad shape, alignment, dtype tag, bitwidth or null pointer */
#define OB_ERR_CUDA 2      /* a CUDA runtime / driver call or a launch failed */
#define OB_ERR_ARCH 3      /* the current device is not sm_100 */
#define OB_ERR_WORKSPACE 4 /* caller-provided workspace too small */

#define OB_F32 0
#define OB_BF16 1

#define OB_ALPHA_EFF 0
#define OB_ALPHA_RAW 1

#define OB_ORDER_I8 0
#define OB_ORDER_BF16 1

typedef void* ob_stream_t; /* cudaStream_t */

int ob_version(void);
/* number of kernels this library has launched in this process (bench.py reports it as gpu_launches) */
int64_t ob_launch_count(void);
const char* ob_last_error_string(void);

/* mean(|W|) over n elements -> out[0]; deterministic two-stage reduction.  Replaces the one-off
 * alpha initialisation `weight.abs().mean()` (quant.py:111-113) and serves as the BitNet absmean
 * re-scaler.  ws: at least ob_absmean_workspace_bytes() bytes. */
size_t ob_absmean_workspace_bytes(void);
int ob_weight_absmean(const float* W, int64_t n, float* out, void* ws, ob_stream_t stream);

/* Wa = W/alpha_eff (IEEE division), codes per quant.py:49-60, packed 4 per byte.
 * packed_i8   [N, K/4] in OB_ORDER_I8   (required);
 * packed_t    [K, N/4] in OB_ORDER_BF16 (the transposed copy grad_x needs; may be NULL). */
int ob_weight_quant_pack(const float* W, const float* alpha, int alpha_mode, int N, int K,
                         int bitwidth, uint8_t* packed_i8, uint8_t* packed_t, ob_stream_t stream);

/* The same for MANY layers and BOTH bitwidths in one launch.  descs_dev: device array of `count` descriptors sorted by
 * tile0 = number of 64 x 64 tiles of all earlier layers (N and K multiples of 64); total_tiles = sum of (N/64)*(K/64).
 * Each W is read once; any of the four outputs may be NULL.  Codes are bit-identical to ob_weight_quant_pack. */
typedef struct ob_pack_desc {
  const float* W;          /* [N, K] fp32 */
  const float* alpha;      /* device scalar */
  uint8_t* packed2;        /* [N, K/4] OB_ORDER_I8, 2-bit (ternary) codes */
  uint8_t* packed2_t;      /* [K, N/4] OB_ORDER_BF16 */
  uint8_t* packed1;        /* 1-bit (binary) codes, same layouts */
  uint8_t* packed1_t;
  int32_t N, K, tile0, reserved;
} ob_pack_desc;
int ob_weight_quant_pack_multi(const ob_pack_desc* descs_dev, int count, int total_tiles, int alpha_mode, ob_stream_t stream);

/* Dense W_hat = alpha_eff * Q in fp32 over n elements: `quantize_weight` / _QuantizeSTE.forward
 * (quant.py:45-70, 95-96). */
int ob_weight_quant_dense(const float* W, const float* alpha, int alpha_mode, int64_t n,
                          int bitwidth, float* w_hat, ob_stream_t stream);

/* _QuantizeSTE.backward (quant.py:72-92) on a dense upstream gradient g [n]:
 * grad_W = g * 1[|Wa| <= 1]; grad_alpha = sum g * term (times sign(alpha) in OB_ALPHA_RAW mode).
 * ws: at least ob_ste_workspace_bytes(n) bytes. */
size_t ob_ste_workspace_bytes(int64_t n);
int ob_weight_ste_backward(const float* g, const float* W, const float* alpha, int alpha_mode,
                           int64_t n, int bitwidth, float* grad_W, float* grad_alpha, void* ws,
                           ob_stream_t stream);

/* Unpack a packed code matrix [R, C/4] back to int8 codes [R, C] (tests, checkpoint export). */
int ob_unpack_codes(const uint8_t* packed, int R, int C, int order, int8_t* codes,
                    ob_stream_t stream);

/* Per-token absmax int8 activation quantiser (north_star (b); no reference counterpart, SURVEY.md
 * section 0 row D4):  s = (1/max(amax,1e-5))*127,  q = clamp(rint(x*s), -128, 127).
 * x [M, K] of dtype x_dtype, q [M, K] int8, scale [M] fp32.  K % 16 == 0. */
int ob_act_quant_i8(const void* x, int x_dtype, int64_t M, int K, int8_t* q, float* scale,
                    ob_stream_t stream);

/* Forward: y[m,n] = (sum_k q[m,k]*Q[n,k]) * (alpha_eff / scale[m]) + bias[n]   (F.linear, quant.py:126)
 * int8 x ternary on tcgen05 (kind::i8, int32 accumulators in TMEM); packed weights are expanded in
 * shared memory.  bias may be NULL.  y_dtype OB_F32 or OB_BF16. */
int ob_gemm_tern_i8_fwd(const int8_t* q, const float* scale, const uint8_t* packed_i8,
                        const float* alpha, int alpha_mode, const float* bias, int M, int N, int K,
                        void* y, int y_dtype, ob_stream_t stream);

/* Backward, step 1: one pass over dY [M, N] (dy_dtype) and q [M, K]:
 *   dys[m,n]  = bf16(dY[m,n] / scale[m])          (operand of both backward GEMMs)
 *   qb[m,k]   = bf16(q[m,k])                      (exact; operand of grad_W; may be NULL)
 *   colsum    = per-row-block partial column sums of dY (for grad_bias; may be NULL)
 * colsum has ob_bwd_colsum_blocks(M) * N floats. */
int ob_bwd_colsum_blocks(int M);
int ob_bwd_prep(const void* dY, int dy_dtype, const float* scale, const int8_t* q, int M, int N,
                int K, void* dys_bf16, void* qb_bf16, float* colsum, ob_stream_t stream);

/* grad_x[m,k] = alpha_eff * scale[m] * sum_n dys[m,n] * Q[n,k]    (LinearBackward of quant.py:126,
 * identity STE through the activation quantiser); bf16 tcgen05, packed_t expanded in shared memory. */
int ob_bwd_dx(const void* dys_bf16, const float* scale, const uint8_t* packed_t, const float* alpha,
              int alpha_mode, int M, int N, int K, void* dx, int dx_dtype, ob_stream_t stream);

/* grad_W_hat = dys^T @ qb  (bf16 tcgen05, split over tokens, fp32 partials in ws), then fused:
 *   grad_W     = grad_W_hat * 1[|W/alpha_eff| <= 1]                       (quant.py:81-82)
 *   grad_alpha = sum grad_W_hat * term  (* sign(alpha) when OB_ALPHA_RAW)  (quant.py:86-91,124)
 *   grad_bias  = column sums of dY from `colsum` (may be NULL together with grad_bias)
 * ws: at least ob_bwd_dw_workspace_bytes(M,N,K) bytes. */
size_t ob_bwd_dw_workspace_bytes(int M, int N, int K);
int ob_bwd_dw(const void* dys_bf16, const void* qb_bf16, const float* colsum, const float* W,
              const float* alpha, int alpha_mode, int bitwidth, int M, int N, int K, float* grad_W,
              float* grad_alpha, float* grad_bias, void* ws, size_t ws_bytes, ob_stream_t stream);

/* ob_bwd_dw reading the saved int8 activation codes q [M, K] themselves (round 2): CTA pairs on [256 x 256] output tiles
 * convert the codes to bf16 (exact) in shared memory, so ob_bwd_prep / ob_bwd_prep_fused are called with qb_bf16 = NULL and
 * the bf16 copy of q never exists in HBM.  Same results as ob_bwd_dw up to the fp32 summation order of the token splits;
 * same workspace size. */
int ob_bwd_dw_q8(const void* dys_bf16, const int8_t* q, const float* colsum, const float* W,
                 const float* alpha, int alpha_mode, int bitwidth, int M, int N, int K, float* grad_W,
                 float* grad_alpha, float* grad_bias, void* ws, size_t ws_bytes, ob_stream_t stream);
/* The same over a stacked batch whose token rows [0, rows2) went through the 2-bit codes and rows [rows2, M) through the
 * 1-bit codes (train.py:83-103 evaluated side by side): one GEMM launch whose token splits never span the boundary and one
 * finaliser - grad_W is the masked sum over all rows, grad_alpha the sum of the two bitwidths' alpha terms, grad_bias the
 * column sums of all rows (colsum from ONE ob_bwd_prep* call over the M rows). */
int ob_bwd_dw_q8_groups(const void* dys_bf16, const int8_t* q, const float* colsum, const float* W,
                        const float* alpha, int alpha_mode, int rows2, int M, int N, int K, float* grad_W,
                        float* grad_alpha, float* grad_bias, void* ws, size_t ws_bytes, ob_stream_t stream);

/* Fused FFN mid-section (conformer.py:36-39): z = dropout(swish(h)) followed by the
 * activation quantiser of the next layer, in one pass.  h [M, K] fp32, K in {256, 512, 1024, 2048}; inv_keep = 1/(1-p).
 * The dropout mask (nn.Dropout of conformer.py:38) comes from one of two places:
 *   keep != NULL          : [M, K] bytes, 1 = keep (explicit mask, used by the parity tests);
 *   keep == NULL, thr != 0: counter-based Philox4x32-10 keyed by `seed`; a block of four 32-bit words is read as eight
 *                           16-bit lanes (lane k = half k % 2, low first, of word k / 2).  With f = element / 4 the
 *                           128-bit counter is (f & ~32, offset) and the element's lane is 4 * ((f >> 5) & 1) +
 *                           element % 4; it is kept iff lane >= thr, thr = round(p * 2^16) < 2^16, and the caller passes
 *                           the exact inv_keep = 65536 / (65536 - thr).  The backward regenerates the same bits from
 *                           (seed, offset): no mask is stored;
 *   keep == NULL, thr == 0: no dropout (inv_keep ignored).
 * ob_swish_drop_bwd: g_h = g_z * keep * inv_keep * swish'(h), n = M * K elements (n % 256 == 0). */
int ob_swish_drop_quant(const float* h, const uint8_t* keep, float inv_keep, uint64_t seed, uint64_t offset,
                        uint32_t drop_threshold, int64_t M, int K, int8_t* q, float* scale, ob_stream_t stream);
int ob_swish_drop_bwd(const float* gz, const float* h, const uint8_t* keep, float inv_keep, uint64_t seed,
                      uint64_t offset, uint32_t drop_threshold, int64_t n, float* gh, ob_stream_t stream);

/* LayerNorm in front of the routed projections (conformer.py:35, 109): y = (x - mean) * rstd * gamma + beta over the
 * last axis C in {128, 256, 512, 1024}, fp32; mean/rstd [M] are kept for the backward.  ob_layernorm_bwd writes dx and
 * the parameter gradients (deterministic two-stage reduction; ws >= ob_layernorm_bwd_workspace_bytes(C)). */
int ob_layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, int64_t M, int C,
                     float* y, float* mean, float* rstd, ob_stream_t stream);
size_t ob_layernorm_bwd_workspace_bytes(int C);
int ob_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd,
                     const float* gamma, int64_t M, int C, float* dx, float* dgamma, float* dbeta,
                     void* ws, ob_stream_t stream);

/* Fusions around the layer inside the routed modules (conformer.py:35-45, 109-138).
 *
 * ob_layernorm_quant_fwd: LayerNorm and the per-token absmax int8 quantiser of the projection(s) that read it, in one
 *   pass: writes only the int8 codes q [M, C], the scales [M] and mean / rstd [M] (the normalised fp32 tensor never
 *   reaches HBM).  q and scale equal ob_act_quant_i8(ob_layernorm_fwd(x)) bit for bit.
 * ob_layernorm_bwd3: ob_layernorm_bwd with the upstream gradient given as dy0 (+ dy1 (+ dy2)) - the gradients of up to
 *   three projections that consumed the same normalised tensor (q, k, v) are summed on load - and, when `resid` is not
 *   NULL, the gradient that reached the module input along the residual path added on store: dx = LN'(...) + resid.
 * ob_gemm_tern_i8_fwd_tail: ob_gemm_tern_i8_fwd (fp32 output) with the module tail in the epilogue,
 *     out[m,n] = resid[m,n] + (keep(m,n) ? y[m,n] * tail_factor * rowmask[row_base + m] : 0),
 *   i.e. x + scale * dropout(y) * frame_mask with tail_factor = scale / (1 - p); keep = 16-bit Philox lane
 *   ((row_base + m) * N + n) % 8 of block ((row_base + m) * N + n) / 8 >= drop_threshold (0 = no dropout), the stream of
 *   ob_residual_dropout_fwd; rowmask (float per global row) may be NULL.  Same bits as the two separate calls.
 * ob_bwd_prep_fused: ob_bwd_prep whose dY is produced on the fly from the gradient g_next of the op that follows the
 *   layer: OB_PREP_TAIL  dY = g_next * factor * rowmask[row] * keep   (backward of the fused tail above),
 *          OB_PREP_SWISH dY = g_next * factor * keep * swish'(h)     (backward of ob_swish_drop_quant in front of the NEXT
 *                                                                      layer; h = this layer's fp32 output [M, N]). */
#define OB_PREP_TAIL 1
#define OB_PREP_SWISH 2
int ob_layernorm_quant_fwd(const float* x, const float* gamma, const float* beta, float eps, int64_t M, int C,
                           int8_t* q, float* scale, float* mean, float* rstd, ob_stream_t stream);
int ob_layernorm_bwd3(const float* dy0, const float* dy1, const float* dy2, const float* x, const float* mean,
                      const float* rstd, const float* gamma, const float* resid, int64_t M, int C, float* dx,
                      float* dgamma, float* dbeta, void* ws, ob_stream_t stream);
int ob_gemm_tern_i8_fwd_tail(const int8_t* q, const float* scale, const uint8_t* packed_i8, const float* alpha,
                             int alpha_mode, const float* bias, int M, int N, int K, const float* resid,
                             const float* rowmask, float tail_factor, uint64_t seed, uint64_t offset,
                             uint32_t drop_threshold, int64_t row_base, float* out, ob_stream_t stream);
int ob_bwd_prep_fused(const float* g_next, int mode, const float* rowmask, const float* h, float factor,
                      uint64_t seed, uint64_t offset, uint32_t drop_threshold, int64_t row_base,
                      const float* scale, const int8_t* q, int M, int N, int K, void* dys_bf16, void* qb_bf16,
                      float* colsum, ob_stream_t stream);

/* Element-wise chain between the attention matmuls of the reference MHSA (conformer.py:118-128), fused:
 * y = nan_to_num(softmax(masked_fill((ac + rel_shift(bd)) * scale))), attn_d = dropout(y).  ac, bd [B,H,T,T] fp32 (bd
 * BEFORE the relative shift), mask [B,T,T] bytes (0 = masked), T <= 2048.  Dropout as in ob_swish_drop_quant: explicit
 * keep [B,H,T,T] bytes, or keep == NULL with drop_threshold != 0 for the Philox stream (element lane + 32 u of row
 * r = (b*H + h)*T + i is 16-bit lane u % 8 of the block with counter ((r * 32 + lane) * 8 + u / 8, offset)), or neither
 * (then attn_d must be NULL and only y is written).  ac, bd, y, attn_d and the gradients are [B,H,T] rows of T floats with a
 * common row pitch ld >= T (elements), so they can be the padded outputs / inputs of ob_gemm_f32.
 * Backward: gradients w.r.t. ac and the un-shifted bd from the gradient w.r.t. attn_d (or y without dropout). */
int ob_relattn_softmax_fwd(const float* ac, const float* bd, const uint8_t* mask, const uint8_t* keep,
                           float inv_keep, uint64_t seed, uint64_t offset, uint32_t drop_threshold, float scale,
                           int B, int H, int T, int ld, float* y, float* attn_d, ob_stream_t stream);
int ob_relattn_softmax_bwd(const float* gd, const float* y, const uint8_t* keep, float inv_keep, uint64_t seed,
                           uint64_t offset, uint32_t drop_threshold, float scale, int B, int H, int T, int ld,
                           float* d_ac, float* d_bd, ob_stream_t stream);

/* Glue of the attention core (conformer.py:113-117): qu = q + u and qw = q + w in one pass over q [M, W] (u, w: [W], the
 * flattened pos_bias_u / pos_bias_v), and for the backward g = ga + gb with sums[0] = column sums of ga (gradient of u),
 * sums[1] = column sums of gb (gradient of w); fixed-order reduction, ws >= ob_add_colsum2_workspace_bytes(M, W).
 * W a multiple of 4 (ob_add_bias2) / of 64 (ob_add_colsum2). */
int ob_add_bias2(const float* q, const float* u, const float* w, int64_t M, int W, float* qu, float* qw,
                 ob_stream_t stream);
size_t ob_add_colsum2_workspace_bytes(int64_t M, int W);
int ob_add_colsum2(const float* ga, const float* gb, int64_t M, int W, float* g, float* sums, void* ws,
                   ob_stream_t stream);

/* Tail of every encoder module (conformer.py:45, 133-138, 163-167): out = x + scale * dropout(y) * rowmask in one pass
 * ([M, C] fp32, rowmask [M] of 0/1 or NULL, C a multiple of 8).  Dropout as in ob_swish_drop_quant (drop_threshold = 0: none);
 * element e is 16-bit lane e % 8 of the Philox block with counter (e / 8, offset).  Backward: g_y = scale * dropmask *
 * rowmask * g (the gradient w.r.t. x is g itself). */
int ob_residual_dropout_fwd(const float* x, const float* y, const float* rowmask, float scale, float inv_keep,
                            uint64_t seed, uint64_t offset, uint32_t drop_threshold, int64_t M, int C, float* out,
                            ob_stream_t stream);
int ob_residual_dropout_bwd(const float* g, const float* rowmask, float scale, float inv_keep, uint64_t seed,
                            uint64_t offset, uint32_t drop_threshold, int64_t M, int C, float* gy, ob_stream_t stream);

/* First layer of the subsampling front-end (conformer.py:177-181): y = relu(conv2d(x, w, bias, stride 2)) with one input
 * channel, 3 x 3 taps and C = 256 output channels.  x [B, T, F] fp32, w [C, 9] (the [C,1,3,3] weight), bias [C] or NULL;
 * y [B, T1, F1, C] channels-last, T1 = (T-3)/2+1, F1 = (F-3)/2+1.  Backward: g [B, T1, F1, C] -> gw [C, 9], gb [C] (the
 * ReLU mask is recomputed from x; the input needs no gradient); ws >= ob_conv1_relu_workspace_bytes(), fixed-order sums. */
size_t ob_conv1_relu_workspace_bytes(void);
int ob_conv1_relu_fwd(const float* x, const float* w, const float* bias, int B, int T, int F, int C, float* y,
                      ob_stream_t stream);
int ob_conv1_relu_bwd(const float* g, const float* x, const float* w, const float* bias, int B, int T, int F, int C,
                      float* gw, float* gb, void* ws, ob_stream_t stream);

/* Greedy CTC decoding (onebit_asr/metrics.py:51-60): per-frame argmax over V (first maximal index), blanks dropped,
 * repeats collapsed.  logits [B, T, V] (dtype tag), lens [B] valid frames; out_tokens [B, T] int32 (compacted, padded
 * with -1), out_lens [B].  ws: at least ob_ctc_decode_workspace_bytes(B, T) bytes. */
size_t ob_ctc_decode_workspace_bytes(int B, int T);
int ob_ctc_greedy_decode(const void* logits, int dtype, int B, int T, int V, const int32_t* lens,
                         int blank_id, int32_t* out_tokens, int32_t* out_lens, void* ws,
                         ob_stream_t stream);

/* CTC loss of the training step from the logits (onebit_asr/losses.py:41-47: log_softmax, transpose, nn.CTCLoss(blank,
 * zero_infinity=True), reduction 'mean'): loss[0] = mean_b(nll_b / max(L_b, 1)), an infinite nll_b counts as 0 and gets a
 * zero gradient.  logits [B, T, V] fp32 with row pitch ld (elements); in_lens [B], tgt_lens [B] int64 on the device;
 * targets [B, Lmax] int64, row pitch tgt_ld, Lmax <= 511.  The caller owns lse [B*T], alpha and beta [B, T, Sp] with
 * Sp = ob_ctc_state_pitch(Lmax), nll [B] and loss [1]; the backward reads them again together with grad_out [1] (the
 * upstream gradient of the loss, on the device) and writes grad_logits [B, T, V] (row pitch ldg): g_b (softmax - state
 * occupancy per label), zero beyond in_lens[b].  Fixed-order reductions (deterministic). */
int ob_ctc_state_pitch(int Lmax);
int ob_ctc_loss_fwd(const float* logits, int64_t ld, const int64_t* in_lens, const int64_t* targets, int64_t tgt_ld,
                    const int64_t* tgt_lens, int B, int T, int V, int Lmax, int blank, float* lse, float* alpha,
                    float* beta, float* nll, float* loss, ob_stream_t stream);
int ob_ctc_loss_bwd(const float* logits, int64_t ld, const int64_t* in_lens, const int64_t* targets, int64_t tgt_ld,
                    const int64_t* tgt_lens, int B, int T, int V, int Lmax, int blank, const float* lse,
                    const float* alpha, const float* beta, const float* nll, const float* grad_out, float* grad_logits,
                    int64_t ldg, ob_stream_t stream);

/* Batched fp32 GEMM on the tensor cores for the non-routed matmuls of the model (attention products conformer.py:113-129,
 * vocabulary projections, 1x1 convolutions - fp32 torch.matmul / nn.Linear / nn.Conv1d in the reference):
 *   D[b0,b1](m, n) (+)= scale * sum_k A[b0,b1](m, k) * B[b0,b1](n, k) + bias[n]
 * passes = 3: every operand is split into tf32 hi + lo parts in shared memory and three products are accumulated in
 * fp32 (error ~2^-22 relative per term: fp32-level); passes = 1: plain tf32.  Operand element (m, k) of A lives at
 * A + b0*a_bs0 + b1*a_bs1 + m*lda + k (K-major) or ... + k*lda + m (a_mn_major = 1); B likewise with (n, k) (K-major
 * B is an [N, K] weight; MN-major B is a [K, N] matrix); D is row-major [M, N] with pitch ldd.  A batch stride of 0
 * broadcasts an input over that batch axis.  All pointers 16-byte aligned, leading dimensions and batch strides
 * multiples of 4 elements; M, N, K arbitrary (> 0).  accumulate = 1 adds into D (one atomic vector add per element, so
 * the result is deterministic unless a D batch stride of 0 makes several batch items share an output).  Unbatched
 * contractions of K >= 2048 with few output tiles are split over K in chunks of 1024 whose products go to the workspace
 * and are summed in fixed order: this fills the SMs and keeps the truncating in-tensor-core accumulation chains short.
 * ws: at least ob_gemm_f32_workspace_bytes(M, N, K, nb0, nb1) bytes (0 for most shapes; ws may then be NULL).
 * Measured error vs fp64: <= 1e-5 of max|D| (passes = 3). */
size_t ob_gemm_f32_workspace_bytes(int M, int N, int K, int nb0, int nb1);
int ob_gemm_f32(const float* A, int a_mn_major, int64_t lda, int64_t a_bs0, int64_t a_bs1, const float* B,
                int b_mn_major, int64_t ldb, int64_t b_bs0, int64_t b_bs1, float* D, int64_t ldd, int64_t d_bs0,
                int64_t d_bs1, const float* bias, float scale, int accumulate, int M, int N, int K, int nb0, int nb1,
                int passes, void* ws, size_t ws_bytes, ob_stream_t stream);

/* Middle of the convolution module (conformer.py:141-167) in the [B, T, C] layout: GLU -> depthwise conv1d (ks odd <= 31
 * taps, zero padding ks/2, per utterance) -> BatchNorm with batch statistics over all B*T frames (biased variance, as
 * nn.BatchNorm1d(track_running_stats=False)) -> swish.  The 1x1 convolutions on either side are ob_gemm_f32 products.
 *   ob_glu_dwconv_bn_fwd: a [B,T,2C] (output of the first 1x1 conv) -> d [B,T,C] = dwconv(a[..,:C] * sigmoid(a[..,C:])) + bias,
 *                         mean[C], rstd[C] of d (Chan-combined tile partials, fp64);  w [C, ks], bias [C] or NULL
 *   ob_bn_swish_fwd:      s = swish(gamma * (d - mean) * rstd + beta)
 *   ob_bn_swish_bwd:      g_s -> g_d; g_gamma_beta [2][C] = (g_beta, g_gamma)
 *   ob_glu_dwconv_bwd:    g_d -> g_a [B,T,2C], g_w [C, ks], g_bias [C] (or NULL)
 * C a multiple of 64.  ws: at least ob_convmod_workspace_bytes(B, T, C) bytes; all reductions have a fixed order.
 * groups > 1: the batch is a stack of `groups` independent passes (B / groups utterances each, contiguous) whose
 * BatchNorm statistics are separate: mean, rstd are [groups][C] and g_gamma_beta is [groups][2][C] (sum over the groups
 * = the parameter gradients). */
size_t ob_convmod_workspace_bytes(int B, int T, int C);
int ob_glu_dwconv_bn_fwd(const float* a, const float* w, const float* bias, int B, int T, int C, int ks, float eps,
                         int groups, float* d, float* mean, float* rstd, void* ws, ob_stream_t stream);
int ob_bn_swish_fwd(const float* d, const float* mean, const float* rstd, const float* gamma, const float* beta,
                    int64_t M, int C, int groups, float* s, ob_stream_t stream);
int ob_bn_swish_bwd(const float* gs, const float* d, const float* mean, const float* rstd, const float* gamma,
                    const float* beta, int64_t M, int C, int groups, float* gd, float* g_gamma_beta, void* ws,
                    ob_stream_t stream);
int ob_glu_dwconv_bwd(const float* gd, const float* a, const float* w, int B, int T, int C, int ks, float* ga,
                      float* gw, float* gbias, void* ws, ob_stream_t stream);

/* Column sums of a row-major fp32 [M, N] matrix (N % 4 == 0): out[n] = sum_m x[m, n], fixed summation order.  Used for the
 * bias gradients of the non-routed linears around the layer.  ws >= ob_colsum_workspace_bytes(M, N). */
size_t ob_colsum_workspace_bytes(int64_t M, int N);
int ob_colsum(const float* x, int64_t M, int N, float* out, void* ws, ob_stream_t stream);

/* Debug/tuning knob (tests and profiling only): key/value pairs, see csrc/ob_gemm.cu. */
int ob_debug_set(int key, int value);

#ifdef __cplusplus
}
#endif
#endif /* ONEBIT_H_ */
