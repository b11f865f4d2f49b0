/*
 * s2svc_b200.h -- C ABI of libs2svc_b200.so: the B200 (sm_100a) kernels behind the seq2seq-vc
 * training hot path.
 *
 * The reference (unilight/seq2seq-vc) is pure Python; it has no FFI of its own.  Each entry point
 * below therefore names the reference *function* whose arithmetic it replaces (file:line relative
 * to the reference root).  INTEGRATION.md shows the ctypes binding a reference maintainer adds.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless its name ends in _host;
 *  - no entry point allocates, synchronises the device, or keeps global mutable state other than a
 *    once-initialised driver-entry-point / function-attribute cache; all are re-entrant across
 *    streams; workspaces are passed by the caller;
 *  - return value: S2S_OK (0) or a negative S2S_ERR_* code; s2s_last_error() returns the message
 *    of the last failure on the calling thread;
 *  - `stream` is a cudaStream_t passed as void* so that the header needs no CUDA include;
 *  - `dtype` arguments are S2S_F32 / S2S_BF16 and describe activations in HBM; statistics,
 *    parameters (gamma/beta/bias), parameter gradients and losses are always float32;
 *  - per-utterance lengths are int32 device arrays.
 */
#ifndef S2SVC_B200_H_
#define S2SVC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S2S_OK 0
#define S2S_ERR_INVALID (-1)
#define S2S_ERR_UNSUPPORTED (-2)
#define S2S_ERR_CUDA (-3)

#define S2S_F32 0
#define S2S_BF16 1

#define S2S_ABI_VERSION 22

const char* s2s_last_error(void);
int s2s_abi_version(void);
/* 0 when the current device is compute capability 10.x (B200), else S2S_ERR_UNSUPPORTED */
int s2s_device_check(void);
/* number of kernel launches issued through this library by the calling process (for bench.py) */
int64_t s2s_launch_count(void);
/* number of s2s_gemm(mode = 1) calls that could not be described to TMA (unaligned strides, N < 8)
 * and were served by the CUDA-core kernel instead */
int64_t s2s_tc_fallback_count(void);
/* test hook: force the tile shape of s2s_gemm(mode = 1): 1 = one CTA per 128-row tile, 2 = CTA pairs on 256-row tiles
 * (cta_group::2), 0 = cost model (also clears a pinned N tile); a value >= 16 pins the N tile to that width instead */
void s2s_debug_gemm_tile(int cg);

/* Counter-based dropout: element idx is dropped iff hash(seed', stream, idx) < p * 2^32 where
 * seed' = seed + (seed_dev ? *seed_dev : 0); kept values are scaled by 1/(1-p).  Backward kernels
 * regenerate the mask from the same triple.  seed_dev lets a CUDA graph replay see a new seed. */
typedef struct {
    float p;
    uint64_t seed;
    uint64_t stream;
    const uint64_t* seed_dev;
} s2s_dropout_t;

/* -------------------------------------------------------------------------------------------
 * GEMM  (replaces every torch.nn.Linear / torch.matmul / Conv1d / Conv2d-as-im2col on the path:
 *        modules/transformer/attention.py:40-111, positionwise_feed_forward.py:30-32,
 *        subsampling.py:58-94, pre_postnets.py:60-66,105-185, models/vtn.py:181-182,249-251)
 *
 *   C[b1,b2][m,n] = epilogue( alpha * sum_{t<taps} sum_{k<K} A[b1,b2][m + t, k] * B[b1,b2][n, t, k] )
 *   A(m,k) = A + b1*a_bs1 + b2*a_bs2 + m*a_rs + k*a_cs      (element strides; a_rs==1 or a_cs==1)
 *   B(n,t,k) = B + b1*b_bs1 + b2*b_bs2 + n*b_rs + t*b_ts + k*b_cs
 *   C(m,n) = C + b1*c_bs1 + b2*c_bs2 + m*c_rs + n           (residual R uses the same strides)
 *   epilogue: (+ bias[n]) -> relu? -> dropout? -> (+ R[m,n]) -> (+ C[m,n] if accumulate)
 *             -> row mask -> store
 *   row mask (mask_period > 0): rows with ((m + mask_offset) % mask_period) outside
 *   [mask_lo, mask_hi) are stored as 0 (halo rows of the zero-padded conv1d layout).
 *   (R: see r_mode below; with r_mode = 1 it gates instead of adding.)
 * mode 0: fp32 CUDA-core path (operands f32 or bf16, fp32 FMA accumulate) -- the numerical yard-stick.
 * mode 1: bf16 tcgen05 tensor-core path (operands must be bf16, TMEM fp32 accumulate).
 * mode 2: fp32-accurate tensor-core path: float32 operands are split into 2 or 3 bf16 pieces each (split_terms = 3 / 6
 *         partial products, concatenated along K in a workspace) and multiplied by the same tcgen05 kernel with fp32
 *         TMEM accumulation -- the parity mode (mel L1 <= 1e-4 against the reference) on the tensor cores.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int M, N, K, taps;
    const void* A; int a_dtype; int64_t a_rs, a_cs, a_bs1, a_bs2;
    const void* B; int b_dtype; int64_t b_rs, b_cs, b_ts, b_bs1, b_bs2;
    void* C; int c_dtype; int64_t c_rs, c_bs1, c_bs2;
    int batch1, batch2;
    const float* bias;
    const void* R;          /* optional residual, dtype c_dtype, strides of C */
    float alpha;
    int relu;
    int accumulate;
    s2s_dropout_t drop;
    int mask_period, mask_offset, mask_lo, mask_hi;
    int r_mode;             /* how R enters: 0 = residual (x + R), 1 = ReLU gate (R > 0 ? x * r_scale : 0), the backward of
                               relu (+ dropout scale) folded into the dX GEMM of the layer above (positionwise_feed_forward.py:30-32) */
    float r_scale;
    int split_terms;        /* mode 2 only: 3 = two-way bf16 split (hi*hi + lo*hi + hi*lo), 6 = three-way split (fp32-accurate) */
    void* ws;               /* mode 2 only: workspace of at least s2s_gemm_workspace_bytes(g) bytes */
    size_t ws_bytes;
} s2s_gemm_t;

int s2s_gemm(const s2s_gemm_t* g, int mode, void* stream);
/* n independent GEMMs issued together.  In mode 1, sets of up to 8 weight-gradient products (bf16 operands contiguous along
 * M / N, float32 C accumulated in place, no epilogue extras, no batch / taps) share ONE persistent tcgen05 launch -- the
 * dW = dy^T x products of a Transformer layer are too small to fill the machine one at a time; anything else is launched
 * one by one, so the call is always equivalent to n s2s_gemm calls. */
int s2s_gemm_grouped(const s2s_gemm_t* gs, int n, int mode, void* stream);
/* bytes of workspace s2s_gemm(g, mode = 2) needs for this problem */
size_t s2s_gemm_workspace_bytes(const s2s_gemm_t* g);

/* -------------------------------------------------------------------------------------------
 * LayerNorm over the last dim, eps = 1e-12 in the reference (modules/transformer/layer_norm.py:12-42)
 * fwd: y = (x - mean) * rstd * gamma + beta ; saves mean / rstd (float32, one per row)
 * bwd: dx (+ dres when given: the gradient arriving over the residual branch) ; dgamma += ,
 *      dbeta += (float32 accumulate into the caller's gradient buffers)
 * ------------------------------------------------------------------------------------------- */
int s2s_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean,
                      float* rstd, int64_t rows, int d, float eps, int dtype, void* stream);
int s2s_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean,
                      const float* rstd, const void* dres, void* dx, float* dgamma, float* dbeta,
                      int64_t rows, int d, int dtype, void* stream);
/* same, with a second output dx_drop = dropout'(dx) (mask regenerated from `drop`): the gradient entering the
 * residual branch `x = LN(drop(f(..)) + res)` of a post-LN block (decoder_layer.py:63-134) without a separate pass */
int s2s_layernorm_bwd_drop(const void* dy, const void* x, const float* gamma, const float* mean,
                           const float* rstd, const void* dres, void* dx, void* dx_drop, const s2s_dropout_t* drop,
                           float* dgamma, float* dbeta, int64_t rows, int d, int dtype, void* stream);

/* Skinny linear layer with 1 <= N <= 4 output features (the stop-token head prob_out,
 * models/vtn.py:182,251): y[r, j] = sum_k x[r, k] w[j, k] + bias[j]; x (rows, K) and w (N, K) in
 * `dtype`.  bwd: dw[j, k] += sum_r dy[r, j] x[r, k] ; dbias[j] += sum_r dy[r, j] ;
 * dx[r, k] (+)= sum_j dy[r, j] w[j, k].  dw / dbias / dx may each be NULL. */
int s2s_skinny_linear_fwd(const void* x, const void* w, const float* bias, void* y, int64_t rows, int K, int N,
                          int dtype, void* stream);
int s2s_skinny_linear_bwd(const void* dy, const void* x, const void* w, float* dw, float* dbias, void* dx,
                          int dx_accumulate, int64_t rows, int K, int N, int dtype, void* stream);

/* out[c] += sum_r x[r, c]  (bias gradients); x row stride ld */
int s2s_colsum(const void* x, int64_t rows, int cols, int64_t ld, float* out, int dtype, void* stream);
/* n column sums in one launch (the bias gradients of one layer's backward, whose gradient tensors are all alive at its end) */
typedef struct {
    const void* x;
    int64_t rows;
    int cols;
    int64_t ld;
    float* out;             /* out[c] += sum_r x[r, c] */
} s2s_colsum_t;
int s2s_colsum_multi(const s2s_colsum_t* items, int n, int dtype, void* stream);
/* dx = dy * (y > 0 ? scale : 0): backward of relu followed by dropout (scale = 1/(1-p)) */
int s2s_relu_bwd(const void* dy, const void* y, void* dx, int64_t n, float scale, int dtype, void* stream);
/* out = a + b (elementwise; out may alias a or b): residual joins that no GEMM epilogue absorbs
 * (models/vtn.py:257-259 postnet residual, gradient joins of the residual branches) */
int s2s_add(const void* a, const void* b, void* out, int64_t n, int dtype, void* stream);
/* dx[r,c] = dy[r,c] * dropout_factor(idx = r*cols + c): backward of an epilogue dropout */
int s2s_dropout_bwd(const void* dy, void* dx, int64_t rows, int cols, const s2s_dropout_t* drop, int dtype,
                    void* stream);

/* -------------------------------------------------------------------------------------------
 * Masked softmax over the key axis (modules/transformer/attention.py:76-85):
 * S is (B, H, T1, ld) with T2 <= ld valid columns, already scaled by 1/sqrt(d_k).  Key j of batch b
 * is visible to query i iff j < klens[b] and (!causal or j <= i).  P = softmax over visible keys,
 * exact zeros elsewhere (also for rows with no visible key).  In place (P may equal S).
 * If Pd != NULL it also receives dropout(P) (attention dropout, attention.py:85).
 * bwd: dS = scale * P * (dP - sum_j P_j dP_j), in place over dP; when drop->p > 0, dP is first
 * multiplied by the regenerated dropout factor.
 * ------------------------------------------------------------------------------------------- */
int s2s_softmax_fwd(const void* S, void* P, void* Pd, const int32_t* klens, int B, int H, int T1, int T2,
                    int64_t ld, int causal, const s2s_dropout_t* drop, int dtype, void* stream);
int s2s_softmax_bwd(const void* P, void* dP, int B, int H, int T1, int T2, int64_t ld, float scale,
                    const s2s_dropout_t* drop, int dtype, void* stream);

/* Fused attention probabilities for small head dimensions (bf16 only; d_k in {16,32,48,64,96,128}):
 *   fwd: P = softmax_s(scale * q k^T) with the same masking contract as s2s_softmax_fwd (attention.py:95-104,76-85)
 *   bwd: dS = scale * P * (dP - sum_s P dP), dP = dctx v^T (+ dAtt when non-NULL), i.e. s2s_gemm + s2s_softmax_bwd
 * q / k / v / dctx are (B, T, H, d_k) views given by element strides (batch, time, head; d_k contiguous); P, dAtt, dS are
 * (B, H, T1, ld) with ld % 8 == 0.  The product is recomputed with warp-level mma.sync in two passes so the (B,H,T1,T2)
 * matrix crosses HBM once per direction: at d_k = 48 the QK^T GEMM is pure epilogue for the tcgen05 tile. */
int s2s_attn_probs_fwd(const void* q, int64_t q_bs, int64_t q_ts, int64_t q_hs, const void* k, int64_t k_bs, int64_t k_ts,
                       int64_t k_hs, void* P, const int32_t* klens, int B, int H, int T1, int T2, int dk, int64_t ld, float scale,
                       int causal, void* stream);
int s2s_attn_probs_bwd(const void* dctx, int64_t d_bs, int64_t d_ts, int64_t d_hs, const void* v, int64_t v_bs, int64_t v_ts,
                       int64_t v_hs, const void* P, const void* dAtt, void* dS, int B, int H, int T1, int T2, int dk, int64_t ld,
                       float scale, void* stream);

/* Flash-style fused multi-head attention on tcgen05 / TMEM (bf16; d_k a multiple of 16 in [16, 128]) -- replaces
 * MultiHeadedAttention.forward_attention (modules/transformer/attention.py:76-111) and its autograd backward:
 *   fwd: ctx[b,t,h,:] = sum_s P[b,h,t,s] v[b,s,h,:],  P = softmax_s(scale * q k^T) under the key-length / causal mask
 *        (masked columns exactly 0; a row without a visible key gives P = 0, ctx = 0), S and P stay in tensor / shared
 *        memory; lse (B, H, T1p) float32, T1p = T1 rounded up to 64, receives the log2-domain row statistic
 *        max + log2(sum) (+inf for an empty row); when P != NULL the normalised probabilities are ALSO written as
 *        (B, H, T1, ld) bf16 with columns >= the visible keys zeroed (source-attention maps are an output of
 *        VTN.forward, models/vtn.py:280-287).
 *   bwd: recomputes P from q, k and lse; dq = dS k, dk = dS^T q, dv = P^T dctx with dS = scale * P * (dctx v^T - D),
 *        D = rowsum(dctx o ctx) (written to dvec (B, H, T1p) as a workspace).  Two launches (dQ; dK + dV), no atomics.
 * q / ctx / dctx / dq are (B, T1, H, d_k) views, k / v / dk / dv (B, T2, H, d_k) views, each given by element strides
 * (batch, time, head; d_k contiguous), 16-byte aligned with strides that are multiples of 8 elements. */
int s2s_attn_fwd_tc(const void* q, int64_t q_bs, int64_t q_ts, int64_t q_hs, const void* k, const void* v, int64_t kv_bs,
                    int64_t kv_ts, int64_t kv_hs, void* ctx, int64_t c_bs, int64_t c_ts, int64_t c_hs, float* lse, void* P,
                    int64_t ld, const int32_t* klens, int B, int H, int T1, int T2, int dk, float scale, int causal, void* stream);
int s2s_attn_bwd_tc(const void* q, int64_t q_bs, int64_t q_ts, int64_t q_hs, const void* k, const void* v, int64_t kv_bs,
                    int64_t kv_ts, int64_t kv_hs, const void* ctx, const void* dctx, int64_t c_bs, int64_t c_ts, int64_t c_hs,
                    const float* lse, float* dvec, void* dq, int64_t dq_bs, int64_t dq_ts, int64_t dq_hs, void* dk_out, void* dv_out,
                    int64_t dkv_bs, int64_t dkv_ts, int64_t dkv_hs, const int32_t* klens, int B, int H, int T1, int T2, int dk,
                    float scale, int causal, void* stream);

/* -------------------------------------------------------------------------------------------
 * ScaledPositionalEncoding (layers/positional_encoding.py:73-106): y = dropout(x + alpha * pe[t])
 * pe is the float32 sinusoid table (>= T rows of d); alpha is a device scalar.
 * bwd: dx = dy * mask ; dalpha += sum(dy * mask * pe)
 * ------------------------------------------------------------------------------------------- */
int s2s_scaled_pe_fwd(const void* x, const float* pe, const float* alpha, void* y, int B, int T, int d,
                      const s2s_dropout_t* drop, int dtype, void* stream);
int s2s_scaled_pe_bwd(const void* dy, const float* pe, void* dx, float* dalpha, int B, int T, int d,
                      const s2s_dropout_t* drop, int dtype, void* stream);

/* -------------------------------------------------------------------------------------------
 * TransformerTTS encoder input layer (models/transformer_tts.py:63-77,139-142): token embedding with
 * the <eos> append and ScaledPositionalEncoding fused:
 *   tok(b, t) = tokens[b, t] for t < ilens[b] ; eos for t == ilens[b] ; padding_idx beyond
 *   y[b, t, :] = dropout(weight[tok(b, t)] + alpha * pe[t]),  t < T_out (= T_in + 1)
 * bwd: dweight[tok] += dy * mask (never for padding_idx) ; dalpha += sum(dy * mask * pe).
 * tokens (B, T_in) int64, weight (V, d) float32.
 * ------------------------------------------------------------------------------------------- */
int s2s_embed_pe_fwd(const int64_t* tokens, const int32_t* ilens, const float* weight, const float* pe,
                     const float* alpha, void* y, int B, int T_in, int T_out, int d, int eos, int padding_idx,
                     const s2s_dropout_t* drop, int dtype, void* stream);
int s2s_embed_pe_bwd(const void* dy, const int64_t* tokens, const int32_t* ilens, const float* pe, float* dweight,
                     float* dalpha, int B, int T_in, int T_out, int d, int eos, int padding_idx,
                     const s2s_dropout_t* drop, int dtype, void* stream);

/* -------------------------------------------------------------------------------------------
 * Conv2dSubsampling front end (modules/transformer/subsampling.py:58-94).
 * conv1: x (B, T, F) float32 -> relu(conv2d(1->C, 3x3, stride 2)) stored channels-last
 *        y1 (B, T1, F1, C), T1 = (T-1)/2, F1 = (F-1)/2;  w (C, 1, 3, 3) float32, bias (C).
 * conv1_bwd: dw += , dbias += from dy1 (already relu-masked by the caller via s2s_relu_bwd).
 * im2col_s2: y1 (B, T1, F1, C) -> col (B*T2*F2, 9*C), column order (kt, kf, c), stride 2.
 * col2im_s2: the adjoint scatter (gather form), dcol -> dy1.
 * ------------------------------------------------------------------------------------------- */
int s2s_conv1_fwd(const float* x, const float* w, const float* bias, void* y1, int B, int T, int F, int C,
                  int dtype, void* stream);
int s2s_conv1_bwd(const float* x, const void* dy1, float* dw, float* dbias, int B, int T, int F, int C,
                  int dtype, void* stream);
int s2s_im2col_s2(const void* y1, void* col, int B, int T1, int F1, int C, int dtype, void* stream);
int s2s_col2im_s2(const void* dcol, void* dy1, int B, int T1, int F1, int C, int dtype, void* stream);
/* the same scatter-add with the ReLU' of conv.0's output y1 (subsampling.py:60-61) applied at the store: dy1 = col2im(dcol) * (y1 > 0) */
int s2s_col2im_s2_relu(const void* dcol, const void* y1, void* dy1, int B, int T1, int F1, int C, int dtype, void* stream);

/* Decoder input glue (models/vtn.py:227-243,523-527): out[b, 0] = 0, out[b, l] = ys[b, l*r - 1]
 * for l >= 1;  ys (B, L, odim) float32 -> out (B, Lr, odim) dtype. */
int s2s_shift_thin(const float* ys, void* out, int B, int L, int Lr, int odim, int r, int dtype, void* stream);
/* Target glue (models/vtn.py:262-274): olens_out = olens - olens % r ; labels_out = labels with
 * labels_out[b, olens_out[b]-1] = 1 ; (B, Lin) -> (B, Lout), Lout <= Lin. */
int s2s_fix_targets(const float* labels, const int32_t* olens, float* labels_out, int32_t* olens_out, int B,
                    int Lin, int Lout, int r, void* stream);

/* -------------------------------------------------------------------------------------------
 * Postnet BatchNorm1d(+tanh) in training/eval mode over a zero-haloed channels-last buffer
 * (modules/pre_postnets.py:105-185).  x is (B, Lp, C) with frames at rows [halo, halo+L) of each
 * utterance; statistics run over all B*L frames INCLUDING padded frames (the reference applies no mask).
 * bn_stats: sums[0:C] += sum x, sums[C:2C] += sum x^2 (float32; caller zeroes)
 * bn_finalize: mean/invstd from sums (biased var, eps) and running-stat update
 *              (momentum, unbiased var; modules see torch.nn.BatchNorm1d); running_* may be NULL.
 * bn_apply: y = act((x - mean) * invstd * gamma + beta), then dropout; `use_tanh` selects act: 0 identity,
 *           1 tanh (postnet), 2 Swish (Conformer ConvolutionModule, modules/conformer/convolution.py:75, halo = 0);
 *           halo rows of y are written as zeros.
 * bn_bwd_reduce: sums[0:C] += sum dz, sums[C:2C] += sum dz * xhat with dz = dy * mask * act'(y)
 * bn_bwd_apply: dx = gamma * invstd * (dz - sums0/n - xhat * sums1/n); dgamma += sums1, dbeta += sums0;
 *               halo rows of dx are zeros.  (eval mode: pass sums == NULL -> dx = gamma*invstd*dz)
 * ------------------------------------------------------------------------------------------- */
int s2s_bn_stats(const void* x, float* sums, int B, int L, int halo, int C, int dtype, void* stream);
int s2s_bn_finalize(const float* sums, float* mean, float* invstd, float* running_mean, float* running_var,
                    int64_t count, int C, float eps, float momentum, void* stream);
int s2s_bn_apply(const void* x, const float* mean, const float* invstd, const float* gamma, const float* beta,
                 void* y, int B, int L, int halo, int C, int use_tanh, const s2s_dropout_t* drop, int dtype,
                 void* stream);
int s2s_bn_bwd_reduce(const void* dy, const void* y, const void* x, const float* mean, const float* invstd,
                      const float* gamma, const float* beta, float* sums, int B, int L, int halo, int C,
                      int use_tanh, const s2s_dropout_t* drop, int dtype, void* stream);
int s2s_bn_bwd_apply(const void* dy, const void* y, const void* x, const float* mean, const float* invstd,
                     const float* gamma, const float* beta, const float* sums, void* dx, float* dgamma,
                     float* dbeta, int B, int L, int halo, int C, int use_tanh, const s2s_dropout_t* drop,
                     int dtype, void* stream);
/* eval-mode statistics: mean = running_mean, invstd = rsqrt(running_var + eps) */
int s2s_bn_eval_stats(const float* running_mean, const float* running_var, float* mean, float* invstd, int C,
                      float eps, void* stream);
/* copy (B, L, C) <-> haloed (B, L + 2*halo, C) ; halo rows zeroed on pad */
int s2s_pad_rows(const void* x, void* y, int B, int L, int halo, int C, int dtype, void* stream);
int s2s_unpad_rows(const void* x, void* y, int B, int L, int halo, int C, int dtype, void* stream);

/* -------------------------------------------------------------------------------------------
 * Seq2SeqLoss (losses/seq2seq_loss.py:13-59): masked mean L1(after) + L1(before) and
 * BCE-with-logits(pos_weight) over frames l < olens[b].  losses[0] = l1, losses[1] = bce.
 * Also writes the gradients of (l1 + bce) w.r.t. after / before / logits (zero on padded frames).
 * after/before/logits are `dtype` with L frames per utterance; ys (B, L_ys, odim) and labels
 * (B, L_labels) are float32 and may be longer than L (only the first L frames are read).
 * workspace: >= 4 floats, zeroed by the call.
 * ------------------------------------------------------------------------------------------- */
int s2s_seq2seq_loss(const void* after, const void* before, const void* logits, const float* ys,
                     const float* labels, const int32_t* olens, int B, int L, int L_ys, int L_labels, int odim,
                     float pos_weight, float* losses, void* d_after, void* d_before, void* d_logits,
                     float* workspace, int dtype, void* stream);

/* -------------------------------------------------------------------------------------------
 * GuidedMultiHeadAttentionLoss (losses/guided_attention_loss.py:6-165):
 * loss = alpha * mean over (b, h, t < olens[b], s < ilens[b]) of att * (1 - exp(-(s/ilen - t/olen)^2 / (2 sigma^2)))
 * att (B, H, T_out, ld) with T_in valid columns.  d_att (optional) receives d loss / d att.
 * workspace: >= 2 floats.
 * ------------------------------------------------------------------------------------------- */
int s2s_guided_attn_loss(const void* att, const int32_t* ilens, const int32_t* olens, int B, int H, int T_out,
                         int T_in, int64_t ld, float sigma, float alpha, float* loss, void* d_att,
                         float* workspace, int dtype, void* stream);

/* -------------------------------------------------------------------------------------------
 * Optimizer tail (trainers/ar_vc.py:99-107): clip_grad_norm_(max_norm) + Adam on flat float32
 * buffers.  sqnorm: out[0] += sum g^2 (caller zeroes).  adam_step: g' = g * grad_scale *
 * min(1, max_norm / (sqrt(sqnorm * grad_scale^2) + 1e-6)); Adam(lr, beta1, beta2, eps) with
 * bias correction for step *step_dev (float32 device scalar, already incremented);
 * lr is read from *lr_dev.  If p_bf16 != NULL the updated parameter is also written as bf16.
 * ------------------------------------------------------------------------------------------- */
int s2s_sqnorm(const float* g, int64_t n, float* out, void* stream);
int s2s_adam_step(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, const float* lr_dev,
                  float beta1, float beta2, float eps, float weight_decay, const float* step_dev,
                  const float* sqnorm, float max_norm, float grad_scale, void* stream);
/* step bookkeeping on device: step[0] += 1 ; seed[0] += 1 */
int s2s_step_advance(float* step, uint64_t* seed, void* stream);

/* dtype conversion / weight packing */
int s2s_cast(const void* in, void* out, int64_t n, int in_dtype, int out_dtype, void* stream);
/* out[n][b][a] (+)= in[n][a][b] : transpose of the two trailing dims (conv weight packing and the
 * adjoint un-packing of the packed gradient) */
int s2s_transpose_last2(const void* in, void* out, int N, int A, int Bd, int in_dtype, int out_dtype,
                        int accumulate, void* stream);

/* Conv1d weight packing (modules/pre_postnets.py:105-185): w (OC, IC, K) float32 ->
 * wp (OC, K, IC) for the forward taps-GEMM and wpt (IC, K, OC) with wpt[ic][u][oc] = w[oc][ic][K-1-u]
 * for the input-gradient taps-GEMM; out_dtype applies to both (either may be NULL). */
int s2s_pack_conv1d_w(const float* w, void* wp, void* wpt, int OC, int IC, int K, int out_dtype, void* stream);

/* -------------------------------------------------------------------------------------------
 * Monotonic alignment search + duration count (modules/alignments.py:63-93, 281-310).
 * log_p (B, T_feats, T_text) float32; per utterance the slice [:feats_lens[b], :text_lens[b]] is
 * searched with the reference's exact arithmetic (float32 sequential prefix on row 0, float64 DP,
 * ">=" tie rule).  paths (B, T_feats) int32 (-1 beyond feats_len), ds (B, T_text) float32 counts.
 * bin_loss[0] = -(1/B) sum_b mean_t log_p[b, t, path_t]; d_log_p (optional, may be NULL) gets its
 * gradient.  workspace: B * T_text * T_feats doubles (query s2s_mas_workspace_bytes).
 * ------------------------------------------------------------------------------------------- */
size_t s2s_mas_workspace_bytes(int B, int T_feats, int T_text);
int s2s_mas(const float* log_p, const int32_t* text_lens, const int32_t* feats_lens, int B, int T_feats,
            int T_text, int32_t* paths, float* ds, float* bin_loss, float* d_log_p, void* workspace,
            size_t workspace_bytes, void* stream);

/* -------------------------------------------------------------------------------------------
 * STFT -> log-mel (bin/preprocess.py:30-92 + librosa semantics): wav (B, n_samples) float32 ->
 * mel (B, 1 + n_samples/hop, n_mels) float32.  center=True reflect padding, window (n_fft) float32
 * (periodic Hann zero-padded to n_fft by the caller), mel_basis (n_mels, 1 + n_fft/2) float32,
 * out = log10(max(eps, |STFT| @ basis^T)) (log_base: 10, 2 or 0 for natural).
 * n_fft must be a power of two <= 4096.
 * ------------------------------------------------------------------------------------------- */
int s2s_logmel(const float* wav, const float* window, const float* mel_basis, float* mel, int B, int n_samples,
               int n_fft, int hop, int n_mels, float eps, float log_base, void* stream);
/* same, with the global mean-variance normalisation of bin/normalize.py:173-193 (sklearn StandardScaler.transform) fused into
 * the store: mel[b,t,m] = (logmel - mean[m]) / scale[m]; mean / scale are (n_mels) float32 (bin/compute_statistics.py). */
int s2s_logmel_norm(const float* wav, const float* window, const float* mel_basis, const float* mean, const float* scale, float* mel,
                    int B, int n_samples, int n_fft, int hop, int n_mels, float eps, float log_base, void* stream);

/* Weight gradient of Conv2dSubsampling's first convolution through the GEMM path (bf16 engine): s2s_conv1_xcol builds the
 * (B T1 F1, 16) patch matrix of the INPUT x (B, T, F) float32 -- nine taps of Conv2d(1, C, 3, 2), a column of ones, six zero columns --
 * in `dtype`; s2s_gemm forms dy1^T x it into a (C, 16) float32 block; s2s_conv1_dw_scatter adds columns 0..8 into dw (C, 9) and
 * column 9 into dbias (C) (subsampling.py:58-63). */
int s2s_conv1_xcol(const float* x, void* xcol, int B, int T, int F, int dtype, void* stream);
int s2s_conv1_dw_scatter(const float* g16, float* dw, float* dbias, int C, void* stream);
/* forward of the same convolution on the GEMM path: w16 (C, 16) = [w (C, 9) | bias | 0 x 6] in `dtype`; y1 = relu(patches x w16^T) through
 * s2s_gemm (the ones column of the patch matrix carries the bias). */
int s2s_conv1_pack_w(const float* w, const float* bias, void* w16, int C, int dtype, void* stream);

/* Generic square-kernel / stride patch gather and its adjoint over channels-last maps (the later convolutions of
 * Conv2dSubsampling2 / 6 / 8, modules/transformer/subsampling.py:108-279: (k, s) = (3, 1), (5, 3), (3, 2)):
 *  s2s_im2col2d: col ((B T2 F2), k*k, C) with col[(b,t2,f2), kt*k+kf, c] = y[b, s t2 + kt, s f2 + kf, c]; T2 = (T1-k)/s + 1, F2 likewise.
 *  s2s_col2im2d: dy (B, T1, F1, C) = scatter-add of dcol; with gate != NULL (the map itself) times ReLU'(gate). */
int s2s_im2col2d(const void* y, void* col, int B, int T1, int F1, int C, int k, int s, int dtype, void* stream);
int s2s_col2im2d(const void* dcol, const void* gate, void* dy, int B, int T1, int F1, int C, int k, int s, int dtype, void* stream);

/* Dataset statistics for the global mean-variance normalisation (bin/compute_statistics.py:128-132: sklearn
 * StandardScaler.partial_fit over every utterance): feats (B, T, D) float32 zero-padded, lens (B) int32 valid frames per utterance
 * (NULL = all T), acc (2 * D + 1) float64 accumulated IN PLACE: acc[c] += sum_r x[r, c], acc[D + c] += sum_r x[r, c]^2,
 * acc[2 * D] += number of valid rows.  The host turns them into mean_ / var_ / scale_ (float64, as sklearn does). */
int s2s_feat_stats(const float* feats, const int* lens, double* acc, int B, int T, int D, void* stream);

/* Griffin-Lim phase reconstruction (vocoder/griffin_lim.py:52-106 -> librosa.griffinlim / stft / istft, center = True), fp32,
 * one utterance, n_fft a power of two in [64, 4096].  Spectra are (T, n_fft/2 + 1) row-major, complex values interleaved (re, im).
 *  s2s_gl_istft:  y (hop * (T - 1)) = istft(mag * angles): inverse rFFT of every frame (imaginary parts of the DC / Nyquist bins
 *                 ignored), synthesis window, overlap-add divided by the sum of squared windows, n_fft/2 trimmed on both sides;
 *                 frames (T, n_fft) is workspace.
 *  s2s_gl_stft:   spec = stft(y) of the T centred frames, padding n_fft/2 with zeros (pad_reflect = 0) or by reflection (1).
 *  s2s_gl_update: angles = rebuilt - c * tprev; angles /= |angles| + tiny; tprev = rebuilt   (n complex elements; c = momentum / (1 + momentum)). */
int s2s_gl_istft(const float* mag, const float* angles, const float* window, float* frames, float* y, int T, int n_fft, int hop,
                 void* stream);
int s2s_gl_stft(const float* y, const float* window, float* spec, int T, int64_t n_samples, int n_fft, int hop, int pad_reflect,
                void* stream);
int s2s_gl_update(const float* rebuilt, float* tprev, float* angles, int64_t n, float c, void* stream);

/* ===========================================================================================
 * Conformer block (AAS-VC encoder / decoder): modules/conformer/encoder_layer.py:79-179,
 * modules/conformer/convolution.py:56-79, modules/transformer/attention.py:209-305.
 * =========================================================================================== */

/* RelPositionMultiHeadedAttention query biases (attention.py:284-287): qu = q + pos_bias_u, qv = q + pos_bias_v.
 * q is (rows, d) with row stride ldq (a slice of the fused QKV buffer); u, v are (d) = (H, d_k) float32. */
int s2s_bias_add2(const void* q, int64_t ldq, const float* u, const float* v, void* qu, void* qv, int64_t rows, int d,
                  int dtype, void* stream);
/* out (row stride ldo) = a + b: joins d(qu) + d(qv) into the dQ slice of the fused dQKV buffer */
int s2s_add_strided(const void* a, const void* b, void* out, int64_t ldo, int64_t rows, int d, int dtype, void* stream);
/* rel_shift (attention.py:237-260) folded into a gather: S[b,h,i,j] += BD[h,b,i,T-1-i+j].  S (B,H,T,ldS) holds
 * matrix_ac (pre-scaled by the GEMM), BD (H,B,T,ldB >= 2T-1) holds matrix_bd before the shift. */
int s2s_relshift_add(void* S, const void* BD, int B, int H, int T, int64_t ldS, int64_t ldB, int dtype, void* stream);
/* adjoint of the gather: dBD[h,b,i,k] = dS[b,h,i,k-(T-1-i)] inside the band, 0 elsewhere (every column written) */
int s2s_relshift_bwd(const void* dS, void* dBD, int B, int H, int T, int64_t ldS, int64_t ldB, int dtype, void* stream);
/* LegacyRelPositionMultiHeadedAttention.rel_shift (modules/transformer/attention.py:138-157, the default of
 * VTN(encoder_type="conformer"), models/vtn.py:83-99): S (B,H,T,ldS) += shift(BD) with BD (H,B,T,ldB >= T) the T x T product
 * (q + pos_bias_v) p^T; the shift pads a zero column, re-views (T, T+1) as (T+1, T) and drops the first row, so rows wrap
 * around.  s2s_relshift_legacy_bwd is its adjoint (dBD from dS, dropped / padding elements zero). */
int s2s_relshift_legacy_add(void* S, const void* BD, int B, int H, int T, int64_t ldS, int64_t ldB, int dtype, void* stream);
int s2s_relshift_legacy_bwd(const void* dS, void* dBD, int B, int H, int T, int64_t ldS, int64_t ldB, int dtype, void* stream);
/* GLU over channels (convolution.py:69): x (rows, 2C) -> y (rows, C) = x[:, :C] * sigmoid(x[:, C:]); dx (rows, 2C) */
int s2s_glu_fwd(const void* x, void* y, int64_t rows, int C, int dtype, void* stream);
int s2s_glu_bwd(const void* dy, const void* x, void* dx, int64_t rows, int C, int dtype, void* stream);
/* depthwise Conv1d over time (convolution.py:37-45,72): x, y (B, T, C) channels-last, w (C, K) float32 (the
 * reference's (C,1,K) weight), bias (C) or NULL, zero padding (K-1)/2 per utterance; K odd.  The reference applies
 * no padding mask, neither do these.  bwd: dx (may be NULL), dw (C, K) += and dbias (C) += (may be NULL; dbias needs dw).
 * K = 7 / 15 / 31 take the shared-memory tiled kernels (one channel per thread, register sliding window). */
int s2s_dwconv_fwd(const void* x, const float* w, const float* bias, void* y, int B, int T, int C, int K, int dtype,
                   void* stream);
int s2s_dwconv_bwd(const void* dy, const void* x, const float* w, void* dx, float* dw, float* dbias, int B, int T, int C,
                   int K, int dtype, void* stream);
/* Swish FFN activation (conformer/swish.py:13-18) + dropout: y = dropout(x * sigmoid(x)); x is kept for backward */
int s2s_swish_fwd(const void* x, void* y, int64_t n, const s2s_dropout_t* drop, int dtype, void* stream);
int s2s_swish_bwd(const void* dy, const void* x, void* dx, int64_t n, const s2s_dropout_t* drop, int dtype,
                  void* stream);
/* y = x * scale * mask1 * mask2: RelPositionalEncoding's x * sqrt(d) between the input-layer dropout and the
 * positional dropout (layers/positional_encoding.py:303-309, conformer/encoder.py:117-123).  Its own adjoint. */
int s2s_scale_dropout(const void* x, void* y, int64_t n, float scale, const s2s_dropout_t* drop1,
                      const s2s_dropout_t* drop2, int dtype, void* stream);
/* y += alpha * x */
int s2s_axpy(const void* x, void* y, int64_t n, float alpha, int dtype, void* stream);
/* out[r, :] = s[r] * x[r, :] with s float32 per row */
int s2s_rowscale(const void* x, const float* s, void* out, int64_t rows, int C, int dtype, void* stream);
/* y[b,i,:] = sum_{j in [start[i], start[i]+count[i])} x[b,j,:]: nearest-neighbour F.interpolate over time
 * (models/aas_vc.py:339-351; count = 1) and its adjoint (runs of destination rows per source row). */
int s2s_gather_rows(const void* x, const int32_t* start, const int32_t* count, void* y, int B, int Tin, int Tout, int C,
                    int dtype, void* stream);
/* LengthRegulator (modules/length_regulator.py:46-97): y[b] = pad(repeat_interleave(x[b], ds[b])) as a gather over per-utterance
 * prefix sums.  s2s_lr_cumsum: cum (B, T+1) int32 = exclusive prefix sums of round(ds * alpha) (ds int64 as the reference's
 * LongTensor; alpha == 1: ds itself; all_ones != 0: every duration 1, the reference's rescue when ALL predicted durations are 0),
 * cum[b, T] = the utterance's output length.  s2s_lr_fwd: y (B, Lmax, D), frames past cum[b, T] = pad_value.  s2s_lr_bwd: the
 * adjoint, dx[b, i] = sum of dy over row i's run of frames. */
int s2s_lr_cumsum(const int64_t* ds, int32_t* cum, int B, int T, float alpha, int all_ones, void* stream);
int s2s_lr_fwd(const void* x, const int32_t* cum, void* y, int B, int T, int Lmax, int D, float pad_value, int dtype, void* stream);
int s2s_lr_bwd(const void* dy, const int32_t* cum, void* dx, int B, int T, int Lmax, int D, int dtype, void* stream);

/* ===========================================================================================
 * AAS-VC alignment block
 * =========================================================================================== */

/* AlignmentModule distance + log-softmax (modules/alignments.py:50-60): logp[b,t,s] =
 * log_softmax_s(-||feats[b,t,:] - text[b,s,:]||_2) with text padding (s >= text_lens[b]) = -inf.
 * feats (B,T_feats,C), text (B,T_text,C) in `dtype`; logp (B,T_feats,T_text) and lse (B,T_feats) float32
 * (lse = log sum_s exp(-dist), kept so that backward recovers dist = -(logp + lse)). */
int s2s_align_logp_fwd(const void* feats, const void* text, const int32_t* text_lens, float* logp, float* lse, int B,
                       int T_feats, int T_text, int C, int dtype, void* stream);
/* backward: W[b,t,s] = (d loss / d dist) / dist (B,T_feats,ldW) in `dtype`, rowsum (B,T_feats), colsum (B,T_text) so
 * that d_feats = rowsum * feats - W text and d_text = colsum * text - W^T feats (two s2s_gemm + s2s_rowscale). */
/* The same log-probabilities through the tensor-core product (bf16 engine): s2s_row_sqnorm gives |row|^2 (fp32) of feats / text,
 * s2s_gemm writes f . x into logp (B, T_feats, T_text) fp32, s2s_align_logp_from_dot turns it in place into
 * log_softmax_s(-sqrt(max(|f|^2 + |x|^2 - 2 f.x, 0))) with the same masking / lse output as s2s_align_logp_fwd. */
int s2s_row_sqnorm(const void* x, float* out, int64_t rows, int C, int dtype, void* stream);
int s2s_align_logp_from_dot(float* logp, const float* feats_sqnorm, const float* text_sqnorm, const int32_t* text_lens, float* lse,
                            int B, int T_feats, int T_text, void* stream);
int s2s_align_logp_bwd(const float* dlogp, const float* logp, const float* lse, const int32_t* text_lens, void* W,
                       float* rowsum, float* colsum, int B, int T_feats, int T_text, int64_t ldW, int dtype,
                       void* stream);
/* ForwardSumLoss (losses/forward_sum_loss.py:26-76): loss = mean_b [ CTC-NLL_b / N_b ] over the lattice
 * [blank,1,blank,...,N,blank] with emission logp + prior for labels and the constant blank_logp for blanks;
 * infeasible utterances contribute 0 (zero_infinity).  prior (B,T_feats,T_text) float32 is the beta-binomial
 * log-prior built on the host (forward_sum_loss.py:78-116).  alpha_ws: (B,T_feats,T_text) float32 workspace.
 * dlogp (may be NULL) receives grad_scale * d loss / d logp exactly as torch's ctc_loss backward defines it for
 * the reference: (exp(lp) - exp(alpha + beta - lp + nll)) / (N_b * B); entries outside (feats_len, text_len) = 0. */
int s2s_forward_sum(const float* logp, const float* prior, const int32_t* text_lens, const int32_t* feats_lens, int B,
                    int T_feats, int T_text, float blank_logp, float* alpha_ws, float* loss, float* dlogp,
                    float grad_scale, void* stream);
/* GaussianUpsampling weights (modules/length_regulator.py:111-154): P[b,t,:] = softmax_s(-delta (t' - c_s)^2),
 * c = cumsum(ds) - ds/2, t' = t for t < feats_lens[b] else 0, s >= text_lens[b] masked.  P (B,T_feats,ldP) `dtype`;
 * the expansion itself is s2s_gemm(P, hs). */
int s2s_gauss_weights(const float* ds, const int32_t* feats_lens, const int32_t* text_lens, void* P, int B, int T_feats,
                      int T_text, int64_t ldP, float delta, int dtype, void* stream);
/* DurationPredictor output masking/clamp + DurationPredictorLoss (modules/duration_predictor.py:98-101,
 * models/aas_vc.py:408-410, losses/duration_predictor_loss.py:29-50): d_outs = min(pre * mask, clamp_max),
 * loss = mean over s < text_lens[b] of (d_outs - log(ds + offset))^2, d_pre = grad_scale * d loss / d pre; when
 * g_douts (B,T_text) is given, d_pre = g_douts * mask * [pre <= clamp_max] instead (chain rule for an external loss
 * on d_outs).  d_outs, loss, d_pre may each be NULL. */
int s2s_duration_loss(const void* pre, const float* ds, const int32_t* text_lens, int B, int T_text, float offset,
                      float clamp_max, float grad_scale, const float* g_douts, float* d_outs, float* loss, void* d_pre,
                      int dtype, void* stream);
/* DurationPredictor.inference + clamp (modules/duration_predictor.py:92-96, models/aas_vc.py:389):
 * d[i] = min(max(rint(exp(pre[i]) - offset), 0), clamp_max) as float32 (integral values). */
int s2s_duration_infer(const void* pre, float* d, int n, float offset, float clamp_max, int dtype, void* stream);

/* ===========================================================================================
 * Single-position decode with a KV cache (VTN.inference, models/vtn.py:302-394; Decoder.forward_one_step,
 * modules/transformer/decoder.py:239-273).  Batch 1; the position is a DEVICE int32 so one CUDA graph serves all steps.
 * =========================================================================================== */
/* y[n] = act(W[n,:] . x + bias[n]) * dropout(pos * N + n) + residual[n];  W (N, K) row-major in `dtype`, x / y / residual
 * vectors in `dtype`, bias float32 (may be NULL), pos_dev may be NULL (position 0 for the dropout counter). */
int s2s_gemv(const void* W, const float* bias, const void* x, const void* residual, void* y, int N, int K, int relu,
             const s2s_dropout_t* drop, const int32_t* pos_dev, int dtype, void* stream);
/* ctx[h,:] = softmax_s(scale * q[h,:] . K[s,h,:]) V[s,h,:] over s < S with S = fixed_S (>= 0: source attention over the
 * encoder memory) or *pos_dev + 1 (self attention); when knew / vnew (H, d_k) are given they are first stored as cache row
 * *pos_dev.  Cache element (s, h, j) at cache + s * row_stride + h * d_k + j.  probs (float32, may be NULL) receives the
 * weights of this step at probs + *pos_dev * probs_step_stride + h * ldp (zero-padded to ldp).  S_cap bounds S. */
int s2s_decode_attn(const void* q, const void* knew, const void* vnew, void* kcache, void* vcache, int64_t row_stride, int H, int dk,
                    int fixed_S, int S_cap, const int32_t* pos_dev, float scale, void* ctx, float* probs, int ldp,
                    int64_t probs_step_stride, int dtype, void* stream);
/* y = x + alpha * pe[*pos_dev, :]  (ScaledPositionalEncoding of the new position, eval mode) */
int s2s_decode_pe(const void* x, const float* pe, const float* alpha, const int32_t* pos_dev, void* y, int d, int dtype, void* stream);
/* end of a step: frames[pos] = feat (r * odim), logits[pos] = logit (r), next decoder input = last generated frame, ++*pos_dev */
int s2s_decode_advance(const void* feat, const void* logit, void* next_in, float* frames, float* logits, int32_t* pos_dev, int odim,
                       int r, int dtype, void* stream);

/* -------------------------------------------------------------------------------------------
 * Stochastic duration predictor of AAS-VC (modules/duration_predictor.py:131-304, modules/vits/flow.py:19-310,
 * modules/vits/transform.py:12-216) -- the shipped recipe's default (egs/arctic/vc2/conf/aas_vc.melmelmel.v1.yaml:57).
 * Everything float32, channels-last (B, T, C) activations, flow state z as (B, 2, T); `tlens` (B) int32 is the text-length
 * mask.  The 1x1 convolutions are s2s_gemm, the channel LayerNorms s2s_layernorm_* (eps 1e-5).  Scalar log-likelihood terms
 * are ACCUMULATED into a per-utterance buffer nll (B) with a sign (atomicAdd; the caller threads one buffer through the chain).
 * ------------------------------------------------------------------------------------------- */
/* exact (erf) GELU: flow.py:165,178 */
int s2s_gelu_fwd(const float* x, float* y, int64_t n, void* stream);
int s2s_gelu_bwd(const float* dy, const float* x, float* dx, int64_t n, void* stream);
/* Conv1d(1 -> C, kernel 1) (ConvFlow.input_conv flow.py:240, post_pre duration_predictor.py:184): y[n, c] = x[n] w[c] + b[c];
 * bwd: dx[n] = sum_c dy[n, c] w[c] (dx may be NULL), dw (C) +=, db (C) += */
int s2s_outer_fwd(const float* x, const float* w, const float* b, float* y, int64_t N, int C, void* stream);
int s2s_outer_bwd(const float* dy, const float* x, const float* w, float* dx, float* dw, float* db, int64_t N, int C, void* stream);
/* dilated depthwise Conv1d over time on the MASKED input (flow.py:150-158,205): y[b,t,c] = bias[c] + sum_j w[c,j] *
 * xm[b, t + (j - (K-1)/2) * dil, c]; bwd: dx (masked), dw (C, K) +=, db (C) +=   (K in {3, 5, 7} for bwd) */
int s2s_dwconv_dilated_fwd(const float* x, const int32_t* tlens, const float* w, const float* bias, float* y, int B, int T, int C,
                           int K, int dil, void* stream);
int s2s_dwconv_dilated_bwd(const float* dy, const float* x, const int32_t* tlens, const float* w, float* dx, float* dw, float* db,
                           int B, int T, int C, int K, int dil, void* stream);
/* Piecewise rational-quadratic spline with linear tails, 10 bins on [-5, 5] (transform.py:44-209), one spline per position:
 * h (B, T, 29) = [10 widths | 10 heights | 9 derivatives] as ConvFlow.proj emits them (widths / heights are divided by
 * sqrt(hidden) inside, flow.py:296-300); x / y / gy / dx are rows of (B, 2, T) flow states given by a batch stride; padded
 * positions give y = 0, lad = 0.  inverse != 0 evaluates the inverse (lad negated, as the reference).  bwd: gradient of
 * sum(gy * y + glad * lad) w.r.t. x and h by forward-mode dual numbers (forward direction only; gy / glad may be NULL). */
int s2s_rq_spline_fwd(const float* x, int64_t x_bs, const float* h, const int32_t* tlens, float* y, int64_t y_bs, float* lad, int B,
                      int T, float hidden, int inverse, void* stream);
int s2s_rq_spline_bwd(const float* x, int64_t x_bs, const float* h, const int32_t* tlens, const float* gy, int64_t gy_bs,
                      const float* glad, float* dx, int64_t dx_bs, float* dh, int B, int T, float hidden, void* stream);
/* ElementwiseAffineFlow (flow.py:96-112) on z (B, 2, T): y = (m + exp(logs) z) mask (inverse: (z - m) exp(-logs) mask);
 * nll[b] += sign * sum_t mask (logs_0 + logs_1) when nll != NULL.  bwd (forward direction): dz, dm (2) +=, dlogs (2) +=. */
int s2s_sdp_affine_fwd(const float* z, const float* m, const float* logs, const int32_t* tlens, float* y, float* nll, float sign,
                       int B, int T, int inverse, void* stream);
int s2s_sdp_affine_bwd(const float* z, const float* logs, const int32_t* tlens, const float* gy, const float* g_nll, float sign,
                       float* dz, float* dm, float* dlogs, int B, int T, void* stream);
/* Posterior head + LogFlow (duration_predictor.py:262-281, flow.py:62-65): from z_q = (z_u, z1) and durations w (B, T):
 * u = sigmoid(z_u) mask, z0 = (w - u) mask, out = (log(max(z0, 1e-5)) mask, z1 mask),
 * nll[b] += sum_t mask (out_0 - logsigmoid(z_u) - logsigmoid(-z_u)). */
int s2s_sdp_head_fwd(const float* zq, const float* w, const int32_t* tlens, float* out, float* nll, int B, int T, void* stream);
int s2s_sdp_head_bwd(const float* zq, const float* w, const int32_t* tlens, const float* gout, const float* g_nll, float* dzq, int B,
                     int T, void* stream);
/* nll[b] += sign * sum_{ch,t} mask 0.5 (log 2 pi + z^2)   (duration_predictor.py:270-273,284-287); bwd: dz (+)= sign g_nll[b] z mask */
int s2s_sdp_gauss_fwd(const float* z, const int32_t* tlens, float* nll, float sign, int B, int T, void* stream);
int s2s_sdp_gauss_bwd(const float* z, const int32_t* tlens, const float* g_nll, float sign, float* dz, int accumulate, int B, int T,
                      void* stream);
/* acc[b] += sign * sum_t x[b, t]  /  out[b, t] = sign * g[b]  (the spline log-determinants and their gradient) */
int s2s_rowsum_acc(const float* x, float* acc, float sign, int B, int T, void* stream);
int s2s_rowbcast(const float* g, float* out, float sign, int B, int T, void* stream);
/* standard-normal draws (counter-based Box-Muller; *seed_dev is added to the seed so CUDA-graph replays differ) */
int s2s_randn(float* out, int64_t n, uint64_t seed, const uint64_t* seed_dev, uint64_t stream_id, void* stream);
/* dur[b, t] = min(ceil(exp(z[b, 0, t]) mask), clamp_max)   (duration_predictor.py:298-304, models/aas_vc.py:393) */
int s2s_sdp_durations(const float* z, const int32_t* tlens, float* dur, float clamp_max, int B, int T, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* S2SVC_B200_H_ */
