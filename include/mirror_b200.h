/*
 * mirror_b200 C ABI — the drop-in boundary of the B200-native MIRROR pre-training path.
 *
 * The reference (TianyiFranklinWang/MIRROR) is pure Python/PyTorch: every device
 * operation of its hot path is a torch library call.  This header
 * declares the entry points that replace those call sites; each comment cites the
 * reference lines (relative to the reference checkout) the entry point stands in for.
 *
 * Conventions:
 *  - plain pointers and sizes only; every pointer is DEVICE memory owned by the caller
 *    (PyTorch's caching allocator); the library never allocates, frees or retains them.
 *  - every call enqueues work on the caller's `stream` and returns immediately;
 *    no entry point synchronises the device or uses the legacy default stream.
 *  - return value 0 = success, otherwise a negative library code or a cudaError_t;
 *    `mirror_last_error()` returns a thread-local message.  No exceptions, no exit().
 *  - re-entrant and thread-safe: backward entry points are called from PyTorch's
 *    autograd thread concurrently with forward calls from the main thread.
 *  - "bf16" buffers are raw 16-bit bfloat16; "f32" are IEEE float.
 */
#ifndef MIRROR_B200_H
#define MIRROR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* mirror_stream_t; /* cudaStream_t */

const char* mirror_last_error(void);
int mirror_abi_version(void);
/* 1 when the current device is compute capability 10.x (tcgen05/TMEM present). */
int mirror_device_supported(void);
/* Graph-safe dropout: while a DEVICE counter is installed (NULL = off, the default), every dropout launch (GEMM epilogues,
 * mirror_act_fwd / mirror_act_bwd) mixes *device_counter into its seed at run time, so a replayed CUDA graph draws fresh
 * masks when the owner bumps the counter between replays (mirror_b200/step.py). */
int mirror_set_dropout_epoch(const void* device_counter);

/* ------------------------------------------------------------------------------------------------
 * Batched GEMM on 5th-gen tensor cores (tcgen05.mma, TMEM accumulators, TMA-fed, persistent).
 *
 *   for every (b2,b1):  acc[M,N] = sum_k A[m,k] * B[n,k]            (bf16 x bf16 -> fp32)
 *   v = alpha*acc + diag*[m==n] + bias[n];  v = act(v);  v = dropout(v);
 *   v += gamma * R[m,n] + gamma2 * R2[m,n] + beta * out_f32_old[m,n];   out_f32 = v;  out_bf16 = bf16(v)
 *
 * Replaces every nn.Linear / einsum / `@` of the hot path:
 *   models/mirror.py:346,654 (_fc1), :594-605 heads, :70-74,:100 RNA qkv/proj, timm Mlp
 *   (:138,:217,:815), :823-827 style heads/prototypes; nystrom_attention to_qkv / sim1-3 /
 *   pinv iterations / attn@v products / to_out (call site models/mirror.py:299-312);
 *   losses/mirror_loss.py:39-40 logits; and all of their autograd backward GEMMs.
 *
 * A is logically [M,K], B is logically [N,K] (C = A * B^T).  Each operand is either
 * "K-major" (K contiguous, `ld` = elements between consecutive m / n) or "MN-major"
 * (m / n contiguous, `ld` = elements between consecutive k), so all of NN/NT/TN/TT are covered.
 * Alignment: base pointers 16 B; ld and batch strides multiples of 8 elements.
 * ---------------------------------------------------------------------------------------------- */
enum { MIRROR_ACT_NONE = 0, MIRROR_ACT_RELU = 1, MIRROR_ACT_GELU = 2 };

typedef struct {
  const void* a;
  const void* b;
  int32_t a_mn_major, b_mn_major;
  int64_t lda, ldb;
  int64_t a_bs1, a_bs2, b_bs1, b_bs2; /* batch strides in elements */
  int32_t M, N, K, batch1, batch2;
  float alpha;
  const float* bias; /* [N] or NULL */
  int32_t act;
  float drop_p; /* 0 = off.  keep(idx) = hash(drop_seed, idx) >= p, idx = ((b2*batch1+b1)*M+m)*N+n */
  uint64_t drop_seed;
  const void* res; /* residual R, NULL = none */
  int32_t res_is_bf16;
  float gamma;
  int64_t ldr, r_bs1, r_bs2;
  float beta;
  float* out_f32; /* may be NULL */
  int64_t ldc32, c32_bs1, c32_bs2;
  void* out_bf16; /* may be NULL */
  int64_t ldc16, c16_bs1, c16_bs2;
  int32_t split_k; /* >1: K is split over CTAs and fp32 partials are atomically added into out_f32
                      (caller pre-zeroes it); only alpha is applied */
  float diag;      /* added to alpha*acc where m == n (e.g. E = I - a2.z of the Moore-Penrose step) */
  const void* res2; /* optional second residual (bf16, same element strides as `res`): v += gamma2 * R2[m,n] */
  float gamma2;
  int32_t res_row_div; /* >1: the residuals are read at row m / res_row_div (a row of R feeds res_row_div consecutive output
                          rows: the landmark-mean backward of the Nystrom layer) */
  int32_t mode;        /* MIRROR_GEMM_*: fused row-softmax epilogues (two GEMM passes, the logits never reach HBM) */
  float* stats;        /* [batch2,batch1,M,nparts] float2 partials exchanged between the two passes, nparts = mirror_gemm_nparts(N) */
} mirror_gemm_args;

/* Fused row softmax of the Nystrom similarity matrices and its backward, as epilogue modes of the
 * GEMM that produces the logits / the probability gradient.  Rows span several N tiles, so each takes two passes over the
 * (cheap, K = head_dim) product: pass 1 leaves per-row partials per half tile, pass 2 recomputes the product and applies them.
 *   ROWSTATS      stats[part] = (max, sum exp) of alpha*acc over the part's columns
 *   SOFTMAX       out = exp(alpha*acc - max) / sum          (bf16 and/or f32)
 *   ROWDOT        stats[part].x = sum_j acc_j * P_j over the part's columns, P = `res` (bf16 probabilities)
 *   SOFTMAX_BWD   out = alpha * P * (acc - sum of partial dots)   (bf16): d logits-before-alpha
 *   SOFTMAX_BWD_DOT  the same with the row dots given directly: stats = [batch2,batch1,M] f32.  When the probabilities
 *                 fed a product O = P V, sum_j G_ij P_ij = dO_i . O_i (mirror_rowdot_bf16), so the ROWDOT pass is not needed.
 * Needs N % 32 == 0; alpha is the only other epilogue term honoured. */
enum { MIRROR_GEMM_NORMAL = 0, MIRROR_GEMM_ROWSTATS = 1, MIRROR_GEMM_SOFTMAX = 2, MIRROR_GEMM_ROWDOT = 3, MIRROR_GEMM_SOFTMAX_BWD = 4,
       MIRROR_GEMM_SOFTMAX_BWD_DOT = 5 };
int mirror_gemm_nparts(int32_t N);

int mirror_gemm_bf16(const mirror_gemm_args* args, mirror_stream_t stream);
/* D = epilogue( sum_t A_t * B_t^T ), 1 <= nterms <= 6: the terms share M, N and the batch dims; K and the operand layouts
 * may differ.  terms[0] carries the epilogue and the outputs.  All products accumulate in the same TMEM tile, so a sum of
 * products costs one epilogue pass (used by the backward of the 6-step Moore-Penrose iteration of nystrom_attention, where
 * autograd would issue one fp32 accumulate pass per product). */
int mirror_gemm_bf16_multi(const mirror_gemm_args* terms, int32_t nterms, mirror_stream_t stream);
/* Same contract on CUDA cores (one thread per output element).  Test/diagnostic tool used by the
 * GPU unit tests to cross-check the tensor-core kernel at sizes where a host reference is slow. */
int mirror_gemm_bf16_simt(const mirror_gemm_args* args, mirror_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * HBM-bound glue (elementwise.cu).  `ld*` are row strides in elements.
 * ---------------------------------------------------------------------------------------------- */
/* dst[r,0:cols_out] = bf16(src[r,0:cols]), zero padded.  Replaces the implicit autocast casts
 * (train_mirror.py:1145) and `h.float()` (models/mirror.py:652). */
int mirror_cast_f32_bf16(const float* src, int64_t rows, int32_t cols, int64_t lds, void* dst, int32_t cols_out, int64_t ldd,
                         mirror_stream_t stream);
/* bf16 split-3 operand: hi = bf16(x), lo = bf16(x-hi); blocks (hi,lo,hi) (order 0) or (hi,hi,lo) (order 1), side by side
 * (stack_rows 0: dst [rows_out, 3*cols_out]) or stacked (stack_rows 1: dst [3*rows_out, cols_out]), zero padded.  One tensor-core
 * GEMM over the tripled contraction then gives hi.hi + lo.hi + hi.lo, i.e. an (almost) fp32 product.  Used where bf16 operand
 * rounding is visible in the result at negligible cost: _fc1 (models/mirror.py:654; its ReLU mask must match the fp32 reference,
 * every flipped mask bit is a 100 % error in that element's weight gradient — SURVEY.md §8c(v)), and every M = batch-sized Linear
 * (RNA encoder, heads, style/prototype heads, contrastive logits), which are weight-bandwidth-bound anyway. */
int mirror_cast_split3(const float* src, int64_t rows, int32_t cols, int64_t lds, void* dst, int64_t rows_out, int32_t cols_out,
                       int32_t stack_rows, int32_t order, mirror_stream_t stream);
/* dst[r,0:cols] = src[r,0:cols] with row strides: gathers `wsi_emb[:, 0, :]` (models/mirror.py:896) into a dense block */
int mirror_copy_rows_f32(const float* src, int64_t lds, int64_t rows, int32_t cols, float* dst, int64_t ldd, mirror_stream_t stream);
/* dst[r, 0:cols] (f32, dense) = src[idx[r], 0:cols], src fp32 or bf16 (src_bf16) with row stride lds, n_src rows: the per-slide
 * patch resampling of datasets/dataset_pretrain.py:157-161 over the packed features of a batch, on the device (mirror_b200/data.py) */
int mirror_gather_rows(const void* src, int32_t src_bf16, int64_t lds, int64_t n_src, const int64_t* idx, int64_t rows, int32_t cols,
                       float* dst, mirror_stream_t stream);
/* out[r] = sum_c a[r,c]*(b[r,c] - sub[r,c]) over contiguous bf16 rows (cols % 8 == 0, sub may be NULL): the row dots dO.O of a
 * softmax backward; `sub` removes a residual that was added to O after the attention product */
int mirror_rowdot_bf16(const void* a, const void* b, const void* sub, int64_t rows, int32_t cols, float* out, mirror_stream_t stream);
/* dst += alpha*src  (gradient accumulation of the autograd graph) */
int mirror_axpy_f32(float* dst, const float* src, int64_t n, float alpha, mirror_stream_t stream);
/* Gradient of the encoder output h:[B,T,E] that the model reads three ways (models/mirror.py:889-905: the whole matrix for the
 * retention decoder, row 0 for the alignment head and the style encoder, rows 1.. as the retention target):
 * out[b,0] = d_full[b,0] + d_cls[b];  out[b,t] = d_full[b,t] + d_tok[b,t-1].  Any of the three may be NULL (= zero).
 * d_tok is a [B,T-1,E] view with batch / row strides tok_bs / tok_ld. */
int mirror_token_fanout_bwd(const float* d_full, const float* d_cls, const float* d_tok, int64_t tok_bs, int64_t tok_ld, int32_t B,
                            int32_t T, int32_t E, float* out, mirror_stream_t stream);
/* out = dropout(act(pre)); keep(idx)=hash(seed,idx)>=p.  GELU of timm Mlp / Block (models/mirror.py:138-143,217-224). */
int mirror_act_fwd(const float* pre, int64_t n, int32_t act, float drop_p, uint64_t seed, void* out_bf16, float* out_f32,
                   mirror_stream_t stream);
/* dx = dy * dropmask * act'(pre) on [B,T,C] views (bs* = batch strides, ld* = row strides; dropout index = dense (b,t,c));
 * for ReLU `pre` may be the saved OUTPUT (same sign test). */
int mirror_act_bwd(const float* dy, int64_t bs_dy, int64_t ld_dy, const float* pre, int64_t bs_pre, int64_t ld_pre, int32_t B,
                   int32_t T, int32_t C, int32_t act, float drop_p, uint64_t seed, void* out_bf16, int64_t bs16, int64_t ld16,
                   float* out_f32, int64_t bs32, int64_t ld32, mirror_stream_t stream);
/* h:[B,1+N+add,E]; h[b,0]=cls, h[b,1+N+j]=h[b,1+j] (wrap-around square padding + cls, models/mirror.py:656-665) */
int mirror_wsi_assemble_fwd(float* h, const float* cls, int32_t B, int32_t N, int32_t add, int32_t E, mirror_stream_t stream);
/* backward of fc1+ReLU+assemble: dpre[b,j,:] = (h[b,1+j,:]>0) * (dh[b,1+j,:] + (j<add ? dh[b,1+N+j,:] : 0)) as bf16 [B,N,E];
 * dcls[e] += sum_b dh[b,0,e] */
int mirror_wsi_embed_bwd(const float* dh, const float* h, int32_t B, int32_t N, int32_t add, int32_t E, void* dpre_bf16, float* dcls,
                         mirror_stream_t stream);
/* mask[b,j] = rank(noise[b,j]) >= keep ? 1 : 0 -- argsort(argsort(noise)) of random_masking (models/mirror.py:516-531,630-647) */
int mirror_rank_mask(const float* noise, int32_t B, int32_t N, int32_t keep, float* mask, mirror_stream_t stream);
/* r[b,t,e] = (t>=first && mask[b,t-first] ? tok[e*tok_stride] : r[b,t,e]) + pos[t,e]  (models/mirror.py:521-527,636-643,692-693,549) */
int mirror_mask_pos_fwd(float* r, const float* mask, const float* tok, int32_t tok_stride, const float* pos, int32_t B, int32_t T,
                        int32_t E, int32_t first, mirror_stream_t stream);
/* dr = masked ? 0 : dy (dr must not alias dy); dpos[t,e] += sum_b dy; dtok[e*tok_stride] += sum of dy over the masked slots */
int mirror_mask_pos_bwd(const float* dy, const float* mask, float* dr, float* dtok, int32_t tok_stride, float* dpos, int32_t B,
                        int32_t T, int32_t E, int32_t first, mirror_stream_t stream);
/* Nyström landmarks: lm[b,j,0:2E] = mean of `seg` consecutive rows of the q and k slots of qkv[B,n,3E] */
int mirror_landmark_fwd(const void* qkv_bf16, void* lm_bf16, int32_t B, int32_t n, int32_t m, int32_t seg, int32_t E,
                        mirror_stream_t stream);
/* out[c] += sum_r x[r,c] (bias gradients) */
int mirror_colsum(const void* x, int32_t is_bf16, int64_t rows, int32_t cols, int64_t ld, float* out, mirror_stream_t stream);
/* z = mu + exp(0.5*logvar)*eps (Normal.rsample with injected eps, models/mirror.py:830-833) and its backward */
int mirror_reparam_fwd(const float* mu, const float* logvar, const float* eps, int64_t n, void* z_bf16, float* z_f32,
                       mirror_stream_t stream);
int mirror_reparam_bwd(const float* dz, const float* logvar, const float* eps, int64_t n, float* dmu, float* dlogvar,
                       mirror_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * LayerNorm / softmax / L2-normalise (norm_softmax.cu)
 * ---------------------------------------------------------------------------------------------- */
/* x:[B,x_rows,E] f32, the first S rows of every slide are normalised (x_rows > S: the encoder's final norm feeds only the
 * N+1 real tokens onward, models/mirror.py:372 `h[:, :-add_length]`, without a copy) -> outputs in a padded layout
 * [B,n_out,E] at row offset `pad` (rows<pad zero): the Nyström layer front-pads with zero rows.  mean/rstd: [B,S].
 * nn.LayerNorm at models/mirror.py:298,350,604 (eps 1e-5) and :122,137,255,494 (eps 1e-6). */
int mirror_layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, int32_t B, int32_t S, int32_t x_rows,
                         int32_t E, int32_t n_out, int32_t pad, void* out_bf16, float* out_f32, float* mean, float* rstd,
                         mirror_stream_t stream);
/* dx:[B,x_rows,E] = (add ? add : 0) + LN-gradient (rows >= S: no LN term); `add` may alias dx (the residual branch of
 * models/mirror.py:312) */
int mirror_layernorm_bwd(const void* dy /* f32, or bf16 when dy_is_bf16 */, int32_t dy_is_bf16, const float* x, const float* gamma,
                         const float* mean, const float* rstd, int32_t B, int32_t S, int32_t x_rows, int32_t E, int32_t n_out,
                         int32_t pad, float* dx, const float* add, float* dgamma, float* dbeta, mirror_stream_t stream);
/* row softmax of the three Nyström similarity matrices */
int mirror_softmax_fwd(const float* x, int64_t rows, int32_t cols, void* y_bf16, float* y_f32, mirror_stream_t stream);
int mirror_softmax_bwd(const void* y_bf16, const float* dy, int64_t rows, int32_t cols, float scale, void* dx_bf16, float* dx_f32,
                       mirror_stream_t stream);
/* F.normalize(x, dim=-1, eps) of the alignment heads (models/mirror.py:539-540,682-683); rows may be strided */
int mirror_l2norm_fwd(const float* x, int64_t ldx, int32_t rows, int32_t cols, float eps, void* y_bf16, float* y_f32, int64_t ldy,
                      float* norm, mirror_stream_t stream);
int mirror_l2norm_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* norm, int32_t rows, int32_t cols,
                      float* dx, int64_t lddx, int32_t accumulate, mirror_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Nyström attention specifics (nystrom.cu); algorithm of nystrom_attention~=0.0.14, call site models/mirror.py:299-312
 * ---------------------------------------------------------------------------------------------- */
/* out[b,t,c] = sum_j w[h(c),j] * v[b,t+j-16,c]  (res_conv: Conv2d(h,h,(33,1),groups=h,bias=False) on the value slot) */
int mirror_res_conv_fwd(const void* qkv_bf16, const float* w, int32_t B, int32_t n, int32_t E, void* out_bf16, mirror_stream_t stream);
/* dv_bf16[B,n,E] = conv^T(dout) (data gradient w.r.t. the value slot);  dw[8,33] += weight gradient */
int mirror_res_conv_bwd(const void* dout_bf16, const void* qkv_bf16, const float* w, int32_t B, int32_t n, int32_t E, void* dv_bf16,
                        float* dw, mirror_stream_t stream);
/* z0 = a2^T / (max_rowsum * max_colsum), maxima over the WHOLE [BH,m,m] tensor (moore_penrose_iter_pinv init).
 * scratch32: 32 bytes of device memory kept by the caller until the backward call. */
int mirror_pinv_init(const float* a2, int32_t BH, int32_t m, void* scratch32, float* z_f32, void* z_bf16, mirror_stream_t stream);
int mirror_pinv_init_bwd(const float* gz0, const void* z0_bf16, int32_t BH, int32_t m, void* scratch32, float* gx, int32_t accumulate,
                         mirror_stream_t stream);
/* The same followed by the row-softmax backward of attn2, fused: ds = scale * P * (g - rowsum(P g)) with
 * g = ga2 + (the pinv_init_bwd terms), P = a2_bf16; ga2 is left untouched.  m % 32 == 0, m <= 512. */
int mirror_pinv_init_softmax_bwd(const float* ga2, const float* gz0, const void* z0_bf16, const void* a2_bf16, int32_t BH, int32_t m,
                                 void* scratch32, float scale, void* ds_bf16, mirror_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * PPEG (ppeg.cu): models/mirror.py:317-331.  wm[49*E], bm[E], dwm[49*E], dbm[E] are caller-owned scratch.
 * ---------------------------------------------------------------------------------------------- */
int mirror_ppeg_fwd(const float* x, const float* w7, const float* w5, const float* w3, const float* b7, const float* b5,
                    const float* b3, int32_t B, int32_t H, int32_t E, float* wm, float* bm, float* y, mirror_stream_t stream);
int mirror_ppeg_bwd(const float* dy, const float* x, const float* wm, int32_t B, int32_t H, int32_t E, float* dx, int32_t accumulate,
                    float* dwm, float* dbm, float* dw7, float* dw5, float* dw3, float* db7, float* db5, float* db3,
                    mirror_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * RNA attention over the 12 chunks of one embedding (rna_attn.cu): models/mirror.py:77-102
 * ---------------------------------------------------------------------------------------------- */
int mirror_rna_attn_fwd(const float* qkv, int32_t B, int32_t E, void* out_bf16, float* out_f32, mirror_stream_t stream);
int mirror_rna_attn_bwd(const float* qkv, const float* dout, int32_t B, int32_t E, void* dqkv_bf16, float* dqkv_f32,
                        mirror_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Flash-style fused softmax products of the Nystrom attention (flash_nystrom.cu).  Replaces, per layer, the einsum +
 * softmax + einsum chains of nystrom_attention (call site models/mirror.py:299-312):
 *   out = softmax(q k_l^T) W + res_conv(v)   and   kv = softmax(q_l k^T) v
 * without the [n x m] / [m x n] probability matrices ever reaching HBM.
 *   out[b,h,r,:] = sum_j softmax_j(alpha x[b,h,r,:] . y[b,h,j,:]) v[b,h,j,:]  (+ res[b,h,r,:]);   lse2 = log2 sum_j 2^(alpha log2e x.y)
 * x: [batch, heads, R, d], y / v: [batch, heads, C, d] bf16 views addressed through (ld, head stride, batch stride) in
 * elements (multiples of 8; bases 16-byte aligned), d % 8 == 0, d <= 128.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const void* x;
  const void* y;
  const void* v;
  int64_t x_ld, x_hs, x_bs, y_ld, y_hs, y_bs, v_ld, v_hs, v_bs;
  int32_t R, C, d, heads, batch;
  float alpha;
  void* out; /* bf16 */
  int64_t o_ld, o_hs, o_bs;
  const void* res; /* bf16 or NULL */
  int64_t r_ld, r_hs, r_bs;
  float* lse2; /* [batch, heads, R] or NULL */
} mirror_flash_args;
int mirror_flash_softmax_pv(const mirror_flash_args* a, mirror_stream_t stream);

/* Backward of the product above by recomputation (autograd of the einsum / softmax / einsum chain), one orientation per call.
 * With P = 2^(alpha log2e S - lse2), dS = alpha P (dP - dot), dot[i] = dO_i . O_i:
 *   cols = 0 (tile = softmax rows):  S = a b^T (x y^T), dP = c dd^T (dO v^T);  out1 = dS b                 (= dX)
 *   cols = 1 (tile = keys):          S = a b^T (y x^T), dP = c dd^T (v dO^T);  out1 = dS b (= dY),  out2 = P dd (= dV)
 * a, c: [batch, heads, T, d];  b, dd: [batch, heads, L, d];  lse2 / dot: [batch, heads, softmax rows] (T for cols = 0, else L).
 * Each output is bf16 or f32 [batch, heads, T, d] through (ld, hs, bs) with an optional bf16 residual read at row / row_div and
 * scaled by rscale (the landmark-mean backward: token t receives d_landmark[t / l] / l; the value residual conv^T(dO)). */
typedef struct {
  void* ptr;
  int32_t is_f32;
  int64_t ld, hs, bs;
  const void* res;
  int64_t r_ld, r_hs, r_bs;
  int32_t row_div;
  float rscale;
} mirror_flash_out;
typedef struct {
  const void* a;
  const void* b;
  const void* c;
  const void* dd;
  int64_t a_ld, a_hs, a_bs, b_ld, b_hs, b_bs, c_ld, c_hs, c_bs, d_ld, d_hs, d_bs;
  int32_t T, L, d, heads, batch, cols;
  float alpha;
  const float* lse2;
  const float* dot;
  mirror_flash_out out1, out2;
} mirror_flash_bwd_args;
int mirror_flash_bwd(const mirror_flash_bwd_args* a, mirror_stream_t stream);
/* debug hook (measurement only, tools/flash_trace.py): CTA 0 of the flash kernels appends (event id << 48 | clock) entries to
 * buf[1..capacity), buf[0] counts them; buf = NULL switches the trace off (the default). */
int mirror_debug_flash_trace(void* buf, int64_t capacity);

/* ------------------------------------------------------------------------------------------------
 * Losses (loss.cu).  Loss values, the temperature scale and upstream gradients are DEVICE scalars.
 * ---------------------------------------------------------------------------------------------- */
/* Fused contrastive loss (contrastive.cu): the B x B logits L = (*scale) X Y^T live only in TMEM.
 * Replaces losses/mirror_loss.py:39-50 (two logits GEMMs + two cross entropies) and losses/info_nce.py:144-164, plus their
 * autograd backward (softmax gradients + four GEMMs).  X: [Br,K], Y: [Bc,K] bf16 row-major (ld in elements, multiples of 8);
 * the positive of row i is column i + diag0 (diag0 = rank * B_local when Y is the all-gathered global batch).
 * stats: lse[i] = ln sum_j e^{L_ij}, diag[i] = L_{i,i+diag0}; `part` is scratch of mirror_contrastive_nsplit(Br,Bc) * Br * 2 floats.
 * Column statistics are the row statistics of the swapped call (X <-> Y). */
int mirror_contrastive_nsplit(int32_t Br, int32_t Bc);
int mirror_contrastive_stats(const void* x, int64_t ldx, const void* y, int64_t ldy, int32_t Br, int32_t Bc, int32_t K,
                             const float* scale, int32_t diag0, float* part, int32_t nsplit, float* lse, float* diag,
                             mirror_stream_t stream);
/* dX[Br,E] (f32, ld lddx) = G Y_value with G_ij = s (a_r[i] e^{L_ij - lse_r[i]} + a_c[j] e^{L_ij - lse_c[j]}
 *                                                    - [j == i + diag0] (a_r[i] + a_c[j])), recomputed tile by tile;
 * *dscale += sum_ij (a_r[i] e^{L_ij - lse_r[i]} - [diag] a_r[i]) (X Y^T)_ij  (may be NULL).  a_c == NULL: one-sided loss.
 * precise = 0: K == D, Y_value = Y.  precise = 1: split-3 operands, K == 3 D, X = [hi|lo|hi], Y's hi block at column 0 and its
 * lo block at column lo_off; G is split into hi + lo too.  D: multiple of 64 (zero padded), E <= D columns are stored. */
int mirror_contrastive_grad(const void* x, int64_t ldx, const void* y, int64_t ldy, int32_t Br, int32_t Bc, int32_t K, int32_t D,
                            int32_t E, int32_t precise, int32_t lo_off, const float* scale, int32_t diag0, const float* lse_r,
                            const float* lse_c, const float* a_r, const float* a_c, float* dx, int64_t lddx, float* dscale,
                            mirror_stream_t stream);
/* per_sample[i] = w_r (lse_r[i] - diag[i]) + w_c (lse_c[i] - diag[i]) (may be NULL); *out = mult * sum_i (may be NULL) */
int mirror_contrastive_loss(const float* lse_r, const float* lse_c, const float* diag, int32_t B, float w_r, float w_c, float mult,
                            float* per_sample, float* out, mirror_stream_t stream);
/* a_r[i] = g[i*g_stride] * mult * w_r, a_c[i] = ... * w_c (a_c may be NULL): upstream gradient -> coefficients of G */
int mirror_contrastive_coef(const float* g, int32_t g_stride, int32_t B, float w_r, float w_c, float mult, float* a_r, float* a_c,
                            mirror_stream_t stream);
/* retention terms, losses/mirror_loss.py:98-103: out = sum_rows mask*mean_e (a-b)^2 / sum mask; batches strided by a_bs / b_bs */
int mirror_masked_mse_fwd(const float* a, int64_t a_bs, const float* b, int64_t b_bs, const float* mask, int32_t B, int32_t T,
                          int32_t E, float* scratch2, float* out, mirror_stream_t stream);
int mirror_masked_mse_bwd(const float* a, int64_t a_bs, const float* b, int64_t b_bs, const float* mask, int32_t B, int32_t T,
                          int32_t E, const float* scratch2, const float* gout, float gw, float* da, int64_t da_bs, int32_t acc_a,
                          float* db, int64_t db_bs, int32_t acc_b, mirror_stream_t stream);
/* style term, losses/mirror_loss.py:105-112, both modalities stacked ([2B,L], n = 2B*L) */
int mirror_gauss_kl_fwd(const float* mu, const float* logvar, int64_t n, int32_t B, float* out, mirror_stream_t stream);
int mirror_gauss_kl_bwd(const float* mu, const float* logvar, int64_t n, int32_t B, const float* gout, float gw, float* dmu,
                        float* dlogvar, mirror_stream_t stream);
/* cluster term, losses/mirror_loss.py:114-119: scores [2B,P], rows 0..B-1 WSI and B..2B-1 RNA */
int mirror_sym_kl_fwd(const float* scores, int32_t B, int32_t P, float* out, mirror_stream_t stream);
int mirror_sym_kl_bwd(const float* scores, int32_t B, int32_t P, const float* gout, float gw, float* dscores_f32, void* dscores_bf16,
                      mirror_stream_t stream);
/* total = sum_i w_i*term_i, losses/mirror_loss.py:121-127 (weights are host floats) */
int mirror_loss_combine(const float* terms5, const float* weights5_host, float* total, mirror_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Step tail (optim.cu) over FLAT fp32 buffers; every per-step scalar is read from device memory (graph-replayable).
 * Replaces optimizer.step() of train_mirror.py:1230 (opt: adam; torch.optim.Adam semantics, or AdamW with decoupled = 1),
 * clip_grad mode "norm" (:1222-1229) and the logit_scale clamp (:1254-1256).
 * ---------------------------------------------------------------------------------------------- */
/* p,g,m,v: [n] f32, 16-byte aligned.  *lr, *step (the 1-based update count, as float) and *grad_scale (may be NULL) are device
 * scalars: g is multiplied by *grad_scale first (the clip coefficient). */
int mirror_adam_step(float* p, const float* g, float* m, float* v, int64_t n, const float* lr, float beta1, float beta2,
                     float eps, float weight_decay, int32_t decoupled, const float* step, const float* grad_scale,
                     mirror_stream_t stream);
/* *out = sum g[i]^2 */
int mirror_grad_sumsq(const float* g, int64_t n, float* out, mirror_stream_t stream);
/* one launch for the scalar bookkeeping: *coef = min(1, max_norm / (sqrt(*sumsq) + 1e-6)) (1 when sumsq is NULL or
 * max_norm <= 0), *step += 1, *clamp_param = clamp(*clamp_param, lo, hi); each pointer may be NULL */
int mirror_tail_scalars(const float* sumsq, float max_norm, float* coef, float* step, float* clamp_param, float lo, float hi,
                        mirror_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MIRROR_B200_H */
