/*
 * mmi_b200.h -- C ABI of libmmi_b200.so: hand-written sm_100a CUDA kernels for the
 * MMinterest training step (hezy18/SegMMInterest, MMinterest/).
 *
 * The reference has no FFI: its hot path is eager PyTorch.  Each entry point below
 * replaces a group of ATen/cuBLAS library calls; the reference call site it replaces
 * is cited as file:line relative to /root/reference/MMinterest/.
 *
 * Conventions
 *   - all pointers are DEVICE pointers owned by the caller (workspace included);
 *   - every call is asynchronous on `stream` (a cudaStream_t) and never syncs the host;
 *   - row-major everywhere, leading dimensions in ELEMENTS;
 *   - dtype codes: MMI_F32 = 0, MMI_BF16 = 1 (activations / table); parameters,
 *     gradients, optimizer state, LayerNorm statistics and reductions are fp32;
 *   - return value: 0 on success, negative MMI_E* otherwise; mmi_last_error() gives
 *     a thread-local message.  No C++ exception crosses this boundary.
 */
#ifndef MMI_B200_H_
#define MMI_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMI_F32 0
#define MMI_BF16 1

#define MMI_OK 0
#define MMI_EINVAL (-1)   /* bad shape / dtype / alignment */
#define MMI_ECUDA (-2)    /* CUDA runtime error at launch   */
#define MMI_ENOSUP (-3)   /* configuration not implemented  */

#define MMI_ACT_NONE 0
#define MMI_ACT_GELU 1    /* erf GELU, kn_util/nn_utils/layers/mlp.py:30-31 */
#define MMI_ACT_RELU 2    /* MLP_Block of the SelfMLP / CrossMLP ablations, models/encoder.py:210-252 */

/* GEMM operand layouts (mmi_gemm): C[M,N] = op(A) * op(B)                          */
#define MMI_GEMM_NT 0     /* A [M,K] row-major, B [N,K] row-major  (y = x W^T)       */
#define MMI_GEMM_NN 1     /* A [M,K] row-major, B [K,N] row-major  (dx = dy W)       */
#define MMI_GEMM_TN 2     /* A [K,M] row-major, B [K,N] row-major  (dW = dy^T x)     */

/* GEMM implementations */
#define MMI_IMPL_SIMT 0   /* fp32-FFMA CUDA-core tiles (strict-parity mode)          */
#define MMI_IMPL_TC 1     /* bf16 tcgen05.mma + TMEM accumulators + TMA operands     */

typedef void* mmi_stream_t; /* cudaStream_t */

/* ---- dropout (every nn.Dropout(0.1) site of the path: models/encoder.py:386,472 embedding; :145-150 attention logits,
 * BEFORE the 1/sqrt(dh) scale; :163-164 attention output projection; kn_util/nn_utils/layers/mlp.py:21-22 after the FFN's
 * GELU; models/encoder.py:198-202 FFN output).  Counter-based: element (row, col) of a site is kept iff
 * bit (col & 31) of keep_word(key, row, col >> 5) is set (segmminterest_b200/csrc/dropout.cuh; numpy twin in
 * oracle/dropout_ref.py), so forward and backward regenerate the same mask and none is stored.  thr8 = 0 switches the
 * site off; otherwise the drop probability is thr8 / 256 and scale must be 256 / (256 - thr8).                      */
typedef struct {
  uint32_t key;    /* per site and per step (host: segmminterest_b200/dropout.py site_key)  */
  uint32_t thr8;   /* 0 = off                                                                */
  float scale;     /* 1 / realised keep probability                                          */
} mmi_dropout;
/* test hook: mask[r * cols + c] = 1 iff element (row0 + r, c) is kept (group of column c = group0 + (c >> 5)).       */
int mmi_dropout_mask(const mmi_dropout* drop, int64_t row0, int64_t rows, int cols, uint32_t group0, uint8_t* mask,
                     mmi_stream_t stream);

int mmi_version(void);
const char* mmi_last_error(void);
/* 1 if the tcgen05 GEMM path was compiled in and the device is sm_100 */
int mmi_has_tc(void);

/* ---- a-1..a-3: gather + zero-pad + mask (+ L1 normalise) ---------------------------
 * replaces utils/dataloader_SegMM.py:301-350 (row gather, _pad_feature_list :251-268)
 * and main_for_seq_leave_earlystop_SegMM.py:272-273 (x / (||x||_1 + 1e-6)).
 * idx[n_tokens] int32 row ids into table[n_rows, din], -1 = pad (row of zeros, mask 0).
 * normalise = 0 gives a bit-exact copy. out_dtype may differ from table_dtype.       */
int mmi_gather_l1norm_fwd(const void* table, int table_dtype, int64_t n_rows, int din,
                          const int32_t* idx, int64_t n_tokens, void* out, int out_dtype,
                          uint8_t* mask, int normalise, mmi_stream_t stream);

/* ---- generic GEMM with fused epilogue ----------------------------------------------
 * replaces every nn.Linear on the path (models/encoder.py:50-62,95-98,163-164,432-445;
 * kn_util/nn_utils/layers/mlp.py:17-24) and their autograd mm/addmm backward.
 *   acc  = op(A) op(B)                                   (fp32 accumulate)
 *   z    = acc + bias[n]                                 (bias may be NULL)
 *   if preact: preact[m,n] = z                           (saved for GELU backward; save_act_grad = 0)
 *              preact[m,n] = gelu'(z)                    (save_act_grad = 1: backward needs no erf)
 *   y    = act(z)
 *   if drop.thr8: y = dropout(y) (and the gelu'(z) written to preact gets the same mask and scale, so the backward
 *                 epilogue's y *= preact applies the dropout gradient for free); element index (m, n), BEFORE `add`
 *   if mul_gelu_grad: y *= gelu'(mul_gelu_grad[m,n])     (dgrad through GELU; mul_is_grad = 0)
 *                     y *= mul_gelu_grad[m,n]            (mul_is_grad = 1: operand saved by save_act_grad)
 *                     y *= mul_gelu_grad[m,n] > 0 ? mul_scale : 0   (mul_is_grad = 2: the operand is the OUTPUT of a
 *                          ReLU (+ dropout) layer, dgrad through dropout(relu(z)): kept and positive <=> output > 0)
 *   if add:  y += add[(m % add_mod) * ld_add + n]        (residual / position embedding)
 *   C    = y                (accumulate == 0)
 *   C   += y                (accumulate == 1, C must be fp32; used for weight grads)
 * in_dtype applies to A, B, add, preact, mul_gelu_grad; out_dtype to C.
 * The tcgen05 path (MMI_IMPL_TC) stores bf16 outputs through shared memory with TMA
 * (cp.async.bulk.tensor) and prefetches the add / mul_gelu_grad tile the same way whenever the
 * operands allow it (bf16, 16-byte aligned rows, add_mod >= M); other cases take a generic
 * register epilogue with identical results.
 * split_k > 1 is only legal with accumulate == 1 (atomic fp32 adds); split_k == 0 lets
 * the tensor-core path pick a split that fills the 148 SMs.                          */
typedef struct {
  int layout;       /* MMI_GEMM_* */
  int impl;         /* MMI_IMPL_* */
  int in_dtype;     /* dtype of A and B */
  int out_dtype;    /* dtype of C */
  int64_t M;
  int64_t N;
  int64_t K;
  const void* A; int64_t lda;
  const void* B; int64_t ldb;
  void* C; int64_t ldc;
  const float* bias;
  int act;
  void* preact; int64_t ld_preact;            /* in_dtype */
  const void* mul_gelu_grad; int64_t ld_mul;  /* in_dtype */
  const void* add; int64_t ld_add; int64_t add_mod; int add_dtype;
  int accumulate;
  int split_k;
  int save_act_grad;   /* 1: preact receives gelu'(z) instead of z */
  int mul_is_grad;     /* 1: mul_gelu_grad already holds gelu'(z); 2: it holds a ReLU (+ dropout) output */
  mmi_dropout drop;    /* applied to act(z) before `add`; thr8 = 0: off (not with accumulate / split-K) */
  float mul_scale;     /* mul_is_grad == 2: survivors' scale of the dropout that followed the ReLU (1 = none) */
  /* fp32 operands on the tensor cores (impl = MMI_IMPL_TC with in_dtype = MMI_F32): A and B are split into three bf16
   * terms each (x = hi + mid + lo, 24 mantissa bits) in this caller-owned workspace and the six products
   * hi*hi, hi*mid, mid*hi, mid*mid, hi*lo, lo*hi accumulate in the fp32 TMEM accumulator -- fp32-grade results (the
   * dropped terms are below 2^-24) at ~1/6 of the bf16 rate instead of the FFMA path's 1/40.
   * At least mmi_gemm_split_workspace(layout, M, N, K) bytes, 1024-byte aligned.                                     */
  void* split_ws;
  int64_t split_ws_bytes;
} mmi_gemm_args;
int mmi_gemm(const mmi_gemm_args* args, mmi_stream_t stream);
int64_t mmi_gemm_split_workspace(int layout, int64_t M, int64_t N, int64_t K);

/* ---- AdaptiveAvgPool1d(out_len) along the token axis (the CrossMLP ablation, models/encoder.py:395,504-506):
 * x [B, L, d] -> y [B, out_len, d], window j = [floor(j L / out_len), ceil((j + 1) L / out_len)).
 * bwd: dx[b, t, :] = sum over the windows that contain t of dy[b, j, :] / window length.                       */
int mmi_adaptive_pool_fwd(const void* x, int dtype, int B, int L, int d, int out_len, void* y, mmi_stream_t stream);
int mmi_adaptive_pool_bwd(const void* dy, int dtype, int B, int L, int d, int out_len, void* dx, mmi_stream_t stream);

/* column sums: out[n] += sum_m X[m,n]  (bias gradients).  out is fp32.               */
int mmi_colsum_acc(const void* x, int dtype, int64_t M, int N, int64_t ldx, float* out,
                   float* workspace, int64_t workspace_floats, mmi_stream_t stream);

/* ---- LayerNorm(eps) over the last dim, fp32 statistics ------------------------------
 * replaces torch.nn.LayerNorm(d, 1e-12) at models/encoder.py:39-40,185-186,383-385.
 * stats[row] = (mean, rstd).                                                          */
int mmi_layernorm_fwd(const void* x, int dtype, int64_t rows, int d, const float* gamma,
                      const float* beta, float eps, void* y, float* stats, mmi_stream_t stream);
/* dx = LN'(dy) (+ add); dgamma += sum dy*xhat; dbeta += sum dy; if dxsum: dxsum[c] += sum_rows dx[row, c]
 * (the bias gradient of the Linear that produced the LayerNorm input, fused so dx is not read again).
 * workspace: at least mmi_layernorm_bwd_workspace(d) floats.                          */
int64_t mmi_layernorm_bwd_workspace(int d);
int mmi_layernorm_bwd(const void* dy, const void* x, int dtype, int64_t rows, int d,
                      const float* gamma, const float* stats, const void* add, void* dx,
                      float* dgamma, float* dbeta, float* dxsum, float* workspace, mmi_stream_t stream);

/* Dropout variants (NULL pointers = that part off; with all three NULL they equal the calls above):
 * fwd:  y = dropout(LN(x))                                   -- the embedding dropout, models/encoder.py:386,472
 * bwd:  dy_drop:  dy is read as dropout'(dy) = mask * scale * dy (backward of the forward variant);
 *       dx_drop + dx_dropped: additionally writes dx_dropped = mask * scale * dx, the gradient that flows into the
 *       Linear whose dropped-out output was added to the residual (encoder.py:163-171,198-202), and dxsum then sums
 *       dx_dropped (that Linear's bias gradient); dx itself stays the residual-branch gradient.                      */
int mmi_layernorm_fwd_drop(const void* x, int dtype, int64_t rows, int d, const float* gamma, const float* beta,
                           float eps, void* y, float* stats, const mmi_dropout* drop, mmi_stream_t stream);
int mmi_layernorm_bwd_drop(const void* dy, const void* x, int dtype, int64_t rows, int d, const float* gamma,
                           const float* stats, const void* add, void* dx, float* dgamma, float* dbeta, float* dxsum,
                           float* workspace, const mmi_dropout* dy_drop, const mmi_dropout* dx_drop, void* dx_dropped,
                           mmi_stream_t stream);

/* ---- a-5/a-6: candidate x history attention -----------------------------------------
 * replaces models/encoder.py:44-73 (get_attn_logits: QK^T, outer-product mask,
 * "set to -10000"), :138-161 (concat of two key blocks, /sqrt(dh), joint softmax, PV).
 * One call handles ONE query side (candidate or history queries) over TWO key blocks
 * with different query projections:
 *     S = [ Qa Ka^T | Qb Kb^T ],  S[!(mq_i & mk_j)] = -10000,  P = softmax(S / sqrt(dh)),
 *     O = P [Va ; Vb]
 * With drop.thr8 != 0 the reference's logits dropout (encoder.py:145-150) sits between the fill and the scale:
 *     S <- dropout(S)  (dropped logits become 0 -- a dropped masked logit is therefore VISIBLE to the softmax, exactly
 *     like the reference), mask row = (b*H + h)*Lq + q, column group of key k of block i = (i << 20) + (k >> 5).
 * Tensors are [B*L, ld] with head h at columns [h*dh, (h+1)*dh).  nblk may be 1.      */
typedef struct {
  const void* q; int64_t ldq;      /* query projection used against this key block */
  const void* k; int64_t ldk;
  const void* v; int64_t ldv;
  const uint8_t* mask_k;           /* [B, Lk] */
  int Lk;
  /* backward outputs (same layout as q/k/v); may be NULL in forward */
  void* dq; int64_t lddq;
  void* dk; int64_t lddk;
  void* dv; int64_t lddv;
  /* optional (NULL = off; MMI_IMPL_TC only): fp32 [H*dh] accumulators that receive (+=, atomics) the column sums of dq /
   * dk / dv taken in fp32 BEFORE rounding to the output dtype -- the bias gradients of the three projections
   * (autograd's dY.sum(0) for models/encoder.py:50-62,95-98), so no separate pass re-reads the gradients.        */
  float* dbq; float* dbk; float* dbv;
} mmi_attn_block;
typedef struct {
  int dtype; int impl;
  int B; int H; int dh; int Lq;
  const uint8_t* mask_q;           /* [B, Lq] */
  int nblk; mmi_attn_block blk[2];
  void* out; int64_t ldo;          /* [B*Lq, H*dh] */
  float* lse;                      /* [B, H, Lq]  log-sum-exp of scaled logits */
  /* backward only */
  const void* dout; int64_t lddo;
  float* delta;                    /* [B, H, Lq] scratch: rowsum(dO * O) */
  mmi_dropout drop;                /* logits dropout; thr8 = 0: off */
  /* mmi_attn_bwd_fused only: fp32 accumulators of the partial dQ tiles of block i, [B*Lq, H*dh] (row-major, leading
   * dimension H*dh), and one int32 counter per (b, h).  Both must be ZERO on entry and are zero again when the call has
   * finished (the last key-tile CTA of every (b, h) converts its dQ columns to blk[i].dq and clears what it read).  */
  float* dq_acc[2];
  int32_t* dq_count[2];
} mmi_attn_args;
int mmi_attn_fwd(const mmi_attn_args* a, mmi_stream_t stream);
/* writes dq for both blocks and delta */
int mmi_attn_bwd_dq(const mmi_attn_args* a, mmi_stream_t stream);
/* writes dk, dv of block `which` (needs lse and delta from the calls above) */
int mmi_attn_bwd_dkv(const mmi_attn_args* a, int which, mmi_stream_t stream);
/* MMI_IMPL_TC, bf16, dh 32: the whole backward of key block `which` in ONE kernel -- dk, dv (owned by the CTA's 128 keys)
 * AND that block's dq (partial tiles reduced through dq_acc[which]) AND their bias-gradient sums; S, dP and the
 * exponential are computed once per score instead of once in mmi_attn_bwd_dq and once in mmi_attn_bwd_dkv.  Needs out,
 * lse and dout (delta is recomputed on the fly and not written).  Returns 1 (no error set) when the configuration is
 * not covered and the caller should use the two-kernel path.                                                         */
int mmi_attn_bwd_fused(const mmi_attn_args* a, int which, mmi_stream_t stream);
/* MMI_IMPL_TC, bf16, dh 32, at most 5 key tiles of 128 over both blocks (640 keys): the WHOLE backward of the query side in
 * one launch, one CTA per (b, h) that owns every key of both blocks -- dq, dk, dv of both blocks and their bias-gradient
 * sums, nothing accumulated through global memory (models/encoder.py:138-161 backward).  Needs out, lse, dout; delta is
 * recomputed and not written.  Returns 1 (no error set) when the shapes are not covered.                            */
int mmi_attn_bwd_all(const mmi_attn_args* a, mmi_stream_t stream);

/* ---- a-9: head Linear(d -> 1)  (models/decoder_leave_focal.py:451,596) ---------------
 * logits[r] = w . x[r] (+ b[0]) (+ add[r]);  b and add may be NULL.  `add` chains two heads for the two-backbone
 * fusions without InteractionAggregation (:624-631: Linear over the sum / the concatenation, or two Linear heads).
 * head_bwd: db may be NULL when another call owns the bias gradient.                                   */
int mmi_head_fwd(const void* x, int dtype, int64_t rows, int d, const float* w,
                 const float* b, const float* add, float* logits, mmi_stream_t stream);
int64_t mmi_head_bwd_workspace(int d);
int mmi_head_bwd(const void* x, int dtype, int64_t rows, int d, const float* w,
                 const float* dlogits, const float* gscale /* device scalar or NULL */,
                 void* dx, float* dw, float* db, float* workspace, mmi_stream_t stream);

/* ---- a-10..a-12: loss --------------------------------------------------------------
 * replaces models/decoder_leave_focal.py:490-572 for loss_type `focal`
 * (my_sigmoid_focal_loss :35-59, alpha 0.5, gamma 2, masked sum / bsz) plus the
 * diagnostics mse / mse2 (:552-558).  gt int64 [B,L] in {1,0,-1,-2}; gt is rewritten in
 * place exactly as the reference does (:534-535) when rewrite_gt != 0.
 * scalars (fp32[16], layout under mmi_loss_fwd_bwd): 0 focal, 1 mse, 2 mse2, 3 loss(=weight*focal).
 * dlogits[B,L] = d loss / d logits (already includes weight and inv_bsz).             */
int mmi_focal_loss_fwd_bwd(const float* logits, int64_t* gt, int B, int L,
                           const float* exposure_prob, float inv_bsz, float weight,
                           int rewrite_gt, float* scalars, float* dlogits, mmi_stream_t stream);

/* General form: optional learnable position bias (models/decoder_leave_focal.py:442-444,497-504:
 * logits += (pos+1) * bias_weight + bias_bias, both [L]) and the default loss `interestBPR`
 * (compute_interest_BPR_all :163-221: rows with view_len < L; pos = logits[row, view_len], the other L-1 logits --
 * pad positions included -- are negatives; -log(clamp(sum_k softmax_k(neg) * sigmoid(neg_k - pos), 1e-8, 1-1e-8)),
 * mean over those rows, times bpr_scale (1 / world size under data parallelism)).
 * The other selectable losses of compute_loss (:539-551) ride in the same launch:
 *   huber      huber_loss(sum hazard_masked [B], view_lengths [B,1], delta 1) -- broadcasts to [B,B] like `mse` (:61-66)
 *   hazard     compute_partial_likelihood_loss (:273-286), rows with view_len == L skipped, / B
 *   surviveCE  compute_leave_prob_CE (:68-97): BCE-with-logits fed exp(h_t), sum over valid / number of valid positions
 *   interestCE / interestKL  compute_interest_leave_CE (:99-161) with use_mask = mask_loss; `*_after_focal` != 0 when
 *              'focal' precedes the loss in loss_type_list (it then sees focal's in-place rewrite of gt, :534-535).
 * loss = sum over the losses switched on of w_* x value  (w_huber is loss_weight['mse'], :563-564).  Batch means over a
 * data-dependent count (interestBPR, surviveCE) and huber's [B,B] mean are taken over this call's rows and scaled by
 * bpr_scale (1 / world size under data parallelism); focal, hazard, interestCE/KL use inv_bsz (1 / global batch).
 * scalars (fp32[16]): 0 focal, 1 mse, 2 mse2, 3 loss, 4 interestBPR, 5 huber, 6 hazard, 7 surviveCE, 8 interestCE,
 * 9 interestKL (10..15 reserved).
 * logits_out (optional) receives logits + bias; dbias_* (optional) are accumulated (+=).             */
typedef struct {
  const float* logits; int64_t* gt; int B; int L;
  const float* exposure_prob;
  const float* bias_weight; const float* bias_bias;
  float inv_bsz; float w_focal; float w_bpr; float bpr_scale;
  int use_focal; int use_bpr; int rewrite_gt;
  float* logits_out; float* scalars; float* dlogits;
  float* dbias_weight; float* dbias_bias;
  int use_huber; int use_hazard; int use_surviveCE; int use_interestCE; int use_interestKL;
  int mask_loss; int ce_after_focal; int kl_after_focal;
  float w_huber; float w_hazard; float w_surviveCE; float w_interestCE; float w_interestKL;
} mmi_loss_args;
int mmi_loss_fwd_bwd(const mmi_loss_args* args, mmi_stream_t stream);

/* ---- ID-embedding inputs and the bilinear fusion head ------------------------------------
 * mmi_id_embed_fwd: models/encoder.py:352-362,426-435,478-488.  out[b,l,c] = (c < tw ? table[ids[b], c]
 *   : pos * frame_w[c-tw] + frame_b[c-tw]) + pe[l,c]   (video: tw = d/2, L = 40; user: tw = d, L = 1, no frame
 *   projection).  table fp32 [n_rows, tw], ids int64 [B], pe fp32 [L, d] or NULL, out [B*L, d].  pos = l, or
 *   frame_pos[b*L + l] (fp32 [B, L], may be NULL): the 'noPos' ablation feeds a random permutation of the frame
 *   positions per interaction (models/encoder.py:428-429).
 * mmi_id_embed_bwd: dtable[ids[b], c] += sum_l de[b,l,c]; dframe_w += sum_{b,l} pos*de; dframe_b += sum de (atomics). */
int mmi_id_embed_fwd(const float* table, int64_t n_rows, int tw, const int64_t* ids, int B, int L, int d,
                     const float* frame_w, const float* frame_b, const float* pe, const float* frame_pos, void* out,
                     int out_dtype, mmi_stream_t stream);
int mmi_id_embed_bwd(const void* de, int dtype, const int64_t* ids, int64_t n_rows, int tw, int B, int L, int d,
                     float* dtable, float* dframe_w, float* dframe_b, const float* frame_pos, mmi_stream_t stream);
/* Row-sparse table gradient for data parallelism (the reference has no DP and keeps nn.Embedding's dense gradient):
 * mmi_id_rows_bwd writes rows[b, c] = sum_l de[b,l,c] (c < tw; fp32 [B, tw]) instead of adding into the table, the frame
 * projection's gradients are accumulated as in mmi_id_embed_bwd; after the ranks have exchanged (ids, rows),
 * mmi_scatter_rows_add does dtable[ids[i], :] += rows[i, :] for the gathered n = world x B rows.                  */
int mmi_id_rows_bwd(const void* de, int dtype, int tw, int B, int L, int d, float* rows, float* dframe_w, float* dframe_b,
                    const float* frame_pos, mmi_stream_t stream);
int mmi_scatter_rows_add(const int64_t* ids, const float* rows, int64_t n, int tw, int64_t n_rows, float* dtable, mmi_stream_t stream);
/* InteractionAggregation (models/decoder_leave_focal.py:411-423) after the two X_h W_h GEMMs:
 *   out[r] = sum_c T[r,c] * Y[r,c] (+ add1[r]) (+ add2[r]);   backward: dT = g*Y, dY = g*T (+ dy_add), g *= gscale[0]. */
int mmi_rowdot_fwd(const void* t, int64_t ldt, const void* y, int64_t ldy, int dtype, int64_t R, int C,
                   const float* add1, const float* add2, float* out, mmi_stream_t stream);
int mmi_rowdot_bwd(const float* g, const float* gscale, const void* t, int64_t ldt, const void* y, int64_t ldy,
                   int dtype, int64_t R, int C, const void* dy_add, void* dt, void* dy, mmi_stream_t stream);

/* ---- a-14: global-norm clip + AdamW on flat fp32 buffers ----------------------------
 * replaces main_for_seq_leave_earlystop_SegMM.py:298-299 (clip_grad_norm_(10.0),
 * torch.optim.AdamW.step).  norm_out[0] = pre-clip global L2 norm, norm_out[1] = clip
 * coefficient.  bf16_out (optional) receives the bf16 copy of the updated parameters.  */
int64_t mmi_clip_adamw_workspace(int64_t n);
int mmi_clip_adamw(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                   int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
                   float max_norm, int step, float* norm_out, void* bf16_out,
                   float* workspace, mmi_stream_t stream);

/* ---- a-15 / 8f-4: validation metrics on the device -------------------------------------------------
 * replaces models/my_evaluation.py: ProbAUC_batch (:73-80) and the per-row metrics of main_eval_batch (:264-357:
 * LeaveMSE's predict_view_length :82-85, LeaveCTR / LeaveCTR_view :87-90, JaccardSim = IoU_Sim length_aware :37-57),
 * fed with interests = sigmoid(logits) * exposure_prob (main...SegMM.py:402-403) and
 * survival = exp(cumsum(log interests)) (my_evaluation.py:273-274).
 * input_kind 0: `logits` are logits (interests formed here); 1: `logits` already holds the interests (exposure_prob
 * unused, may be NULL) -- what main_eval_batch receives.
 * logits fp32 [B,L], gt int64 [B,L] in {1,0,-1,-2} (or the {1,0,-2} focal leaves behind), L <= 64.
 * rows fp32 [B,6]: pred_view_length, view_length, duration, LeaveCTR, LeaveCTR_view, JaccardSim.
 * out fp32[4]: 0 ProbAUC over all positions with gt != -2 (label gt==-1 -> 0; exact Mann-Whitney statistic with
 * ties 1/2 == sklearn.roc_auc_score; NaN when one class is missing, where sklearn raises), 1 n_pos, 2 n_neg.
 * workspace: mmi_eval_metrics_workspace(B, L) bytes, 8-byte aligned.                                            */
int64_t mmi_eval_metrics_workspace(int B, int L);
int mmi_eval_metrics(const float* logits, const int64_t* gt, int B, int L, const float* exposure_prob,
                     int input_kind, void* workspace, float* rows, float* out, mmi_stream_t stream);

/* fp32 -> bf16 cast, optionally transposed: src [rows, cols] -> dst [cols, rows]      */
int mmi_cast_bf16(const float* src, void* dst, int64_t rows, int64_t cols, int transpose,
                  mmi_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MMI_B200_H_ */
