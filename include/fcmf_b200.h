/*
 * fcmf_b200 -- C ABI of the B200-native FCMF fusion path (sm_100a).
 *
 * The reference (sonbui25/Multimodal-Aspect-Category-Sentiment-Analysis) is pure Python/PyTorch and has
 * no FFI of its own; these entry points are what a torch custom-op layer binds
 * for the aten op groups on the hot path.  Each declaration names the reference code it replaces.
 *
 * Conventions: raw DEVICE pointers, caller-allocated outputs, row-major, leading dimensions in ELEMENTS,
 * stream-ordered on `stream` (a cudaStream_t), no implicit synchronisation, no global mutable state except
 * the launch counter and the attention-engine switch (TMA descriptors are encoded per call and passed by value).  Every function returns 0 on success or a
 * negative FCMF_ERR_* code; fcmf_last_error() returns the message of the last failure on this thread.
 * `dtype` is the storage type of activations (FCMF_F32 or FCMF_BF16); parameters that stay fp32
 * (biases, LayerNorm gamma/beta, statistics, weight gradients) are typed `float*`.
 */
#ifndef FCMF_B200_H_
#define FCMF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FCMF_ABI_VERSION 4

enum { FCMF_F32 = 0, FCMF_BF16 = 1 };
enum { FCMF_ERR_ARG = -1, FCMF_ERR_CUDA = -2, FCMF_ERR_UNSUPPORTED = -3 };
/* GEMM engines: AUTO picks tcgen05 for bf16 when the shape allows, the SIMT kernel otherwise. */
enum { FCMF_ENGINE_AUTO = 0, FCMF_ENGINE_SIMT = 1, FCMF_ENGINE_TCGEN05 = 2 };
/* GEMM epilogues */
enum {
  FCMF_EPI_NONE = 0,  /* D = acc + bias                                                         */
  FCMF_EPI_GELU = 1,  /* aux(out, optional) = acc + bias ; D = erf-GELU(aux)   mm_modeling.py:10-15,311-314 */
  FCMF_EPI_TANH = 2,  /* D = tanh(acc + bias)                                  mm_modeling.py:425-431       */
  FCMF_EPI_DGELU = 3  /* D = acc * dGELU/dx(aux(in))   (backward of BertIntermediate)                       */
};

/* Dropout of one site (the reference's nn.Dropout modules: mm_modeling.py:186,233,274,322, roi_modeling.py:77,
 * fcmf_multimodal.py:17). No mask tensor exists: keep(row, col) is a pure function of (seed + *seed_dev, row, col) that the
 * forward and the backward kernel both evaluate (csrc/common.cuh: drop_rowseed / drop_pair). p == 0 (or a NULL
 * fcmf_dropout pointer) = off; p is quantised to multiples of 1/65536; kept values are scaled by 1/(1-p).
 * seed_dev may be NULL; when set it is a DEVICE uint64 that the caller advances between steps (lets a replayed CUDA
 * graph draw fresh masks). */
typedef struct {
  float p;
  uint64_t seed;
  const uint64_t* seed_dev;
} fcmf_dropout;

/* 1 if element (row, col) of a dropout site with this p and EFFECTIVE seed (seed + *seed_dev) is kept, else 0. Host
 * evaluation of exactly the function the kernels run (lets a caller reproduce or audit a mask without a mask tensor). */
int fcmf_dropout_keep(float p, uint64_t seed, uint64_t row, uint32_t col);

int fcmf_abi_version(void);
/* sizeof / offsetof audit of the by-value structs (0 sizeof(fcmf_dropout), 1 .seed, 2 .seed_dev, 3 sizeof(fcmf_seg), 4 .idx,
 * 5 sizeof(fcmf_attn_desc), 6 .mask_add, 7 .bias, 8 .scale, 9 .causal, 10 .drop, 11 .engine, 12 fcmf_seg.groups; -1 otherwise): lets a binding check its
 * mirror of the layouts against the compiled library. */
int fcmf_abi_layout(int which);
const char* fcmf_last_error(void);
/* Number of kernels this library has launched in this process (bench.py reports it as gpu_launches). */
long long fcmf_kernel_launches(void);
/* sm count, compute capability; fails (FCMF_ERR_CUDA) when no device is present. */
int fcmf_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- dense contractions --------------------------------------------------------------------------- */
/* D[M,N] = epi(A[M,K] . B[N,K]^T + bias[N]).  Replaces nn.Linear forward (B = weight) and its input
 * gradient (B = weight^T): fcmf_pretraining.py:49-50,102-103 (vismap2text/roimap2text),
 * mm_modeling.py:194-196,241-243 (query/key/value), :277,:312,:325 (dense), :429 (pooler.dense),
 * roi_modeling.py:153-155,180 (box linears), fcmf_multimodal.py:50 (classifier via fcmf_cls_ce_*). */
int fcmf_gemm_tn(const void* A, int64_t lda, const void* B, int64_t ldb, const float* bias,
                 void* D, int64_t ldd, void* aux, int64_t ldaux,
                 int64_t M, int64_t N, int64_t K, int epi, int dtype, int engine, void* stream);

/* D(fp32)[M,N] = A[M,K] . B[N,K]^T, bf16 operands, the reduction split across CTAs (fp32 atomics; D is zeroed by the call).
 * For contractions with a small output and a long reduction: the input gradient of IAOGDecoder.dense (mm_modeling.py:645),
 * [B*T, H] over K = vocabulary. tcgen05 engine only (FCMF_ERR_UNSUPPORTED otherwise). */
int fcmf_gemm_tn_f32(const void* A, int64_t lda, const void* B, int64_t ldb, float* D, int64_t ldd, int64_t M, int64_t N,
                     int64_t K, int dtype, void* stream);

/* dW[N,K] (+)= dY[M,N]^T . X[M,K] ; db[N] (+)= column sums of dY (db may be NULL).  fp32 outputs.
 * Replaces autograd's weight/bias gradient of every nn.Linear above. */
int fcmf_gemm_wgrad(const void* dY, int64_t lddy, const void* X, int64_t ldx, float* dW, float* db,
                    int64_t M, int64_t N, int64_t K, int accumulate, int dtype, int engine, void* stream);

/* Planning query (host only, launches nothing): tiling of the tcgen05 weight-gradient GEMM for this shape -- cta_pairs = 1
 * when the CTA-pair kernel is used, tiles = 256x256 (or 128xBN) output tiles, splits = split-K factor, workers = CTA pairs (or
 * CTAs) that walk the tiles*splits work units in waves. Wave efficiency = units / (ceil(units / workers) * workers). */
int fcmf_gemm_wgrad_plan(int64_t M, int64_t N, int64_t K, int32_t* cta_pairs, int32_t* tiles, int32_t* splits,
                         int32_t* workers);

/* ---- row-wise kernels ------------------------------------------------------------------------------ */
/* y[m,:] = gamma * (s - mean)/sqrt(var + eps) + beta with s = x[m,:] + res[res_idx ? res_idx[m] : m, :]
 * (res may be NULL).  TF-style LayerNorm: biased variance, eps inside the sqrt (mm_modeling.py:166-171),
 * fused with the residual add of BertSelfOutput/BertOutput (mm_modeling.py:276-280, 324-328).
 * With dropout (mm_modeling.py:278, 326): s = dropout(x)[m,:] + res[...]; mask element = keep(row m, column). */
int fcmf_ln_fwd(const void* x, const void* res, const int32_t* res_idx, const float* gamma, const float* beta,
                void* y, float* mean, float* rstd, int64_t M, int64_t H, float eps, const fcmf_dropout* drop,
                int dtype, void* stream);
/* ds = dLN/ds . (dy + dy_add)  (same shape as x; dy_add may be NULL: the second gradient stream of a
 * residual fan-out); dgamma/dbeta[H] accumulate (+=) in fp32. With dropout, ds is the gradient of the residual
 * and dx (required then, ignored otherwise) = keep * ds / (1-p) is the gradient of x.
 * dy_every > 0: dy is compact [ceil(M / dy_every), H] and holds the gradient of rows 0, dy_every, 2*dy_every, ... only; every
 * other row's upstream gradient is zero (BertPooler keeps token 0 of each per-image branch, mm_modeling.py:428). 0 = dense. */
int fcmf_ln_bwd(const void* dy, const void* dy_add, const void* x, const void* res, const int32_t* res_idx, const float* gamma,
                const float* mean, const float* rstd, void* ds, void* dx, float* dgamma, float* dbeta,
                int64_t M, int64_t H, const fcmf_dropout* drop, int64_t dy_every, int dtype, void* stream);

/* add[b, j] = (1 - mask[b, j]) * -10000 for j < n   (fcmf_pretraining.py:53-56, 97-100, 133-136) */
int fcmf_mask_additive(const int64_t* mask, int64_t ldmask, float* add, int64_t rows, int64_t n, void* stream);

/* out[o,:] = sum_{g<G} src[idx[o*G+g], :]   (idx < 0 entries are skipped).  Covers concat/gather (G=1) of
 * fcmf_pretraining.py:114,127-131 and the reductions of shared rows in backward. out may accumulate. */
int fcmf_gather_sum_rows(const void* src, int64_t ldsrc, const int32_t* idx, void* out, int64_t ldout,
                         int64_t n_out, int64_t G, int64_t width, int accumulate, int dtype, void* stream);

/* out = dy * (1 - y*y)   (backward of the pooler tanh, mm_modeling.py:430) */
int fcmf_dtanh(const void* dy, const void* y, void* out, int64_t n, int dtype, void* stream);

/* dst(bf16 or f32) <- src(f32) with optional transpose of a [rows, cols] matrix (weight staging). */
int fcmf_cast_matrix(const float* src, void* dst, int64_t rows, int64_t cols, int transpose, int dtype, void* stream);
/* The same with a destination row stride ld_dst (elements) >= the destination's row length: stages a weight whose leading
 * extent is not a multiple of 8 (the 250 002-row vocabulary projection of the IAOG decoder, mm_modeling.py:645) into a
 * zero-padded buffer the tensor-core GEMM accepts; elements outside [rows, cols] are not written. */
int fcmf_cast_matrix_ld(const float* src, void* dst, int64_t rows, int64_t cols, int64_t ld_dst, int transpose, int dtype,
                        void* stream);
int fcmf_cast_to_f32(const void* src, float* dst, int64_t n, int dtype, void* stream);

/* ---- attention -------------------------------------------------------------------------------------- */
/* One launch covers NP problems x `heads` heads; the aspect dimension is folded into NP
 * (run_multimodal_fcmf.py:464-475 loop -> one launch).  Query rows are the concatenation of up to two
 * segments, keys/values likewise; segment s of problem p lives at  base + (idx[p]*rows + r)*ld + h*dh.
 * Scores = q.k/sqrt(dh) + mask_add[p / mask_div, j] + bias[p,h,i,j]; softmax; ctx = P.V.
 * Replaces BertCoAttention.forward (mm_modeling.py:240-266), BertSelfAttention.forward (:193-219) and
 * box_attention (roi_modeling.py:14-47). */
typedef struct {
  const void* ptr;      /* NULL => segment absent */
  int64_t ld;           /* row stride, elements */
  int32_t rows;         /* rows per group */
  int32_t groups;       /* number of groups the tensor holds (idx values lie in [0, groups)); 0 = unknown: the TMA-fed
                           attention kernels then are not used (they need the extent for their tensor maps) */
  const int32_t* idx;   /* [NP] group index of problem p; NULL => p */
} fcmf_seg;

typedef struct {
  fcmf_seg q[2], k[2], v[2];
  const float* mask_add; int64_t ld_mask; int32_t mask_div;   /* NULL => no mask */
  const float* bias;                                          /* [NP, heads, Lq, Lk] fp32 or NULL */
  int32_t NP, heads, dh;
  float scale;
  int32_t causal;       /* != 0: scores of keys j > query i are REPLACED by -1e4 (masked_fill of the IAOG decoder,
                           mm_modeling.py:115-124; applies to self- and cross-attention alike). CUDA-core engine only. */
  fcmf_dropout drop;    /* dropout on the probabilities AFTER the softmax (mm_modeling.py:213, 260; roi_modeling.py:42-43):
                           ctx = (keep * P / (1-p)) . V; mask element = keep(row (p*heads + h)*Lq + i, column j). */
  int32_t engine;       /* attention engine of THIS call (FCMF_ENGINE_*); 0 = the process default (fcmf_set_attn_engine).
                           Per-call so that concurrent host threads (the reference's nn.DataParallel fallback,
                           run_multimodal_fcmf.py:241-244) never share a mutable switch. */
} fcmf_attn_desc;

/* DEFAULT attention engine of the process (used by calls whose descriptor says engine = 0): FCMF_ENGINE_AUTO (tcgen05 for bf16, head_dim 64, no bias, 16 <= L <= 320;
 * CUDA-core otherwise), FCMF_ENGINE_SIMT or FCMF_ENGINE_TCGEN05 (fail if unsupported). Process-wide. */
int fcmf_set_attn_engine(int engine);
int fcmf_attn_fwd(const fcmf_attn_desc* d, void* ctx, int64_t ldctx, float* lse, int dtype, void* stream);
/* Per-problem gradients (the caller reduces rows shared between problems with fcmf_gather_sum_rows):
 * dq [NP, Lq, heads*dh], dk/dv [NP, Lk, heads*dh] in `dtype`, dbias [NP, heads, Lq, Lk] fp32 or NULL.
 * delta_ws: caller-provided scratch, fp32 [NP, heads, Lq]. */
int fcmf_attn_bwd(const fcmf_attn_desc* d, const void* ctx, int64_t ldctx, const void* dctx, int64_t lddctx,
                  const float* lse, float* delta_ws, void* dq, void* dk, void* dv, float* dbias, int dtype,
                  void* stream);

/* ---- geometric ROI relations (roi_modeling.py:79-138, 149-162, 40) ------------------------------------ */
/* boxes f64 [G, NR, 4] (x_min,x_max,y_min,y_max) -> emb f32 [G, NR, NR, 64] (computed in f64; the 8-entry
 * frequency table 1/1000^(k/8) is passed from the HOST in f32, exactly as the reference builds it, lines 123-125)
 * and bias[G, heads, NR, NR] = log(max(relu(wg_w[h].emb + wg_b[h]), 1e-6)). */
int fcmf_box_geometry_fwd(const double* boxes, const float* wg_w, const float* wg_b, const float* freq8_host,
                          float* emb, float* bias, int64_t G, int32_t NR, int32_t heads, void* stream);
/* d wg_w [heads,64], d wg_b [heads] (+=) from dbias [G, heads, NR, NR]; dz_ws: scratch of the same size as dbias. */
int fcmf_box_geometry_bwd(const float* emb, const float* wg_w, const float* wg_b, const float* dbias,
                          float* dz_ws, float* d_wg_w, float* d_wg_b, int64_t G, int32_t NR, int32_t heads,
                          void* stream);

/* ---- classifier head + loss (fcmf_multimodal.py:50, run_multimodal_fcmf.py:290,474-478) --------------- */
/* logits[R,C] = pooled[R,H] . Wc[C,H]^T + bc ; loss_rows[R] = CE(logits[r], labels[r]) (label outside [0,C) is
 * ignored); probs[R,C] saved for backward.  labels/probs/loss_rows may be NULL (logits only).  C <= 32.
 * drop: dropout on `pooled` before the classifier (fcmf_multimodal.py:49), mask = keep(row r, column k). */
int fcmf_cls_ce_fwd(const void* pooled, const float* Wc, const float* bc, const int64_t* labels,
                    float* logits, float* probs, float* loss_rows, int64_t R, int64_t H, int32_t C,
                    const fcmf_dropout* drop, int dtype, void* stream);
/* dlogits = dlogits_in if given, else (probs - onehot(labels)) * row_scale; written to dlogits_ws [R,C];
 * dpooled[R,H] in `dtype`; dWc[C,H] (+=), dbc[C] (+=) in fp32. */
int fcmf_cls_ce_bwd(const void* pooled, const float* Wc, const float* probs, const int64_t* labels,
                    const float* dlogits_in, float row_scale, float* dlogits_ws, void* dpooled, float* dWc,
                    float* dbc, int64_t R, int64_t H, int32_t C, const fcmf_dropout* drop, int dtype, void* stream);

/* ---- wide softmax cross-entropy: the IAOG decoder's vocabulary loss (run_pretraining_fcmf.py:320-322) --------- */
/* logits [R, V] (row stride ld elements, any 2/4-byte row alignment), labels [R] int64; rows whose label equals
 * ignore_index (or lies outside [0, V)) contribute 0. loss_rows[r] = logsumexp(logits[r]) - logits[r, label]; lse[r] is
 * saved for the backward pass. The caller divides the sum by the number of counted rows (CrossEntropyLoss 'mean'). */
int fcmf_vocab_ce_fwd(const void* logits, int64_t ld, const int64_t* labels, int64_t ignore_index,
                      float* loss_rows, float* lse, int64_t R, int64_t V, int dtype, void* stream);
/* dlogits[r, j] = (softmax(logits[r])[j] - [j == label[r]]) * scale[0]; scale is a DEVICE scalar (upstream gradient /
 * counted rows, so no host synchronisation is needed); dlogits may alias logits (in place). */
int fcmf_vocab_ce_bwd(const void* logits, int64_t ld, const int64_t* labels, int64_t ignore_index,
                      const float* lse, const float* scale, void* dlogits, int64_t ldd, int64_t R, int64_t V,
                      int dtype, void* stream);

/* ---- optimizer tail (run_multimodal_fcmf.py:483-489: clip_grad_norm_(1.0) + torch.optim.AdamW.step) -------------- */
/* One table entry per parameter tensor (fp32): the 4 parameter groups of the reference only differ in lr / weight_decay. */
typedef struct {
  float* p; const float* g; float* m; float* v;   /* parameter, gradient, exp_avg, exp_avg_sq */
  int64_t n;
  float lr, wd;
} fcmf_opt_tensor;
/* table: DEVICE array of fcmf_opt_tensor; (blk_tensor[b], blk_chunk[b]): block b handles elements
 * [chunk*8192, (chunk+1)*8192) of that tensor. sumsq[0] = sum of squares of every gradient element. */
int fcmf_opt_sumsq(const void* table, const int32_t* blk_tensor, const int32_t* blk_chunk, int64_t n_blocks,
                   float* sumsq, void* stream);
/* coef[0] = min(1, max_norm / (sqrt(sumsq[0]) + 1e-6)) (max_norm <= 0: 1); norm_out[0] = sqrt(sumsq[0]) (may be NULL). */
int fcmf_opt_clip_coef(const float* sumsq, float max_norm, float* coef, float* norm_out, void* stream);
/* g *= coef[0] (coef may be NULL); AdamW update of every tensor of the table at optimizer step `step` (>= 1). */
int fcmf_opt_adamw(const void* table, const int32_t* blk_tensor, const int32_t* blk_chunk, int64_t n_blocks,
                   const float* coef, float beta1, float beta2, float eps, int64_t step, int write_back_grad, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FCMF_B200_H_ */
