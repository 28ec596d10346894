// Classifier head fused with the softmax cross-entropy over the folded (sample x aspect) rows.
// Reference: FCMF.forward tail (fcmf_multimodal.py:50), criterion (run_multimodal_fcmf.py:290, 474-478).
// R = B*A rows (a few hundred) x C <= 32 classes: latency-bound, one warp per row.
#include "common.cuh"

namespace fcmf {

constexpr int HEAD_MAX_C = 32;

template <typename T>
__global__ void cls_ce_fwd_kernel(const T* __restrict__ pooled, const float* __restrict__ Wc, const float* __restrict__ bc,
                                  const int64_t* __restrict__ labels, float* __restrict__ logits,
                                  float* __restrict__ probs, float* __restrict__ loss_rows, int64_t R, int H, int C,
                                  fcmf_dropout drop) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  const T* x = pooled + r * H;
  const DropCfg dc = make_drop(drop);                 // dropout on the pooled vector (fcmf_multimodal.py:49)
  const uint32_t rseed = dc.thr16 ? drop_rowseed(dc.seed, (uint64_t)r) : 0u;
  float mine = 0.f;                                   // lane c keeps logit c
  for (int c = 0; c < C; ++c) {
    float s = 0.f;
    for (int k = lane; k < H; k += 32) {
      float xv = to_f(x[k]);
      if (dc.thr16) xv = drop_keep(rseed, (uint32_t)k, dc.thr16) ? xv * dc.inv_keep : 0.f;
      s = fmaf(xv, Wc[(int64_t)c * H + k], s);
    }
    s = warp_sum(s) + bc[c];
    if (lane == c) mine = s;
  }
  const float v = lane < C ? mine : -INFINITY;
  const float mx = warp_max(v);
  const float e = lane < C ? expf(v - mx) : 0.f;
  const float sum = warp_sum(e);
  if (lane < C) {
    logits[r * C + lane] = mine;
    if (probs) probs[r * C + lane] = e / sum;
  }
  if (labels && loss_rows) {
    const int64_t y = labels[r];
    const float picked = __shfl_sync(0xffffffffu, mine, (y >= 0 && y < C) ? (int)y : 0);
    if (lane == 0) loss_rows[r] = (y >= 0 && y < C) ? (mx + logf(sum) - picked) : 0.f;
  }
}

template <typename T>
__global__ void cls_ce_bwd_rows_kernel(const float* __restrict__ Wc, const float* __restrict__ probs,
                                       const int64_t* __restrict__ labels, const float* __restrict__ dlogits_in,
                                       float row_scale, float* __restrict__ dlogits, T* __restrict__ dpooled,
                                       int64_t R, int H, int C, fcmf_dropout drop) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  const DropCfg dc = make_drop(drop);
  const uint32_t rseed = dc.thr16 ? drop_rowseed(dc.seed, (uint64_t)r) : 0u;
  float dl = 0.f;
  if (lane < C) {
    if (dlogits_in) dl = dlogits_in[r * C + lane];
    else {
      const int64_t y = labels[r];
      dl = (y >= 0 && y < C) ? (probs[r * C + lane] - (lane == y ? 1.f : 0.f)) * row_scale : 0.f;
    }
    dlogits[r * C + lane] = dl;
  }
  for (int k = lane; k < H; k += 32) {
    float s = 0.f;
    for (int c = 0; c < C; ++c) s = fmaf(__shfl_sync(0xffffffffu, dl, c), Wc[(int64_t)c * H + k], s);
    if (dc.thr16) s = drop_keep(rseed, (uint32_t)k, dc.thr16) ? s * dc.inv_keep : 0.f;
    dpooled[r * H + k] = from_f<T>(s);
  }
}

template <typename T>
__global__ void cls_ce_bwd_w_kernel(const T* __restrict__ pooled, const float* __restrict__ dlogits,
                                    float* __restrict__ dWc, float* __restrict__ dbc, int64_t R, int H, int C,
                                    fcmf_dropout drop) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y;
  if (k >= H) return;
  const DropCfg dc = make_drop(drop);
  float s = 0.f, sb = 0.f;
  for (int64_t r = 0; r < R; ++r) {
    const float d = dlogits[r * C + c];
    float xv = to_f(pooled[r * H + k]);
    if (dc.thr16) xv = drop_keep(drop_rowseed(dc.seed, (uint64_t)r), (uint32_t)k, dc.thr16) ? xv * dc.inv_keep : 0.f;
    s = fmaf(d, xv, s);
    sb += d;
  }
  dWc[(int64_t)c * H + k] += s;
  if (k == 0) dbc[c] += sb;
}

}  // namespace fcmf

using namespace fcmf;

extern "C" int fcmf_cls_ce_fwd(const void* pooled, const float* Wc, const float* bc, const int64_t* labels,
                               float* logits, float* probs, float* loss_rows, int64_t R, int64_t H, int32_t C,
                               const fcmf_dropout* drop, int dtype, void* stream) {
  FCMF_CHECK_ARG(R >= 0 && H > 0 && C > 0 && C <= HEAD_MAX_C, "cls_ce_fwd: bad shape R=%lld H=%lld C=%d", (long long)R, (long long)H, C);
  FCMF_CHECK_ARG(drop_check(drop) == 0, "cls_ce_fwd: dropout p must be in [0, 1)");
  if (R == 0) return 0;
  cudaStream_t st = as_stream(stream);
  const unsigned grid = (unsigned)((R + 3) / 4);
  const fcmf_dropout dr = drop_or_off(drop);
  if (dtype == FCMF_BF16) cls_ce_fwd_kernel<bf16><<<grid, 128, 0, st>>>((const bf16*)pooled, Wc, bc, labels, logits, probs, loss_rows, R, (int)H, C, dr);
  else if (dtype == FCMF_F32) cls_ce_fwd_kernel<float><<<grid, 128, 0, st>>>((const float*)pooled, Wc, bc, labels, logits, probs, loss_rows, R, (int)H, C, dr);
  else return fail(FCMF_ERR_ARG, "cls_ce_fwd: bad dtype %d", dtype);
  FCMF_LAUNCH_OK();
  return 0;
}

extern "C" int fcmf_cls_ce_bwd(const void* pooled, const float* Wc, const float* probs, const int64_t* labels,
                               const float* dlogits_in, float row_scale, float* dlogits_ws, void* dpooled, float* dWc,
                               float* dbc, int64_t R, int64_t H, int32_t C, const fcmf_dropout* drop, int dtype, void* stream) {
  FCMF_CHECK_ARG(R >= 0 && H > 0 && C > 0 && C <= HEAD_MAX_C, "cls_ce_bwd: bad shape");
  FCMF_CHECK_ARG(drop_check(drop) == 0, "cls_ce_bwd: dropout p must be in [0, 1)");
  const fcmf_dropout dr = drop_or_off(drop);
  FCMF_CHECK_ARG(dlogits_in || (probs && labels), "cls_ce_bwd: need dlogits_in or (probs, labels)");
  FCMF_CHECK_ARG(dlogits_ws && dpooled && dWc && dbc, "cls_ce_bwd: null buffer");
  if (R == 0) return 0;
  cudaStream_t st = as_stream(stream);
  const unsigned grid = (unsigned)((R + 3) / 4);
  dim3 gw((unsigned)((H + 127) / 128), (unsigned)C);
  if (dtype == FCMF_BF16) {
    cls_ce_bwd_rows_kernel<bf16><<<grid, 128, 0, st>>>(Wc, probs, labels, dlogits_in, row_scale, dlogits_ws, (bf16*)dpooled, R, (int)H, C, dr);
    cls_ce_bwd_w_kernel<bf16><<<gw, 128, 0, st>>>((const bf16*)pooled, dlogits_ws, dWc, dbc, R, (int)H, C, dr);
  } else if (dtype == FCMF_F32) {
    cls_ce_bwd_rows_kernel<float><<<grid, 128, 0, st>>>(Wc, probs, labels, dlogits_in, row_scale, dlogits_ws, (float*)dpooled, R, (int)H, C, dr);
    cls_ce_bwd_w_kernel<float><<<gw, 128, 0, st>>>((const float*)pooled, dlogits_ws, dWc, dbc, R, (int)H, C, dr);
  } else return fail(FCMF_ERR_ARG, "cls_ce_bwd: bad dtype %d", dtype);
  FCMF_LAUNCH_OK();
  return 0;
}
