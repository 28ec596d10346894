// Single-query folded attention (Lq == 1): the "live rows" form of the fusion path, where only the [CLS] query of
// every (sample, aspect, image) problem reaches an output (BertPooler keeps token 0, mm_modeling.py:425-431).
// One warp = one (problem, head); a block = consecutive problems of one head. No shared-memory staging of K/V (each key row is
// touched twice by one warp and served by L1/L2); lanes map to KEYS for the score / dP dot products (16-byte row loads, q and dO broadcast from
// shared memory) and to head DIMENSIONS for the P.V / dS.K accumulations and for every global store (coalesced
// 128-byte rows). fp32 math, bf16 or fp32 storage, head_dim <= 128, Lk <= Q1_MAX_LK.
#include "common.cuh"
#include "attn.h"

namespace fcmf {

constexpr int Q1_WARPS = 4;
constexpr int Q1_MAX_LK = 512;
constexpr int Q1_MAX_DH = 128;

template <typename T>
__device__ __forceinline__ float dot_row(const T* __restrict__ row, const float* __restrict__ vec, int dh) {
  constexpr int N = Vec16<T>::N;
  float s = 0.f;
  for (int c = 0; c < dh; c += N) {
    Vec16<T> k;
    k.load(row + c);
#pragma unroll
    for (int j = 0; j < N; ++j) s = fmaf(k.v[j], vec[c + j], s);
  }
  return s;
}

template <typename T>
__global__ void __launch_bounds__(Q1_WARPS * 32)
attn_q1_fwd_kernel(AttnDev a, T* __restrict__ ctx, int64_t ldctx, float* __restrict__ lse) {
  __shared__ float qs[Q1_WARPS][Q1_MAX_DH];
  __shared__ float ps[Q1_WARPS][Q1_MAX_LK];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // a block = Q1_WARPS CONSECUTIVE problems of ONE head: in the folded layout consecutive problems are the images of one (sample, aspect)
  // and read the same text K/V rows, so the block's warps share them through L1 instead of each pulling them from L2
  const int h = (int)(blockIdx.x % (unsigned)a.heads);
  const int p = (int)(blockIdx.x / (unsigned)a.heads) * Q1_WARPS + warp;
  if (p >= a.NP) return;
  const int dh = a.dh, Lk = a.Lk;
  float* q = qs[warp];
  float* pr = ps[warp];
  const T* qrow = seg_row<T>(a.q, p, 0, h, dh);
  for (int d = lane; d < dh; d += 32) q[d] = to_f(qrow[d]);
  __syncwarp();
  const float* madd = a.mask_add ? a.mask_add + (int64_t)(p / a.mask_div) * a.ld_mask : nullptr;
  float mx = -INFINITY;
  for (int j = lane; j < Lk; j += 32) {
    float s = dot_row<T>(seg_row<T>(a.k, p, j, h, dh), q, dh) * a.scale;       // scale before the mask add
    if (madd) s += madd[j];
    pr[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  const DropCfg dc = make_drop(a.drop);
  const uint32_t rseed = dc.thr16 ? drop_rowseed(dc.seed, attn_drop_row(a, p, h, 0)) : 0u;
  for (int j = lane; j < Lk; j += 32) {
    const float e = __expf(pr[j] - mx);
    sum += e;                                                    // denominator before dropout
    pr[j] = (dc.thr16 && !drop_keep(rseed, (uint32_t)j, dc.thr16)) ? 0.f : e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  const float inv = (dc.thr16 ? dc.inv_keep : 1.0f) / sum;
  T* orow = ctx + (int64_t)p * ldctx + (int64_t)h * dh;
  for (int d = lane; d < dh; d += 32) {
    float o = 0.f;
#pragma unroll 4
    for (int j = 0; j < Lk; ++j) o = fmaf(pr[j], to_f(seg_row<T>(a.v, p, j, h, dh)[d]), o);
    orow[d] = from_f<T>(o * inv);
  }
  if (lane == 0 && lse) lse[(int64_t)p * a.heads + h] = mx + __logf(sum);
}

template <typename T>
__global__ void __launch_bounds__(Q1_WARPS * 32)
attn_q1_bwd_kernel(AttnDev a, const T* __restrict__ ctx, int64_t ldctx, const T* __restrict__ dctx, int64_t lddctx,
                   const float* __restrict__ lse, T* __restrict__ dq, T* __restrict__ dk, T* __restrict__ dv) {
  __shared__ float qs[Q1_WARPS][2 * Q1_MAX_DH];      // q | dO
  __shared__ float ps[Q1_WARPS][2 * Q1_MAX_LK];      // P | dS
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // a block = Q1_WARPS CONSECUTIVE problems of ONE head: in the folded layout consecutive problems are the images of one (sample, aspect)
  // and read the same text K/V rows, so the block's warps share them through L1 instead of each pulling them from L2
  const int h = (int)(blockIdx.x % (unsigned)a.heads);
  const int p = (int)(blockIdx.x / (unsigned)a.heads) * Q1_WARPS + warp;
  if (p >= a.NP) return;
  const int dh = a.dh, Lk = a.Lk, HD = a.heads * dh;
  float* q = qs[warp];
  float* go = q + Q1_MAX_DH;
  float* P = ps[warp];
  float* dS = P + Q1_MAX_LK;
  const T* qrow = seg_row<T>(a.q, p, 0, h, dh);
  const T* orow = ctx + (int64_t)p * ldctx + (int64_t)h * dh;
  const T* grow = dctx + (int64_t)p * lddctx + (int64_t)h * dh;
  float dl = 0.f;
  for (int d = lane; d < dh; d += 32) {
    q[d] = to_f(qrow[d]);
    const float g = to_f(grow[d]);
    go[d] = g;
    dl = fmaf(g, to_f(orow[d]), dl);
  }
  dl = warp_sum(dl);                                               // delta = dO . O
  __syncwarp();
  const float l = lse[(int64_t)p * a.heads + h];
  const float* madd = a.mask_add ? a.mask_add + (int64_t)(p / a.mask_div) * a.ld_mask : nullptr;
  const DropCfg dc = make_drop(a.drop);
  const uint32_t rseed = dc.thr16 ? drop_rowseed(dc.seed, attn_drop_row(a, p, h, 0)) : 0u;
  for (int j = lane; j < Lk; j += 32) {
    float s = dot_row<T>(seg_row<T>(a.k, p, j, h, dh), q, dh) * a.scale;
    if (madd) s += madd[j];
    const float pj = __expf(s - l);
    float dp = dot_row<T>(seg_row<T>(a.v, p, j, h, dh), go, dh);
    float pd = pj;
    if (dc.thr16) {
      const bool keep = drop_keep(rseed, (uint32_t)j, dc.thr16);
      pd = keep ? pj * dc.inv_keep : 0.f;
      dp = keep ? dp * dc.inv_keep : 0.f;
    }
    P[j] = pd;                                                     // dropped probability: dV = P_drop^T . dO
    dS[j] = pj * (dp - dl) * a.scale;                              // scale folded in: dq and dk both carry it
  }
  __syncwarp();
  T* dqrow = dq + (int64_t)p * HD + (int64_t)h * dh;
  for (int d = lane; d < dh; d += 32) {
    const float qd = q[d], gd = go[d];
    float acc = 0.f;
#pragma unroll 4
    for (int j = 0; j < Lk; ++j) {
      const float ds = dS[j];
      acc = fmaf(ds, to_f(seg_row<T>(a.k, p, j, h, dh)[d]), acc);
      const int64_t o = ((int64_t)p * Lk + j) * HD + (int64_t)h * dh + d;
      dk[o] = from_f<T>(ds * qd);
      dv[o] = from_f<T>(P[j] * gd);
    }
    dqrow[d] = from_f<T>(acc);
  }
}

bool attn_q1_supported(const AttnDev& a) {
  if (a.Lq != 1 || a.bias != nullptr || a.causal || a.Lk > Q1_MAX_LK || a.dh > Q1_MAX_DH) return false;
  const int vec = 8;                                               // 16-byte row loads for bf16 (4 floats for f32)
  if (a.dh % vec) return false;
  for (int s = 0; s < 2; ++s) {
    const SegDev* segs[2] = {&a.k[s], &a.v[s]};
    for (const SegDev* g : segs)
      if (g->ptr && g->rows && ((reinterpret_cast<uintptr_t>(g->ptr) & 15u) || (g->ld % vec))) return false;
  }
  return true;
}

int attn_q1_fwd(const AttnDev& a, void* ctx, int64_t ldctx, float* lse, int dtype, cudaStream_t st) {
  const unsigned grid = (unsigned)(((int64_t)a.NP + Q1_WARPS - 1) / Q1_WARPS * a.heads);
  if (dtype == FCMF_BF16) attn_q1_fwd_kernel<bf16><<<grid, Q1_WARPS * 32, 0, st>>>(a, (bf16*)ctx, ldctx, lse);
  else attn_q1_fwd_kernel<float><<<grid, Q1_WARPS * 32, 0, st>>>(a, (float*)ctx, ldctx, lse);
  FCMF_LAUNCH_OK();
  return 0;
}

int attn_q1_bwd(const AttnDev& a, const void* ctx, int64_t ldctx, const void* dctx, int64_t lddctx, const float* lse,
                void* dq, void* dk, void* dv, int dtype, cudaStream_t st) {
  const unsigned grid = (unsigned)(((int64_t)a.NP + Q1_WARPS - 1) / Q1_WARPS * a.heads);
  if (dtype == FCMF_BF16)
    attn_q1_bwd_kernel<bf16><<<grid, Q1_WARPS * 32, 0, st>>>(a, (const bf16*)ctx, ldctx, (const bf16*)dctx, lddctx, lse,
                                                             (bf16*)dq, (bf16*)dk, (bf16*)dv);
  else
    attn_q1_bwd_kernel<float><<<grid, Q1_WARPS * 32, 0, st>>>(a, (const float*)ctx, ldctx, (const float*)dctx, lddctx, lse,
                                                              (float*)dq, (float*)dk, (float*)dv);
  FCMF_LAUNCH_OK();
  return 0;
}

}  // namespace fcmf
