// Generic folded multi-head attention, CUDA-core engine (fp32 math, bf16 or fp32 storage).
//
// One CTA = one (problem, head): the "other side" matrix (K,V for forward/dQ; Q,dO for dK/dV) is staged once
// in shared memory with a padded row stride (conflict-free for both the lane<->row dot products and the
// lane<->column accumulations); each warp then owns whole output rows, so no atomics are needed.
// Backward follows the flash formulation: probabilities are recomputed from the saved log-sum-exp.
#include "common.cuh"
#include "attn.h"
#include <atomic>
#include <stdlib.h>

namespace fcmf {

constexpr int AT_WARPS = 8;
constexpr int AT_THREADS = AT_WARPS * 32;
constexpr int AT_MAX_DPL = 4;                 // head dim <= 128

// Stage `rows` x dh of a two-segment matrix into smem (fp32, stride dh+1).
template <typename T>
__device__ __forceinline__ void stage_matrix(float* dst, const SegDev (&s)[2], int p, int h, int rows, int dh) {
  const int ld = dh + 1;
  for (int e = threadIdx.x; e < rows * dh; e += blockDim.x) {
    const int r = e / dh, d = e - r * dh;
    dst[r * ld + d] = to_f(seg_row<T>(s, p, r, h, dh)[d]);
  }
}

// ------------------------------------------------------------------------------------------------ forward
template <typename T>
__global__ void __launch_bounds__(AT_THREADS)
attn_fwd_kernel(AttnDev a, T* __restrict__ ctx, int64_t ldctx, float* __restrict__ lse) {
  extern __shared__ float sm[];
  const int p = blockIdx.x / a.heads, h = blockIdx.x % a.heads;
  const int dh = a.dh, ld = dh + 1, Lk = a.Lk, Lq = a.Lq;
  float* Ks = sm;
  float* Vs = Ks + Lk * ld;
  float* qs = Vs + Lk * ld;                    // [AT_WARPS][dh]
  float* ps = qs + AT_WARPS * dh;              // [AT_WARPS][Lk]
  stage_matrix<T>(Ks, a.k, p, h, Lk, dh);
  stage_matrix<T>(Vs, a.v, p, h, Lk, dh);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* q = qs + warp * dh;
  float* pr = ps + warp * Lk;
  const float* madd = a.mask_add ? a.mask_add + (int64_t)(p / a.mask_div) * a.ld_mask : nullptr;
  const DropCfg dc = make_drop(a.drop);
  for (int i = warp; i < Lq; i += AT_WARPS) {
    const T* qrow = seg_row<T>(a.q, p, i, h, dh);
    for (int d = lane; d < dh; d += 32) q[d] = to_f(qrow[d]);
    __syncwarp();
    const uint32_t rseed = dc.thr16 ? drop_rowseed(dc.seed, attn_drop_row(a, p, h, i)) : 0u;
    const float* brow = a.bias ? a.bias + (((int64_t)p * a.heads + h) * Lq + i) * Lk : nullptr;
    float mx = -INFINITY;
    for (int j = lane; j < Lk; j += 32) {
      const float* kr = Ks + j * ld;
      float s = 0.f;
      for (int d = 0; d < dh; ++d) s = fmaf(q[d], kr[d], s);
      s *= a.scale;                                              // scale BEFORE the mask add (mm_modeling.py:204-206)
      if (madd) s += madd[j];
      if (brow) s += brow[j];
      if (a.causal && j > i) s = -1e4f;
      pr[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < Lk; j += 32) {
      const float e = __expf(pr[j] - mx);
      sum += e;                                                  // the softmax denominator is taken BEFORE dropout
      pr[j] = (dc.thr16 && !drop_keep(rseed, (uint32_t)j, dc.thr16)) ? 0.f : e;
    }
    sum = warp_sum(sum);
    const float inv = (dc.thr16 ? dc.inv_keep : 1.0f) / sum;
    __syncwarp();
    T* orow = ctx + ((int64_t)p * Lq + i) * ldctx + (int64_t)h * dh;
    for (int d = lane; d < dh; d += 32) {
      float o = 0.f;
      for (int j = 0; j < Lk; ++j) o = fmaf(pr[j], Vs[j * ld + d], o);
      orow[d] = from_f<T>(o * inv);
    }
    if (lane == 0 && lse) lse[((int64_t)p * a.heads + h) * Lq + i] = mx + __logf(sum);
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------ dQ (+delta, +dbias)
template <typename T>
__global__ void __launch_bounds__(AT_THREADS)
attn_bwd_dq_kernel(AttnDev a, const T* __restrict__ ctx, int64_t ldctx, const T* __restrict__ dctx, int64_t lddctx,
                   const float* __restrict__ lse, T* __restrict__ dq, float* __restrict__ delta,
                   float* __restrict__ dbias) {
  extern __shared__ float sm[];
  const int p = blockIdx.x / a.heads, h = blockIdx.x % a.heads;
  const int dh = a.dh, ld = dh + 1, Lk = a.Lk, Lq = a.Lq, HD = a.heads * dh;
  float* Ks = sm;
  float* Vs = Ks + Lk * ld;
  float* qs = Vs + Lk * ld;                    // [AT_WARPS][2*dh]  (q, dO)
  float* ps = qs + AT_WARPS * 2 * dh;          // [AT_WARPS][Lk]    (dS)
  stage_matrix<T>(Ks, a.k, p, h, Lk, dh);
  stage_matrix<T>(Vs, a.v, p, h, Lk, dh);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* q = qs + warp * 2 * dh;
  float* go = q + dh;
  float* dS = ps + warp * Lk;
  const float* madd = a.mask_add ? a.mask_add + (int64_t)(p / a.mask_div) * a.ld_mask : nullptr;
  const DropCfg dc = make_drop(a.drop);
  for (int i = warp; i < Lq; i += AT_WARPS) {
    const uint32_t rseed = dc.thr16 ? drop_rowseed(dc.seed, attn_drop_row(a, p, h, i)) : 0u;
    const T* qrow = seg_row<T>(a.q, p, i, h, dh);
    const T* orow = ctx + ((int64_t)p * Lq + i) * ldctx + (int64_t)h * dh;
    const T* grow = dctx + ((int64_t)p * Lq + i) * lddctx + (int64_t)h * dh;
    float dl = 0.f;
    for (int d = lane; d < dh; d += 32) {
      q[d] = to_f(qrow[d]);
      const float g = to_f(grow[d]);
      go[d] = g;
      dl = fmaf(g, to_f(orow[d]), dl);
    }
    dl = warp_sum(dl);                                           // delta_i = dO_i . O_i = sum_j P_ij dP_ij
    __syncwarp();
    const int64_t stat = ((int64_t)p * a.heads + h) * Lq + i;
    const float l = lse[stat];
    if (lane == 0) delta[stat] = dl;
    const float* brow = a.bias ? a.bias + stat * Lk : nullptr;
    for (int j = lane; j < Lk; j += 32) {
      const float* kr = Ks + j * ld;
      const float* vr = Vs + j * ld;
      float s = 0.f, dp = 0.f;
      for (int d = 0; d < dh; ++d) { s = fmaf(q[d], kr[d], s); dp = fmaf(go[d], vr[d], dp); }
      s *= a.scale;
      if (madd) s += madd[j];
      if (brow) s += brow[j];
      const bool dead = a.causal && j > i;
      if (dead) s = -1e4f;
      const float pj = __expf(s - l);
      if (dc.thr16) dp = drop_keep(rseed, (uint32_t)j, dc.thr16) ? dp * dc.inv_keep : 0.f;   // dP = mask/(1-p) * (dO . v_j)
      const float ds = dead ? 0.f : pj * (dp - dl);
      dS[j] = ds;
      if (dbias) dbias[stat * Lk + j] = ds;
    }
    __syncwarp();
    T* dqrow = dq + ((int64_t)p * Lq + i) * HD + (int64_t)h * dh;
    for (int d = lane; d < dh; d += 32) {
      float o = 0.f;
      for (int j = 0; j < Lk; ++j) o = fmaf(dS[j], Ks[j * ld + d], o);
      dqrow[d] = from_f<T>(o * a.scale);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------ dK, dV
template <typename T>
__global__ void __launch_bounds__(AT_THREADS)
attn_bwd_dkv_kernel(AttnDev a, const T* __restrict__ dctx, int64_t lddctx, const float* __restrict__ lse,
                    const float* __restrict__ delta, T* __restrict__ dk, T* __restrict__ dv) {
  extern __shared__ float sm[];
  const int p = blockIdx.x / a.heads, h = blockIdx.x % a.heads;
  const int dh = a.dh, ld = dh + 1, Lk = a.Lk, Lq = a.Lq, HD = a.heads * dh;
  float* Qs = sm;                              // [Lq][dh+1]
  float* Gs = Qs + Lq * ld;                    // [Lq][dh+1]  dO
  float* ls = Gs + Lq * ld;                    // [Lq] lse
  float* dl = ls + Lq;                         // [Lq] delta
  float* kv = dl + Lq;                         // [AT_WARPS][2*dh]  (k_j, v_j)
  float* ps = kv + AT_WARPS * 2 * dh;          // [AT_WARPS][2*Lq]  (P, dS)
  stage_matrix<T>(Qs, a.q, p, h, Lq, dh);
  for (int e = threadIdx.x; e < Lq * dh; e += blockDim.x) {
    const int r = e / dh, d = e - r * dh;
    Gs[r * ld + d] = to_f(dctx[((int64_t)p * Lq + r) * lddctx + (int64_t)h * dh + d]);
  }
  const int64_t stat0 = ((int64_t)p * a.heads + h) * Lq;
  for (int i = threadIdx.x; i < Lq; i += blockDim.x) { ls[i] = lse[stat0 + i]; dl[i] = delta[stat0 + i]; }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* kj = kv + warp * 2 * dh;
  float* vj = kj + dh;
  float* P = ps + warp * 2 * Lq;
  float* dS = P + Lq;
  const float* madd = a.mask_add ? a.mask_add + (int64_t)(p / a.mask_div) * a.ld_mask : nullptr;
  const DropCfg dc = make_drop(a.drop);
  for (int j = warp; j < Lk; j += AT_WARPS) {
    const T* krow = seg_row<T>(a.k, p, j, h, dh);
    const T* vrow = seg_row<T>(a.v, p, j, h, dh);
    for (int d = lane; d < dh; d += 32) { kj[d] = to_f(krow[d]); vj[d] = to_f(vrow[d]); }
    __syncwarp();
    const float mj = madd ? madd[j] : 0.f;
    for (int i = lane; i < Lq; i += 32) {
      const float* qr = Qs + i * ld;
      const float* gr = Gs + i * ld;
      float s = 0.f, dp = 0.f;
      for (int d = 0; d < dh; ++d) { s = fmaf(qr[d], kj[d], s); dp = fmaf(gr[d], vj[d], dp); }
      s = s * a.scale + mj;
      if (a.bias) s += a.bias[(stat0 + i) * Lk + j];
      const bool dead = a.causal && j > i;
      if (dead) s = -1e4f;
      const float pj = __expf(s - ls[i]);
      float pd = pj;                                             // dropped probability (feeds dV)
      if (dc.thr16) {
        const bool keep = drop_keep(drop_rowseed(dc.seed, attn_drop_row(a, p, h, i)), (uint32_t)j, dc.thr16);
        pd = keep ? pj * dc.inv_keep : 0.f;
        dp = keep ? dp * dc.inv_keep : 0.f;
      }
      P[i] = pd;
      dS[i] = dead ? 0.f : pj * (dp - dl[i]);
    }
    __syncwarp();
    T* dkrow = dk + ((int64_t)p * Lk + j) * HD + (int64_t)h * dh;
    T* dvrow = dv + ((int64_t)p * Lk + j) * HD + (int64_t)h * dh;
    for (int d = lane; d < dh; d += 32) {
      float ak = 0.f, av = 0.f;
      for (int i = 0; i < Lq; ++i) { ak = fmaf(dS[i], Qs[i * ld + d], ak); av = fmaf(P[i], Gs[i * ld + d], av); }
      dkrow[d] = from_f<T>(ak * a.scale);
      dvrow[d] = from_f<T>(av);
    }
    __syncwarp();
  }
}

static std::atomic<int> g_attn_engine{0};
int attn_engine() { return g_attn_engine.load(std::memory_order_relaxed); }
bool attn_ws_enabled() {
  static const bool on = [] { const char* e = getenv("FCMF_ATTN_WS"); return !(e && e[0] == '0'); }();
  return on;
}

static int to_dev(const fcmf_attn_desc* d, AttnDev* o) {
  FCMF_CHECK_ARG(d != nullptr, "attn: null descriptor");
  for (int s = 0; s < 2; ++s) {
    o->q[s] = {d->q[s].ptr, d->q[s].ld, d->q[s].ptr ? d->q[s].rows : 0, d->q[s].groups, d->q[s].idx};
    o->k[s] = {d->k[s].ptr, d->k[s].ld, d->k[s].ptr ? d->k[s].rows : 0, d->k[s].groups, d->k[s].idx};
    o->v[s] = {d->v[s].ptr, d->v[s].ld, d->v[s].ptr ? d->v[s].rows : 0, d->v[s].groups, d->v[s].idx};
  }
  FCMF_CHECK_ARG(d->q[0].ptr && d->k[0].ptr && d->v[0].ptr, "attn: segment 0 of q/k/v is required");
  FCMF_CHECK_ARG(o->k[0].rows == o->v[0].rows && o->k[1].rows == o->v[1].rows, "attn: k/v segment rows differ");
  o->mask_add = d->mask_add; o->ld_mask = d->ld_mask; o->mask_div = d->mask_div > 0 ? d->mask_div : 1;
  o->bias = d->bias;
  o->NP = d->NP; o->heads = d->heads; o->dh = d->dh; o->scale = d->scale; o->causal = d->causal ? 1 : 0;
  o->drop = d->drop;
  o->engine = (d->engine >= FCMF_ENGINE_SIMT && d->engine <= FCMF_ENGINE_TCGEN05) ? d->engine : 0;
  FCMF_CHECK_ARG(drop_check(&d->drop) == 0, "attn: dropout p must be in [0, 1)");
  o->Lq = o->q[0].rows + o->q[1].rows;
  o->Lk = o->k[0].rows + o->k[1].rows;
  FCMF_CHECK_ARG(o->NP >= 0 && o->heads > 0 && o->dh > 0 && o->dh <= 32 * AT_MAX_DPL && o->Lq > 0 && o->Lk > 0,
                 "attn: bad shape NP=%d heads=%d dh=%d Lq=%d Lk=%d", o->NP, o->heads, o->dh, o->Lq, o->Lk);
  FCMF_CHECK_ARG((int64_t)o->NP * o->heads < (1LL << 31), "attn: too many problems");
  return 0;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  FCMF_CHECK_ARG(bytes <= 227 * 1024, "attn: %zu bytes of shared memory needed (> 227 KB): sequence too long for the CUDA-core engine", bytes);
  if (bytes > 48 * 1024) FCMF_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

}  // namespace fcmf

using namespace fcmf;

extern "C" int fcmf_set_attn_engine(int engine) {
  FCMF_CHECK_ARG(engine >= FCMF_ENGINE_AUTO && engine <= FCMF_ENGINE_TCGEN05, "set_attn_engine: bad engine %d", engine);
  g_attn_engine.store(engine, std::memory_order_relaxed);
  return 0;
}

extern "C" int fcmf_attn_fwd(const fcmf_attn_desc* d, void* ctx, int64_t ldctx, float* lse, int dtype, void* stream) {
  AttnDev a;
  if (int r = to_dev(d, &a)) return r;
  FCMF_CHECK_ARG(dtype == FCMF_F32 || dtype == FCMF_BF16, "attn_fwd: bad dtype %d", dtype);
  if (a.NP == 0) return 0;
  {
    const int eng = a.engine ? a.engine : attn_engine();
    const bool ok = dtype == FCMF_BF16 && attn_tc_supported(a, ldctx, ctx);
    if (eng == FCMF_ENGINE_TCGEN05 && !ok) return fail(FCMF_ERR_UNSUPPORTED, "attn_fwd: tcgen05 engine needs bf16, head_dim 64, no bias, 16 <= L <= 320");
    if (ok && eng != FCMF_ENGINE_SIMT && attn_ws_enabled() && attn_ws_supported(a, ldctx, ctx)) return attn_ws_fwd(a, ctx, ldctx, lse, as_stream(stream));
    if (ok && eng != FCMF_ENGINE_SIMT) return attn_tc_fwd(a, ctx, ldctx, lse, as_stream(stream));
    if (eng != FCMF_ENGINE_SIMT && attn_q1_supported(a)) return attn_q1_fwd(a, ctx, ldctx, lse, dtype, as_stream(stream));
  }
  const size_t smem = sizeof(float) * ((size_t)2 * a.Lk * (a.dh + 1) + (size_t)AT_WARPS * (a.dh + a.Lk));
  const unsigned grid = (unsigned)(a.NP * a.heads);
  cudaStream_t st = as_stream(stream);
  if (dtype == FCMF_BF16) {
    if (int r = set_smem(attn_fwd_kernel<bf16>, smem)) return r;
    attn_fwd_kernel<bf16><<<grid, AT_THREADS, smem, st>>>(a, (bf16*)ctx, ldctx, lse);
  } else {
    if (int r = set_smem(attn_fwd_kernel<float>, smem)) return r;
    attn_fwd_kernel<float><<<grid, AT_THREADS, smem, st>>>(a, (float*)ctx, ldctx, lse);
  }
  FCMF_LAUNCH_OK();
  return 0;
}

extern "C" int fcmf_attn_bwd(const fcmf_attn_desc* d, const void* ctx, int64_t ldctx, const void* dctx, int64_t lddctx,
                             const float* lse, float* delta, void* dq, void* dk, void* dv, float* dbias, int dtype,
                             void* stream) {
  AttnDev a;
  if (int r = to_dev(d, &a)) return r;
  FCMF_CHECK_ARG(dtype == FCMF_F32 || dtype == FCMF_BF16, "attn_bwd: bad dtype %d", dtype);
  FCMF_CHECK_ARG(lse && delta && dq && dk && dv, "attn_bwd: null buffer");
  if (a.NP == 0) return 0;
  cudaStream_t st = as_stream(stream);
  {
    const int eng = a.engine ? a.engine : attn_engine();
    const bool ok = dtype == FCMF_BF16 && dbias == nullptr && attn_tc_supported(a, ldctx, ctx) && (lddctx % 8) == 0;
    if (eng == FCMF_ENGINE_TCGEN05 && !ok) return fail(FCMF_ERR_UNSUPPORTED, "attn_bwd: tcgen05 engine needs bf16, head_dim 64, no bias, 16 <= L <= 320");
    if (ok && eng != FCMF_ENGINE_SIMT && attn_ws_enabled() && attn_ws_bwd_supported(a, ldctx, ctx, lddctx) &&
        (reinterpret_cast<uintptr_t>(dctx) & 15u) == 0)
      return attn_ws_bwd(a, ctx, ldctx, dctx, lddctx, lse, dq, dk, dv, st);
    if (ok && eng != FCMF_ENGINE_SIMT) return attn_tc_bwd(a, ctx, ldctx, dctx, lddctx, lse, delta, dq, dk, dv, st);
    if (eng != FCMF_ENGINE_SIMT && dbias == nullptr && attn_q1_supported(a)) return attn_q1_bwd(a, ctx, ldctx, dctx, lddctx, lse, dq, dk, dv, dtype, st);
  }
  const size_t smem_q = sizeof(float) * ((size_t)2 * a.Lk * (a.dh + 1) + (size_t)AT_WARPS * (2 * a.dh + a.Lk));
  const size_t smem_kv = sizeof(float) * ((size_t)2 * a.Lq * (a.dh + 1) + 2 * (size_t)a.Lq + (size_t)AT_WARPS * (2 * a.dh + 2 * a.Lq));
  const unsigned grid = (unsigned)(a.NP * a.heads);
  int rc = 0;
  if (dtype == FCMF_BF16) {
    if ((rc = set_smem(attn_bwd_dq_kernel<bf16>, smem_q)) == 0 && (rc = set_smem(attn_bwd_dkv_kernel<bf16>, smem_kv)) == 0) {
      attn_bwd_dq_kernel<bf16><<<grid, AT_THREADS, smem_q, st>>>(a, (const bf16*)ctx, ldctx, (const bf16*)dctx, lddctx, lse, (bf16*)dq, delta, dbias);
      attn_bwd_dkv_kernel<bf16><<<grid, AT_THREADS, smem_kv, st>>>(a, (const bf16*)dctx, lddctx, lse, delta, (bf16*)dk, (bf16*)dv);
    }
  } else {
    if ((rc = set_smem(attn_bwd_dq_kernel<float>, smem_q)) == 0 && (rc = set_smem(attn_bwd_dkv_kernel<float>, smem_kv)) == 0) {
      attn_bwd_dq_kernel<float><<<grid, AT_THREADS, smem_q, st>>>(a, (const float*)ctx, ldctx, (const float*)dctx, lddctx, lse, (float*)dq, delta, dbias);
      attn_bwd_dkv_kernel<float><<<grid, AT_THREADS, smem_kv, st>>>(a, (const float*)dctx, lddctx, lse, delta, (float*)dk, (float*)dv);
    }
  }
  if (rc) return rc;
  FCMF_LAUNCH_OK();
  return 0;
}
