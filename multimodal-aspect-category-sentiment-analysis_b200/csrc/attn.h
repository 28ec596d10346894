// Internal attention interface shared by the CUDA-core engine (attention.cu) and the tcgen05 engine (attn_tc.cu).
#pragma once
#include "common.cuh"

namespace fcmf {

struct SegDev { const void* ptr; int64_t ld; int rows; int groups; const int32_t* idx; };
struct AttnDev {
  SegDev q[2], k[2], v[2];
  const float* mask_add; int64_t ld_mask; int mask_div;
  const float* bias;
  int NP, heads, dh, Lq, Lk;
  int causal;             // key j > query i => score := -1e4 (masked_fill semantics of the IAOG decoder, mm_modeling.py:115-124)
  float scale;
  int engine;             // engine of this call (0: process default)
  fcmf_dropout drop;      // dropout on the attention probabilities (mm_modeling.py:213, 260; roi_modeling.py:42-43); p == 0: off
};

// dropout row index of query i of (problem p, head h): the mask of an attention launch is keep(row, key j)
__host__ __device__ __forceinline__ uint64_t attn_drop_row(const AttnDev& a, int p, int h, int i) {
  return ((uint64_t)p * (uint64_t)a.heads + (uint64_t)h) * (uint64_t)a.Lq + (uint64_t)i;
}

template <typename T>
__device__ __forceinline__ const T* seg_row(const SegDev (&s)[2], int p, int r, int h, int dh) {
  const SegDev& g = (r < s[0].rows) ? s[0] : s[1];
  const int rl = (r < s[0].rows) ? r : r - s[0].rows;
  const int64_t grp = g.idx ? g.idx[p] : p;
  return reinterpret_cast<const T*>(g.ptr) + (grp * g.rows + rl) * g.ld + (int64_t)h * dh;
}


// tcgen05 engine (attn_tc.cu): bf16, head_dim 64, no per-pair bias
bool attn_tc_supported(const AttnDev& a, int64_t ldctx, const void* ctx);
int attn_tc_fwd(const AttnDev& a, void* ctx, int64_t ldctx, float* lse, cudaStream_t st);
int attn_tc_bwd(const AttnDev& a, const void* ctx, int64_t ldctx, const void* dctx, int64_t lddctx, const float* lse,
                float* delta, void* dq, void* dk, void* dv, cudaStream_t st);
// warp-specialised TMA-fed tcgen05 engine (attn_ws.cu): bf16, head_dim 64, padded key length <= 192
bool attn_ws_supported(const AttnDev& a, int64_t ldctx, const void* ctx);
int attn_ws_fwd(const AttnDev& a, void* ctx, int64_t ldctx, float* lse, cudaStream_t st);
bool attn_ws_bwd_supported(const AttnDev& a, int64_t ldctx, const void* ctx, int64_t lddctx);
int attn_ws_bwd(const AttnDev& a, const void* ctx, int64_t ldctx, const void* dctx, int64_t lddctx, const float* lse,
                void* dq, void* dk, void* dv, cudaStream_t st);
bool attn_ws_enabled();  // FCMF_ATTN_WS=0 in the environment keeps the cp.async kernels of attn_tc.cu (A/B measurements)
int attn_engine();      // 0 auto, 1 CUDA-core, 2 tcgen05 (fcmf_set_attn_engine)

// single-query kernels (attn_q1.cu): Lq == 1, no per-pair bias, either dtype
bool attn_q1_supported(const AttnDev& a);
int attn_q1_fwd(const AttnDev& a, void* ctx, int64_t ldctx, float* lse, int dtype, cudaStream_t st);
int attn_q1_bwd(const AttnDev& a, const void* ctx, int64_t ldctx, const void* dctx, int64_t lddctx, const float* lse,
                void* dq, void* dk, void* dv, int dtype, cudaStream_t st);

}  // namespace fcmf
