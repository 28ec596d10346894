// Forward kernel of the warp-specialised, TMA-fed tcgen05 attention (design notes: attn_ws.cuh).
#include "attn_ws.cuh"

namespace fcmf {
namespace ws {

struct FwdMaps { CUtensorMap q0f, q0t, q1, k0, k1, v0, v1, c0f, c0t, c1; };   // c*: the context output (TMA stores)

// Shared-memory plan. NKB = 3 has no room for separate P buffers: P block 0 then ALIASES the Q tile (dead once S is in
// TMEM) and the O staging slab aliases P block 1; the smaller shapes keep Q, P and the staging slab apart and ring deeper.
template <int NKB> struct FwdCfg {
  static constexpr bool kAlias = NKB == 3;
  static constexpr int QS = kAlias ? 3 : 4;                           // Q tile ring
  static constexpr int KS = NKB == 1 ? 3 : 2;                         // K/V item ring
  static constexpr int PB = kAlias ? NKB - 1 : NKB;                   // own P blocks per warpgroup
  static constexpr uint32_t kKV = 2 * NKB * BLK_B;                    // K then V of one item
  static constexpr uint32_t kQ0 = 0;
  static constexpr uint32_t kKV0 = QS * TILE_B;
  static constexpr uint32_t kPx0 = kKV0 + KS * kKV;                   // P blocks of the two warpgroups
  static constexpr uint32_t kStg0 = kPx0 + 2 * PB * TILE_B;           // O staging slabs (separate only when !kAlias)
  static constexpr uint32_t kMsk0 = kStg0 + (kAlias ? 0 : 2 * TILE_B);
  static constexpr uint32_t kBar0 = kMsk0 + KS * NKB * 64 * 4;
  static constexpr uint32_t kSmem = kBar0 + 256 + 1024;               // barriers + TMEM slot, alignment slack
};

// tile cursor over this CTA's (item, query tile) sequence
struct Cursor {
  int item, qt, it, t;           // item id, query tile, item ordinal of this CTA, tile ordinal of this CTA
};
__device__ __forceinline__ void advance(Cursor& c, int n_qt, int stride) {
  ++c.t;
  if (++c.qt == n_qt) { c.qt = 0; c.item += stride; ++c.it; }
}

// ------------------------------------------------------------------------------------------- forward
template <int NKB, bool DROP>
__global__ void __launch_bounds__(THREADS, 1)
attn_ws_fwd_kernel(const __grid_constant__ FwdMaps M, const Params P, bf16* __restrict__ ctx, int64_t ldctx, float* __restrict__ lse) {
  using Cfg = FwdCfg<NKB>;
  extern __shared__ uint8_t raw[];
  uint8_t* sm = raw + ((1024u - (s32(raw) & 1023u)) & 1023u);
  uint8_t* Qs = sm + Cfg::kQ0;
  uint8_t* KVs = sm + Cfg::kKV0;
  uint8_t* Px = sm + Cfg::kPx0;
  // P block b of warpgroup w for the tile in Q slot `slot`; the O staging slab of warpgroup w
  auto p_blk = [&](int w, int slot, int b) -> uint8_t* {
    if (Cfg::kAlias) return b == 0 ? Qs + slot * TILE_B : Px + (w * Cfg::PB + (b - 1)) * TILE_B;
    return Px + (w * Cfg::PB + b) * TILE_B;
  };
  auto stg_blk = [&](int w) -> uint8_t* { return Cfg::kAlias ? Px + (w * Cfg::PB) * TILE_B : sm + Cfg::kStg0 + w * TILE_B; };
  float* msk = reinterpret_cast<float*>(sm + Cfg::kMsk0);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + Cfg::kBar0);
  uint64_t* q_full = bars;                 // [4]
  uint64_t* q_empty = bars + 4;            // [4]
  uint64_t* kv_full = bars + 8;            // [3]
  uint64_t* kv_empty = bars + 11;          // [3]
  uint64_t* s_full = bars + 14;            // [2]
  uint64_t* p_full = bars + 16;            // [2]
  uint64_t* o_full = bars + 18;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_qt = P.n_qt, stride = gridDim.x;

  // K/V buffers hold stale rows between the loaded boxes and the end of the last key block: make them finite once
  for (uint32_t i = threadIdx.x * 16; i < Cfg::KS * Cfg::kKV; i += THREADS * 16) *reinterpret_cast<uint4*>(KVs + i) = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); }
    for (int i = 0; i < 3; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 128); mbar_init(&o_full[i], 1); }
    fence_init();
    prefetch_map(&M.q0f); prefetch_map(&M.q0t); prefetch_map(&M.q1); prefetch_map(&M.k0); prefetch_map(&M.k1);
    prefetch_map(&M.v0); prefetch_map(&M.v1); prefetch_map(&M.c0f); prefetch_map(&M.c0t); prefetch_map(&M.c1);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  fence_async();                           // the zero fill above (generic proxy) before any TMA write (async proxy)
  tc_before();
  __syncthreads();
  tc_after();
  const uint32_t tm = *tmem_slot;

  if (warp == 0) {
    // ===================================================================== producer
    Cursor c{(int)blockIdx.x, 0, 0, 0};
    while (c.item < P.items) {
      const int p = c.item / P.heads, h = c.item - p * P.heads;
      const int st = c.it % Cfg::KS;
      mbar_wait(&kv_empty[st], ((c.it / Cfg::KS) & 1) ^ 1, 1);
      {                                    // additive key mask of this item, log2 domain; -inf on padding positions
        const float* madd = P.mask_add ? P.mask_add + (int64_t)(p / P.mask_div) * P.ld_mask : nullptr;
        float* m = msk + st * NKB * 64;
        for (int j = lane; j < NKB * 64; j += 32) {
          const int lk = logical_row(P.kl, j);
          m[j] = lk >= 0 ? (madd ? madd[lk] * kLog2e : 0.f) : -INFINITY;
        }
      }
      __syncwarp();
      int gq0 = 0, gq1 = 0;
      if (lane == 0) {
        const int gk0 = P.kidx0 ? P.kidx0[p] : p, gv0 = P.vidx0 ? P.vidx0[p] : p;
        gq0 = P.qidx0 ? P.qidx0[p] : p;
        uint8_t* Ks = KVs + st * Cfg::kKV;
        uint8_t* Vs = Ks + NKB * BLK_B;
        uint32_t bytes = 2u * (uint32_t)P.kl.rows0p * 128u;
        if (P.kl.rows1) bytes += 2u * (uint32_t)P.kl.rows1p * 128u;
        mbar_expect_tx(&kv_full[st], bytes);
        tma3(Ks, &M.k0, &kv_full[st], h * 64, 0, gk0);
        tma3(Vs, &M.v0, &kv_full[st], h * 64, 0, gv0);
        if (P.kl.rows1) {
          const int gk1 = P.kidx1 ? P.kidx1[p] : p, gv1 = P.vidx1 ? P.vidx1[p] : p;
          tma3(Ks + P.kl.rows0p * 128, &M.k1, &kv_full[st], h * 64, 0, gk1);
          tma3(Vs + P.kl.rows0p * 128, &M.v1, &kv_full[st], h * 64, 0, gv1);
        }
        if (P.ql.rows1) gq1 = P.qidx1 ? P.qidx1[p] : p;
      }
      for (int qt = 0; qt < n_qt; ++qt) {
        const int slot = c.t % Cfg::QS;
        mbar_wait(&q_empty[slot], ((c.t / Cfg::QS) & 1) ^ 1, 2);
        if (lane == 0) {
          uint8_t* dst = Qs + slot * TILE_B;
          const int r0 = qt * 128;
          const int n0p = min(max(P.ql.rows0p - r0, 0), 128);
          const bool seg1_here = P.ql.rows1 > 0 && P.ql.rows0p >= r0 && P.ql.rows0p < r0 + 128;
          uint32_t bytes = (uint32_t)n0p * 128u + (seg1_here ? (uint32_t)P.ql.rows1p * 128u : 0u);
          mbar_expect_tx(&q_full[slot], bytes);
          if (n0p == 128) tma3(dst, &M.q0f, &q_full[slot], h * 64, r0, gq0);
          else if (n0p > 0) tma3(dst, &M.q0t, &q_full[slot], h * 64, r0, gq0);
          if (seg1_here) tma3(dst + (P.ql.rows0p - r0) * 128, &M.q1, &q_full[slot], h * 64, 0, gq1);
        }
        advance(c, n_qt, stride);
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (one thread)
    if (lane == 0) {
      const int ksteps_total = (P.kl.total + 15) >> 4;
      auto issue_s = [&](const Cursor& c) {
        const int slot = c.t % Cfg::QS, st = c.it % Cfg::KS, w = c.t & 1;
        mbar_wait(&q_full[slot], (c.t / Cfg::QS) & 1, 3);
        if (c.qt == 0) mbar_wait(&kv_full[st], (c.it / Cfg::KS) & 1, 4);
        tc_after();
        const uint32_t q = s32(Qs + slot * TILE_B), k = s32(KVs + st * Cfg::kKV);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma(tm + w * 256, sdesc(q + kk * 32, 16), sdesc(k + kk * 32, 16), idesc(NKB * 64, 0, 0), kk > 0 ? 1u : 0u);
        commit(&s_full[w]);
      };
      auto issue_pv = [&](const Cursor& c) {
        const int slot = c.t % Cfg::QS, st = c.it % Cfg::KS, w = c.t & 1;
        mbar_wait(&p_full[w], (c.t >> 1) & 1, 5);
        tc_after();
        const uint32_t v = s32(KVs + st * Cfg::kKV + NKB * BLK_B);
#pragma unroll
        for (int b = 0; b < NKB; ++b) {
          const uint32_t pb = s32(p_blk(w, slot, b));
          const int ks = min(4, ksteps_total - b * 4);
          for (int kk = 0; kk < ks; ++kk)
            umma(tm + w * 256 + 192, sdesc(pb + kk * 32, 16), sdesc(v + b * BLK_B + kk * 2048, 8192), idesc(64, 0, 1), (b > 0 || kk > 0) ? 1u : 0u);
        }
        commit(&o_full[w]);
        commit(&q_empty[slot]);
        if (c.qt == n_qt - 1) commit(&kv_empty[st]);
      };
      Cursor cs{(int)blockIdx.x, 0, 0, 0}, cp = cs;
      for (int i = 0; i < 2 && cs.item < P.items; ++i) { issue_s(cs); advance(cs, n_qt, stride); }
      while (cp.item < P.items) {
        issue_pv(cp);
        advance(cp, n_qt, stride);
        if (cs.item < P.items) { issue_s(cs); advance(cs, n_qt, stride); }
      }
    }
  } else {
    // ===================================================================== softmax warpgroups (tiles t = w, w+2, ...)
    const int w = (warp - 2) >> 2, quad = warp & 3;
    const int trow = quad * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const uint32_t tS = tm + w * 256 + lane_addr, tO = tS + 192;
    const float scale2 = P.scale * kLog2e;
    const int NC = (P.kl.total + 31) >> 5;                      // 32-column chunks that hold real keys
    DropCfg dc;
    if (DROP) dc = make_drop(P.drop);
    const int gap = P.kl.rows0p - P.kl.rows0;
    const bool st_thread = ((warp - 2) & 3) == 0 && lane == 0;  // issues this warpgroup's TMA stores
    Cursor c{(int)blockIdx.x, 0, 0, 0};
    if (w == 1) advance(c, n_qt, stride);
    while (c.item < P.items) {
      const int p = c.item / P.heads, h = c.item - p * P.heads;
      const int slot = c.t % Cfg::QS, st = c.it % Cfg::KS;
      const float* m = msk + st * NKB * 64;
      const int r0 = c.qt * 128;
      const int lrow = logical_row(P.ql, r0 + trow);
      // warp-uniform: does any of this warp's 32 tile rows hold a real query?
      const int wlo = r0 + quad * 32, whi = wlo + 32;
      const bool wact = wlo < P.ql.rows0 || (P.ql.rows1 > 0 && wlo < P.ql.rows0p + P.ql.rows1 && whi > P.ql.rows0p);
      mbar_wait(&kv_full[st], (c.it / Cfg::KS) & 1, 6);        // the mask vector of this item is visible
      mbar_wait(&s_full[w], (c.t >> 1) & 1, 7);
      tc_after();
      float mx = -INFINITY, sum = 0.f;
      float2 sum2 = f2(0.f, 0.f);
      if (wact) {
#pragma unroll 1
        for (int cc = 0; cc < NC; ++cc) {
          uint32_t r[32];
          tmem_ld32(tS + cc * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 m4 = *reinterpret_cast<const float4*>(m + cc * 32 + j);
            const float2 a = __ffma2_rn(u2f2(r[j], r[j + 1]), f2(scale2, scale2), f2(m4.x, m4.y));
            const float2 b = __ffma2_rn(u2f2(r[j + 2], r[j + 3]), f2(scale2, scale2), f2(m4.z, m4.w));
            mx = fmaxf(mx, fmaxf(fmaxf(a.x, a.y), fmaxf(b.x, b.y)));
          }
        }
      }
      // the staging slab (= P block 1 when aliased) of this warpgroup's previous tile has been read by its TMA store
      if (st_thread) tma_store_wait_read();
      wg_bar(w);
      if (wact) {
        uint32_t rseed = 0;
        if (DROP) rseed = drop_rowseed(dc.seed, ((uint64_t)p * (uint64_t)P.heads + (uint64_t)h) * (uint64_t)P.Lq + (uint64_t)max(lrow, 0));
#pragma unroll 1
        for (int cc = 0; cc < NC; ++cc) {
          uint32_t r[32];
          float v[32];
          tmem_ld32(tS + cc * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 m4 = *reinterpret_cast<const float4*>(m + cc * 32 + j);
            const float2 a = __fadd2_rn(__ffma2_rn(u2f2(r[j], r[j + 1]), f2(scale2, scale2), f2(m4.x, m4.y)), f2(-mx, -mx));
            const float2 b = __fadd2_rn(__ffma2_rn(u2f2(r[j + 2], r[j + 3]), f2(scale2, scale2), f2(m4.z, m4.w)), f2(-mx, -mx));
            v[j] = ex2_approx(a.x); v[j + 1] = ex2_approx(a.y); v[j + 2] = ex2_approx(b.x); v[j + 3] = ex2_approx(b.y);
            sum2 = __fadd2_rn(sum2, __fadd2_rn(f2(v[j], v[j + 1]), f2(v[j + 2], v[j + 3])));
          }
          if (DROP) {                                            // the denominator keeps the dropped terms
            if ((cc + 1) * 32 <= P.kl.rows0p) {                  // chunk inside segment 0: padded position == key index
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                const uint32_t hsh = drop_pair(rseed, (uint32_t)(cc * 32 + j));
                v[j] = drop_keep_lo(hsh, dc.thr16) ? v[j] : 0.f;
                v[j + 1] = drop_keep_hi(hsh, dc.thr16) ? v[j + 1] : 0.f;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int x = cc * 32 + j;
                const int lk = x < P.kl.rows0p ? x : x - gap;    // padding positions hold exact zeros already
                v[j] = drop_keep(rseed, (uint32_t)lk, dc.thr16) ? v[j] : 0.f;
              }
            }
          }
          store_row32(p_blk(w, slot, cc >> 1), trow, (cc & 1) * 32, v);
        }
        sum = sum2.x + sum2.y;
      }
      fence_async();                                             // P (generic proxy) -> the MMA (async proxy)
      tc_before();
      mbar_arrive(&p_full[w]);
      mbar_wait(&o_full[w], (c.t >> 1) & 1, 8);
      tc_after();
      // O rows -> bf16 staging slab (a thread owns a row: direct global stores would send 16-byte pieces to 32 different
      // lines per instruction -- measured: the LSU wavefronts of those stores bounded the round-1 kernels) -> TMA store
      uint8_t* stg = stg_blk(w);
      if (wact) {
        uint32_t r0v[32], r1v[32];
        tmem_ld32(tO, r0v);
        tmem_ld32(tO + 32, r1v);
        tmem_ld_wait();
        const float osc = (DROP ? dc.inv_keep : 1.0f) / sum;
        stage_out32(stg, trow, 0, r0v, osc);
        stage_out32(stg, trow, 32, r1v, osc);
        if (lrow >= 0 && lse) lse[((int64_t)p * P.heads + h) * P.Lq + lrow] = (mx + __log2f(sum)) * kLn2;
      }
      tc_before();                                               // this tile's tcgen05.ld before the MMAs of tile t+2
      fence_async();
      wg_bar(w);
      if (st_thread) {
        const int n0p = min(max(P.ql.rows0p - r0, 0), 128);
        const bool seg1_here = P.ql.rows1 > 0 && P.ql.rows0p >= r0 && P.ql.rows0p < r0 + 128;
        if (n0p == 128) tma3_store(&M.c0f, stg, h * 64, r0, p);
        else if (n0p > 0) tma3_store(&M.c0t, stg, h * 64, r0, p);
        if (seg1_here) tma3_store(&M.c1, stg + (P.ql.rows0p - r0) * 128, h * 64, 0, p);
        tma_store_commit();
      }
      advance(c, n_qt, stride);
      if (c.item < P.items) advance(c, n_qt, stride);
    }
    if (st_thread) tma_store_wait_all();                         // shared memory must outlive the bulk stores
  }
  tc_before();
  __syncthreads();
  if (warp == 1) { tc_after(); tmem_dealloc(tm, 512); }
}

}  // namespace ws

bool attn_ws_supported(const AttnDev& a, int64_t ldctx, const void* ctx) {
  using namespace ws;
  if (a.dh != 64 || a.bias != nullptr || a.causal) return false;
  if ((ldctx % 8) || (reinterpret_cast<uintptr_t>(ctx) & 15u)) return false;
  for (int s = 0; s < 2; ++s)
    if (!seg_ok(a.q[s]) || !seg_ok(a.k[s]) || !seg_ok(a.v[s])) return false;
  const RowLay ql = row_lay(a.q), kl = row_lay(a.k);
  if (a.k[0].rows != a.v[0].rows || a.k[1].rows != a.v[1].rows) return false;
  if (key_blocks(kl) > 3 || kl.rows0p > 256) return false;
  if (a.Lq < 16 || !tiles_ok(ql, 128)) return false;
  return (int64_t)a.NP * a.heads < (1LL << 30);
}

int attn_ws_fwd(const AttnDev& a, void* ctx, int64_t ldctx, float* lse, cudaStream_t st) {
  using namespace ws;
  Params P;
  fill_params(a, &P);
  FwdMaps M;
  if (int r = role_maps(a.q, P.ql, a.heads, 128, &M.q0f, &M.q0t, &M.q1)) return r;
  // keys / values: ONE box of rows0p rows per segment-0 load (<= 256)
  if (int r = make_map3(&M.k0, a.k[0].ptr, (int64_t)a.heads * 64, a.k[0].rows, a.k[0].groups, a.k[0].ld, (int64_t)a.k[0].rows * a.k[0].ld, P.kl.rows0p)) return r;
  if (int r = make_map3(&M.v0, a.v[0].ptr, (int64_t)a.heads * 64, a.v[0].rows, a.v[0].groups, a.v[0].ld, (int64_t)a.v[0].rows * a.v[0].ld, P.kl.rows0p)) return r;
  if (P.kl.rows1) {
    if (int r = make_map3(&M.k1, a.k[1].ptr, (int64_t)a.heads * 64, a.k[1].rows, a.k[1].groups, a.k[1].ld, (int64_t)a.k[1].rows * a.k[1].ld, P.kl.rows1p)) return r;
    if (int r = make_map3(&M.v1, a.v[1].ptr, (int64_t)a.heads * 64, a.v[1].rows, a.v[1].groups, a.v[1].ld, (int64_t)a.v[1].rows * a.v[1].ld, P.kl.rows1p)) return r;
  } else { M.k1 = M.k0; M.v1 = M.v0; }
  {   // context output [NP][Lq][heads*64] as two virtual row segments (rows [0, rows0) and [rows0, Lq) of every problem)
    const int64_t cols = (int64_t)a.heads * 64, gs = (int64_t)a.Lq * ldctx;
    const int tail = P.ql.rows0p % 128;
    if (int r = make_map3(&M.c0f, ctx, cols, P.ql.rows0, a.NP, ldctx, gs, 128)) return r;
    if (int r = make_map3(&M.c0t, ctx, cols, P.ql.rows0, a.NP, ldctx, gs, tail ? tail : 8)) return r;
    if (P.ql.rows1) { if (int r = make_map3(&M.c1, (const bf16*)ctx + (int64_t)P.ql.rows0 * ldctx, cols, P.ql.rows1, a.NP, ldctx, gs, P.ql.rows1p)) return r; }
    else M.c1 = M.c0f;
  }
  const int nkb = key_blocks(P.kl);
  const bool drop = a.drop.p > 0.f;
  const unsigned grid = (unsigned)std::min<int64_t>(P.items, sm_count());
#define WS_LAUNCH(NB, DR)                                                                                       \
  {                                                                                                             \
    auto kern = attn_ws_fwd_kernel<NB, DR>;                                                                     \
    FCMF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FwdCfg<NB>::kSmem)); \
    kern<<<grid, THREADS, FwdCfg<NB>::kSmem, st>>>(M, P, (bf16*)ctx, ldctx, lse);                               \
  }
  switch (nkb) {
    case 1: if (drop) WS_LAUNCH(1, true) else WS_LAUNCH(1, false) break;
    case 2: if (drop) WS_LAUNCH(2, true) else WS_LAUNCH(2, false) break;
    default: if (drop) WS_LAUNCH(3, true) else WS_LAUNCH(3, false) break;
  }
#undef WS_LAUNCH
  FCMF_LAUNCH_OK();
  return 0;
}

}  // namespace fcmf
