// ONE backward kernel of the warp-specialised tcgen05 attention: S and dP are computed once per (query tile, key block)
// and feed dQ, dK and dV together (the round-1 design ran a dQ kernel and a dK/dV kernel that each recomputed S, dP and
// the exponentials). Design notes common to forward and backward: attn_ws.cuh.
//
// Work item = (problem, head). Iterations of an item: key block g (64 keys, outer) x query tile qt (128 rows, inner).
//   A(g, qt):  S  = Q_qt . K_g^T          dP = dO_qt . V_g^T                          -> TMEM (two buffers, one per team)
//   team:      P  = exp2(S*scale*log2e + mask - lse*log2e),  dP' = dropout(dP),  dS = P o (dP' - delta)
//              P' (= dropout(P)) and dS as bf16 images [128 q rows x 64 keys] in shared memory
//   B(g, qt):  dQ_qt += dS . K_g          (A = dS image, K-major;  B = K_g as MN-major operand)
//              dV_g  += P'^T . dO_qt      (A = P' image read MN-major: M = keys;  B = dO tile as MN-major operand)
//              dK_g  += dS^T . Q_qt       (A = dS image read MN-major;            B = Q tile as MN-major operand)
// dV_g / dK_g leave TMEM after the last query tile of block g, dQ_qt after the last key block; both go through a bf16
// staging slab and a TMA store (thread-per-row global stores were the bound of the round-1 kernels: 32 lines per request).
// The MN-major A operand is 128 "rows" (keys) wide while a key block has 64: the upper half reads the neighbouring image
// and produces 64 accumulator lanes nobody reads.
//
// Roles (640 threads): warp 0 TMA producer (+ mask slices), warp 1 MMA issuer (one thread, polls A/B readiness so that a
// late load never blocks a ready B step), warps 2-3 delta = rowsum(dO o O), warps 4-19 two teams of 8 warps that take
// alternate iterations (thread = one query row x 32 of the 64 key columns).
#include "attn_ws.cuh"

namespace fcmf {
namespace ws {

constexpr int BW_THREADS = 640;
constexpr int KVS = 3;                          // (K | V) block ring

struct BwdMaps {
  CUtensorMap q0f, q0t, q1;                     // loads: Q tiles (128 rows)
  CUtensorMap g0f, g0t, g1;                     //        dO tiles (two virtual row segments of [NP][Lq][HD])
  CUtensorMap o0f, o0t, o1;                     //        O tiles (only when delta comes from dO . O: more than one key block)
  CUtensorMap k0f, k0t, k1;                     //        K blocks (64 rows)
  CUtensorMap v0f, v0t, v1;
  CUtensorMap dq0f, dq0t, dq1;                  // stores
  CUtensorMap dk0f, dk0t, dk1;
  CUtensorMap dv0f, dv0t, dv1;
};

// delta_i = sum_j dO_ij O_ij = sum_j P_ij dP'_ij. With ONE key block (text->image: 49 keys) a team holds the whole row of P
// and dP' and forms delta itself (one exchange between the two threads of a row): no O traffic at all, ring of three
// (Q | dO) slots. With more blocks the producer also loads the O tile and two dedicated warps form dO . O from shared
// memory (and stage lse): two (Q | dO | O) slots -- the same 96 KB.
template <bool ONEBLK> struct BwdSmem {
  static constexpr int QGS = ONEBLK ? 3 : 2;
  static constexpr uint32_t kSlot = (ONEBLK ? 2 : 3) * TILE_B;
  static constexpr uint32_t kPdS0 = 0;                                  // [team][P 16K | dS 16K]
  static constexpr uint32_t kQG0 = 4 * TILE_B;                          // [slot][Q 16K | dO 16K (| O 16K)]
  static constexpr uint32_t kKV0 = kQG0 + QGS * kSlot;                  // [slot][K 8K | V 8K]
  static constexpr uint32_t kMsk0 = kKV0 + KVS * 2 * BLK_B;             // [slot][64] f32 additive key mask, log2 domain
  static constexpr uint32_t kDlt0 = kMsk0 + KVS * 64 * 4;               // [slot][delta 128 | lse*log2e 128] f32; ONEBLK: [team][half][128] partial sums
  static constexpr uint32_t kBar0 = kDlt0 + 3 * 256 * 4;
  static constexpr uint32_t kSmem = kBar0 + 512 + 1024;
};

// iteration cursor of one CTA: item (outer), key block g, query tile qt (inner)
struct It {
  int item, it, g, qt, n;
};
__device__ __forceinline__ void next(It& c, int nkb, int n_qt, int stride) {
  ++c.n;
  if (++c.qt == n_qt) {
    c.qt = 0;
    if (++c.g == nkb) { c.g = 0; c.item += stride; ++c.it; }
  }
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {     // non-blocking
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(s32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void team_bar(int e) { asm volatile("bar.sync %0, 256;" ::"r"(e + 1) : "memory"); }
// MN-major A operand: 64-element M blocks are `lbo` bytes apart, 8-row (reduction) groups 1024 B apart
__device__ __forceinline__ uint64_t sdesc_mn(uint32_t saddr, uint32_t lbo_bytes) { return sdesc(saddr, lbo_bytes); }

template <bool DROP, bool ONEBLK>
__global__ void __launch_bounds__(BW_THREADS, 1)
attn_ws_bwd_kernel(const __grid_constant__ BwdMaps M, const Params P, int nkb, const bf16* __restrict__ ctx, int64_t ldctx,
                   const float* __restrict__ lse) {
  using S = BwdSmem<ONEBLK>;
  constexpr int QGS = S::QGS;
  constexpr uint32_t SLOT = S::kSlot;
  extern __shared__ uint8_t raw[];
  uint8_t* sm = raw + ((1024u - (s32(raw) & 1023u)) & 1023u);
  uint8_t* PdS = sm + S::kPdS0;
  uint8_t* QG = sm + S::kQG0;
  uint8_t* KV = sm + S::kKV0;
  float* msk = reinterpret_cast<float*>(sm + S::kMsk0);
  float* dlt = reinterpret_cast<float*>(sm + S::kDlt0);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + S::kBar0);
  uint64_t* qg_full = bars;            // [3]
  uint64_t* qg_empty = bars + 3;       // [3]
  uint64_t* kv_full = bars + 6;        // [3]
  uint64_t* kv_empty = bars + 9;       // [3]
  uint64_t* d_full = bars + 12;        // [3] delta of the tile in QG slot s is in dlt[s]
  uint64_t* d_empty = bars + 15;       // [3]
  uint64_t* s_full = bars + 18;        // [2] S / dP of team e in TMEM
  uint64_t* pds_full = bars + 20;      // [2] P' / dS images of team e written
  uint64_t* pds_empty = bars + 22;     // [2] ... and consumed by B
  // "complete" signals are PER TEAM: a team waits for its own barrier in strictly increasing phases. (One barrier per query
  // tile / key block shared by alternating teams let a team that ran ahead test the parity of a phase two completions
  // away -- indistinguishable from the one already complete.)
  uint64_t* dq_full = bars + 24;       // [2] team e: dQ of the tile of its last-key-block iteration is complete
  uint64_t* dq_empty = bars + 26;      // [2] per query tile: the dQ accumulator has been drained
  uint64_t* dkv_full = bars + 28;      // [2] team e: dK / dV of the key block of its last-query-tile iteration are complete
  uint64_t* dkv_empty = bars + 30;     // the dK / dV accumulators have been drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 31);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_qt = P.n_qt, stride = gridDim.x;

  // stale rows of the rings (beyond the loaded boxes) feed masked columns / zero probabilities: keep them finite
  for (uint32_t i = threadIdx.x * 16; i < S::kMsk0; i += BW_THREADS * 16) *reinterpret_cast<uint4*>(sm + i) = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 3; ++i) {
      mbar_init(&qg_full[i], 1); mbar_init(&qg_empty[i], 1); mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1);
      mbar_init(&d_full[i], ONEBLK ? 1 : 64); mbar_init(&d_empty[i], 256);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1); mbar_init(&pds_full[i], 256); mbar_init(&pds_empty[i], 1);
      mbar_init(&dq_full[i], 1); mbar_init(&dq_empty[i], 256); mbar_init(&dkv_full[i], 1);
    }
    mbar_init(dkv_empty, 256);
    fence_init();
    const CUtensorMap* mp = &M.q0f;
    for (int i = 0; i < (int)(sizeof(BwdMaps) / sizeof(CUtensorMap)); ++i) prefetch_map(mp + i);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  fence_async();
  tc_before();
  __syncthreads();
  tc_after();
  const uint32_t tm = *tmem_slot;
  // TMEM columns: team e: S at e*128, dP at e*128 + 64; dQ of tile qt at 256 + qt*64; dV at 384, dK at 448

  if (warp == 0) {
    // ===================================================================== producer
    It c{(int)blockIdx.x, 0, 0, 0, 0};
    auto load_qg = [&](int it, int qt, int h, int p, int gq0, int gq1) {
      const int tt = it * n_qt + qt, slot = tt % QGS;
      mbar_wait(&qg_empty[slot], ((tt / QGS) & 1) ^ 1, 1);
      if (lane == 0) {
        uint8_t* dq_ = QG + slot * SLOT;
        uint8_t* dg_ = dq_ + TILE_B;
        uint8_t* do_ = dg_ + TILE_B;
        const int r0 = qt * 128;
        const int n0p = min(max(P.ql.rows0p - r0, 0), 128);
        const bool seg1_here = P.ql.rows1 > 0 && P.ql.rows0p >= r0 && P.ql.rows0p < r0 + 128;
        mbar_expect_tx(&qg_full[slot], (ONEBLK ? 2u : 3u) * ((uint32_t)n0p * 128u + (seg1_here ? (uint32_t)P.ql.rows1p * 128u : 0u)));
        if (n0p == 128) {
          tma3(dq_, &M.q0f, &qg_full[slot], h * 64, r0, gq0); tma3(dg_, &M.g0f, &qg_full[slot], h * 64, r0, p);
          if (!ONEBLK) tma3(do_, &M.o0f, &qg_full[slot], h * 64, r0, p);
        } else if (n0p > 0) {
          tma3(dq_, &M.q0t, &qg_full[slot], h * 64, r0, gq0); tma3(dg_, &M.g0t, &qg_full[slot], h * 64, r0, p);
          if (!ONEBLK) tma3(do_, &M.o0t, &qg_full[slot], h * 64, r0, p);
        }
        if (seg1_here) {
          const int off = (P.ql.rows0p - r0) * 128;
          tma3(dq_ + off, &M.q1, &qg_full[slot], h * 64, 0, gq1);
          tma3(dg_ + off, &M.g1, &qg_full[slot], h * 64, 0, p);
          if (!ONEBLK) tma3(do_ + off, &M.o1, &qg_full[slot], h * 64, 0, p);
        }
      }
    };
    auto load_kv = [&](int it, int g, int h, int p, int gk0, int gk1, int gv0, int gv1) {
      const int kb = it * nkb + g, slot = kb % KVS;
      mbar_wait(&kv_empty[slot], ((kb / KVS) & 1) ^ 1, 2);
      {                                  // mask slice of this key block (log2 domain; -inf on padding positions)
        const float* madd = P.mask_add ? P.mask_add + (int64_t)(p / P.mask_div) * P.ld_mask : nullptr;
        for (int j = lane; j < 64; j += 32) {
          const int lk = logical_row(P.kl, g * 64 + j);
          msk[slot * 64 + j] = lk >= 0 ? (madd ? madd[lk] * kLog2e : 0.f) : -INFINITY;
        }
      }
      __syncwarp();
      if (lane == 0) {
        uint8_t* dk_ = KV + slot * 2 * BLK_B;
        uint8_t* dv_ = dk_ + BLK_B;
        const int r0 = g * 64;
        const int n0p = min(max(P.kl.rows0p - r0, 0), 64);
        const bool seg1_here = P.kl.rows1 > 0 && P.kl.rows0p >= r0 && P.kl.rows0p < r0 + 64;
        mbar_expect_tx(&kv_full[slot], 2u * ((uint32_t)n0p * 128u + (seg1_here ? (uint32_t)P.kl.rows1p * 128u : 0u)));
        if (n0p == 64) { tma3(dk_, &M.k0f, &kv_full[slot], h * 64, r0, gk0); tma3(dv_, &M.v0f, &kv_full[slot], h * 64, r0, gv0); }
        else if (n0p > 0) { tma3(dk_, &M.k0t, &kv_full[slot], h * 64, r0, gk0); tma3(dv_, &M.v0t, &kv_full[slot], h * 64, r0, gv0); }
        if (seg1_here) {
          const int off = (P.kl.rows0p - r0) * 128;
          tma3(dk_ + off, &M.k1, &kv_full[slot], h * 64, 0, gk1);
          tma3(dv_ + off, &M.v1, &kv_full[slot], h * 64, 0, gv1);
        }
      }
    };
    for (; c.item < P.items; c.item += stride, ++c.it) {
      const int p = c.item / P.heads, h = c.item - p * P.heads;
      int gq0 = 0, gq1 = 0, gk0 = 0, gk1 = 0, gv0 = 0, gv1 = 0;
      if (lane == 0) {
        gq0 = P.qidx0 ? P.qidx0[p] : p; gk0 = P.kidx0 ? P.kidx0[p] : p; gv0 = P.vidx0 ? P.vidx0[p] : p;
        if (P.ql.rows1) gq1 = P.qidx1 ? P.qidx1[p] : p;
        if (P.kl.rows1) { gk1 = P.kidx1 ? P.kidx1[p] : p; gv1 = P.vidx1 ? P.vidx1[p] : p; }
      }
      // in consumption order: block 0, the query tiles, the remaining blocks
      load_kv(c.it, 0, h, p, gk0, gk1, gv0, gv1);
      for (int qt = 0; qt < n_qt; ++qt) load_qg(c.it, qt, h, p, gq0, gq1);
      for (int g = 1; g < nkb; ++g) load_kv(c.it, g, h, p, gk0, gk1, gv0, gv1);
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (one thread)
    if (lane == 0) {
      int my_items = 0;
      for (int i = blockIdx.x; i < P.items; i += stride) ++my_items;
      const int total = my_items * nkb * n_qt;
      It ca{(int)blockIdx.x, 0, 0, 0, 0}, cb = ca;
      auto a_ready = [&](const It& c) -> bool {
        const int kb = c.it * nkb + c.g, tt = c.it * n_qt + c.qt;
        if (!mbar_test(&kv_full[kb % KVS], (kb / KVS) & 1)) return false;
        return mbar_test(&qg_full[tt % QGS], (tt / QGS) & 1);
      };
      auto b_ready = [&](const It& c) -> bool {
        const int e = c.n & 1, u = c.n >> 1, kb = c.it * nkb + c.g;
        if (!mbar_test(&pds_full[e], u & 1)) return false;
        if (c.qt == 0 && !mbar_test(dkv_empty, (kb & 1) ^ 1)) return false;
        if (c.g == 0 && !mbar_test(&dq_empty[c.qt], (c.it & 1) ^ 1)) return false;
        return true;
      };
      auto issue_a = [&](const It& c) {
        const int kb = c.it * nkb + c.g, tt = c.it * n_qt + c.qt, e = c.n & 1;
        tc_after();
        const uint32_t q = s32(QG + (tt % QGS) * SLOT), dO = q + TILE_B;
        const uint32_t k = s32(KV + (kb % KVS) * 2 * BLK_B), v = k + BLK_B;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma(tm + e * 128, sdesc(q + kk * 32, 16), sdesc(k + kk * 32, 16), idesc(64, 0, 0), kk > 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma(tm + e * 128 + 64, sdesc(dO + kk * 32, 16), sdesc(v + kk * 32, 16), idesc(64, 0, 0), kk > 0 ? 1u : 0u);
        commit(&s_full[e]);
      };
      auto issue_b = [&](const It& c) {
        const int kb = c.it * nkb + c.g, tt = c.it * n_qt + c.qt, e = c.n & 1;
        tc_after();
        const uint32_t q = s32(QG + (tt % QGS) * SLOT), dO = q + TILE_B;
        const uint32_t k = s32(KV + (kb % KVS) * 2 * BLK_B);
        const uint32_t pimg = s32(PdS + e * 2 * TILE_B), dsimg = pimg + TILE_B;
        const int ks_k = (min(64, P.kl.total - c.g * 64) + 15) >> 4;         // 16-key steps that hold real keys
        const int ks_q = (min(128, P.ql.total - c.qt * 128) + 15) >> 4;      // 16-row steps that hold real queries
        for (int kk = 0; kk < ks_k; ++kk)                                      // dQ_qt += dS . K_g
          umma(tm + 256 + c.qt * 64, sdesc(dsimg + kk * 32, 16), sdesc(k + kk * 2048, 8192), idesc(64, 0, 1), (c.g > 0 || kk > 0) ? 1u : 0u);
        for (int kk = 0; kk < ks_q; ++kk)                                      // dV_g += P'^T . dO_qt
          umma(tm + 384, sdesc_mn(pimg + kk * 2048, TILE_B), sdesc(dO + kk * 2048, 8192), idesc(64, 1, 1), (c.qt > 0 || kk > 0) ? 1u : 0u);
        for (int kk = 0; kk < ks_q; ++kk)                                      // dK_g += dS^T . Q_qt
          umma(tm + 448, sdesc_mn(dsimg + kk * 2048, TILE_B), sdesc(q + kk * 2048, 8192), idesc(64, 1, 1), (c.qt > 0 || kk > 0) ? 1u : 0u);
        commit(&pds_empty[e]);
        if (c.qt == n_qt - 1) { commit(&dkv_full[e]); commit(&kv_empty[kb % KVS]); }
        if (c.g == nkb - 1) { commit(&dq_full[e]); commit(&qg_empty[tt % QGS]); }
      };
      const long long t0 = clock64();
      long long last = t0;
      while (cb.n < total) {
        bool progressed = false;
        if (ca.n < total && ca.n < cb.n + 2 && a_ready(ca)) { issue_a(ca); next(ca, nkb, n_qt, stride); progressed = true; }
        if (cb.n < ca.n && b_ready(cb)) { issue_b(cb); next(cb, nkb, n_qt, stride); progressed = true; }
        if (progressed) last = clock64();
        else {
          __nanosleep(32);
#ifndef FCMF_WS_NO_TRAP
          if (clock64() - last > 9000000000LL) {                   // later than the bounded waits of the other roles: they name the barrier
            const int e = cb.n & 1, u = cb.n >> 1, kb = cb.it * nkb + cb.g;
            printf("fcmf attn_ws_bwd: MMA issuer stalled (block %d, A %d, B %d of %d; B waits: pds_full %d dkv_empty %d dq_empty %d)\n",
                   (int)blockIdx.x, ca.n, cb.n, total, (int)mbar_test(&pds_full[e], u & 1),
                   (int)(cb.qt != 0 || mbar_test(dkv_empty, (kb & 1) ^ 1)), (int)(cb.g != 0 || mbar_test(&dq_empty[cb.qt], (cb.it & 1) ^ 1)));
            __trap();
          }
#endif
        }
      }
    }
  } else if (warp < 4) {
    // ===================================================================== delta = rowsum(dO o O) and lse*log2e, 64 threads
    // (more than one key block only) thread t owns tile rows t and t + 64; O and dO tiles are read from shared memory
    if (!ONEBLK) {
      const int t2 = (warp - 2) * 32 + lane;
      int it = 0;
      for (int item = blockIdx.x; item < P.items; item += stride, ++it) {
        const int p = item / P.heads, h = item - p * P.heads;
        for (int qt = 0; qt < n_qt; ++qt) {
          const int tt = it * n_qt + qt, slot = tt % QGS;
          float l2v[2];
#pragma unroll
          for (int i = 0; i < 2; ++i) {                              // in flight while the tiles land
            const int lrow = logical_row(P.ql, qt * 128 + t2 + i * 64);
            l2v[i] = lrow >= 0 ? lse[((int64_t)p * P.heads + h) * P.Lq + lrow] : INFINITY;   // raw (scaled at the store below: no wait here); padded rows: P = exp2(. - inf) = 0
          }
          mbar_wait(&d_empty[slot], ((tt / QGS) & 1) ^ 1, 3);
          mbar_wait(&qg_full[slot], (tt / QGS) & 1, 4);
          const uint8_t* dO = QG + slot * SLOT + TILE_B;
          const uint8_t* O = dO + TILE_B;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int r = t2 + i * 64;
            float part = 0.f;
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
              const uint4 ov = *reinterpret_cast<const uint4*>(O + swz(r, ch));
              const uint4 gv = *reinterpret_cast<const uint4*>(dO + swz(r, ch));
              const __nv_bfloat162* ho = reinterpret_cast<const __nv_bfloat162*>(&ov);
              const __nv_bfloat162* hg = reinterpret_cast<const __nv_bfloat162*>(&gv);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 fo = __bfloat1622float2(ho[j]), fg = __bfloat1622float2(hg[j]);
                part = fmaf(fo.x, fg.x, fmaf(fo.y, fg.y, part));
              }
            }
            dlt[slot * 256 + r] = part;                              // padded rows: O and dO are zero-filled
            dlt[slot * 256 + 128 + r] = l2v[i] * kLog2e;
          }
          mbar_arrive(&d_full[slot]);
        }
      }
    }
  } else {
    // ===================================================================== softmax teams (iterations n = e, e+2, ...)
    const int e = (warp - 4) >> 3, quad = warp & 3, half = ((warp - 4) >> 2) & 1;
    const int trow = quad * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const uint32_t tS = tm + e * 128 + half * 32 + lane_addr, tP = tS + 64;
    const float scale2 = P.scale * kLog2e;
    uint8_t* pimg = PdS + e * 2 * TILE_B;
    uint8_t* dsimg = pimg + TILE_B;
    const bool st_thread = (warp - 4) == e * 8 && lane == 0;
    const int gap = P.kl.rows0p - P.kl.rows0;
    DropCfg dc;
    if (DROP) dc = make_drop(P.drop);
    const float log2_inv_keep = DROP ? log2f(dc.inv_keep) : 0.f, keep_prob = DROP ? 1.0f / dc.inv_keep : 1.0f;
    It c{(int)blockIdx.x, 0, 0, 0, 0};
    if (e == 1) next(c, nkb, n_qt, stride);
    auto load_l2 = [&](const It& x) -> float {                        // ONEBLK: lse of this thread's row, prefetched one iteration ahead
      if (x.item >= P.items) return INFINITY;
      const int lr = logical_row(P.ql, x.qt * 128 + trow);
      const int xp = x.item / P.heads, xh = x.item - xp * P.heads;
      return lr >= 0 ? lse[((int64_t)xp * P.heads + xh) * P.Lq + lr] : INFINITY;      // RAW: scaling it here would wait for the load at once
    };
    float l2_next = ONEBLK ? load_l2(c) : 0.f;
    uint32_t n_dkv = 0, n_dq = 0;                                  // epilogues this team has run: phases of its "complete" barriers
    while (c.item < P.items) {
      const int p = c.item / P.heads, h = c.item - p * P.heads;
      const int kb = c.it * nkb + c.g, tt = c.it * n_qt + c.qt, u = c.n >> 1;
      const int kvs = kb % KVS, qs = tt % QGS;
      const int r0 = c.qt * 128;
      const int lrow = logical_row(P.ql, r0 + trow);
      const int wlo = r0 + quad * 32, whi = wlo + 32;
      const bool wact = wlo < P.ql.rows0 || (P.ql.rows1 > 0 && wlo < P.ql.rows0p + P.ql.rows1 && whi > P.ql.rows0p);
      float l2 = l2_next * kLog2e, dl = 0.f;                      // the prefetched value is consumed one iteration after its load
      if (ONEBLK) {
        It cn = c;
        next(cn, nkb, n_qt, stride);
        if (cn.item < P.items) next(cn, nkb, n_qt, stride);
        l2_next = load_l2(cn);
      }
      mbar_wait(&kv_full[kvs], (kb / KVS) & 1, 5);                 // mask slice visible
      if (!ONEBLK) {
        mbar_wait(&d_full[qs], (tt / QGS) & 1, 6);                 // delta and lse*log2e of the tile visible
        dl = dlt[qs * 256 + trow];
        l2 = dlt[qs * 256 + 128 + trow];
      }
      // train(): P carries the 1/(1-p) of the kept entries from the exponent (pI = P / (1 - p)), so the per-element products
      // P' = keep * P / (1-p) and dP' = keep * dP / (1-p) become selects: dS = P (dP' - delta) = pI (keep * dP - delta (1-p))
      if (DROP) { l2 -= log2_inv_keep; dl *= keep_prob; }
      mbar_wait(&s_full[e], u & 1, 7);
      tc_after();
      // P' / dS images of this team: consumed by B two iterations ago; staging reads of its TMA stores have finished
      if (st_thread) tma_store_wait_read();
      mbar_wait(&pds_empty[e], (u & 1) ^ 1, 8);
      team_bar(e);
      const float* m = msk + kvs * 64 + half * 32;
      uint32_t rseed = 0;
      if (DROP) rseed = drop_rowseed(dc.seed, ((uint64_t)p * (uint64_t)P.heads + (uint64_t)h) * (uint64_t)P.Lq + (uint64_t)max(lrow, 0));
      const int x0 = c.g * 64 + half * 32;                         // padded key position of this thread's first column
      const bool pair_path = x0 + 32 <= P.kl.rows0p;               // inside segment 0: padded position == key index
      auto keep2 = [&](int xa, bool& k0, bool& k1) {                // dropout keep bits of the column pair (xa, xa + 1)
        if (pair_path) {
          const uint32_t hsh = drop_pair(rseed, (uint32_t)xa);
          k0 = drop_keep_lo(hsh, dc.thr16); k1 = drop_keep_hi(hsh, dc.thr16);
        } else {
          const int xb = xa + 1;
          k0 = drop_keep(rseed, (uint32_t)(xa < P.kl.rows0p ? xa : xa - gap), dc.thr16);
          k1 = drop_keep(rseed, (uint32_t)(xb < P.kl.rows0p ? xb : xb - gap), dc.thr16);
        }
      };
      if (!ONEBLK) {
        if (wact) {
#pragma unroll
          for (int sc = 0; sc < 2; ++sc) {                         // two 16-column sub-chunks (register budget)
            uint32_t rs[16], rp[16];
            float pv[16], dsv[16];
            tmem_ld16(tS + sc * 16, rs);
            tmem_ld16(tP + sc * 16, rp);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              const float2 m2 = *reinterpret_cast<const float2*>(m + sc * 16 + j);
              const float2 xe = __fadd2_rn(__ffma2_rn(u2f2(rs[j], rs[j + 1]), f2(scale2, scale2), m2), f2(-l2, -l2));
              const float p0 = ex2_approx(xe.x), p1 = ex2_approx(xe.y);
              float d0 = __uint_as_float(rp[j]), d1 = __uint_as_float(rp[j + 1]);
              float q0 = p0, q1 = p1;
              if (DROP) {
                bool k0, k1;
                keep2(x0 + sc * 16 + j, k0, k1);
                q0 = k0 ? p0 : 0.f; d0 = k0 ? d0 : 0.f;
                q1 = k1 ? p1 : 0.f; d1 = k1 ? d1 : 0.f;
              }
              pv[j] = q0; pv[j + 1] = q1;
              const float2 ds2 = __fmul2_rn(f2(p0, p1), __fadd2_rn(f2(d0, d1), f2(-dl, -dl)));
              dsv[j] = ds2.x; dsv[j + 1] = ds2.y;
            }
            store_row16(pimg, trow, half * 32 + sc * 16, pv);
            store_row16(dsimg, trow, half * 32 + sc * 16, dsv);
          }
        }
      } else {
        // one key block: delta_i = sum_j P_ij dP'_ij from the registers of the two threads that share row i
        float pk[32];
        uint32_t kbits = 0xffffffffu;
        float part = 0.f;
        float* red = dlt + e * 256;                                 // [half][128]
        if (wact) {
#pragma unroll
          for (int sc = 0; sc < 2; ++sc) {
            uint32_t rs[16], rp[16];
            float pv[16];
            tmem_ld16(tS + sc * 16, rs);
            tmem_ld16(tP + sc * 16, rp);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              const float2 m2 = *reinterpret_cast<const float2*>(m + sc * 16 + j);
              const float2 xe = __fadd2_rn(__ffma2_rn(u2f2(rs[j], rs[j + 1]), f2(scale2, scale2), m2), f2(-l2, -l2));
              const float p0 = ex2_approx(xe.x), p1 = ex2_approx(xe.y);
              float d0 = __uint_as_float(rp[j]), d1 = __uint_as_float(rp[j + 1]);
              float q0 = p0, q1 = p1;
              if (DROP) {
                bool k0, k1;
                keep2(x0 + sc * 16 + j, k0, k1);
                if (!k0) kbits &= ~(1u << (sc * 16 + j));
                if (!k1) kbits &= ~(2u << (sc * 16 + j));
                q0 = k0 ? p0 : 0.f; d0 = k0 ? d0 : 0.f;
                q1 = k1 ? p1 : 0.f; d1 = k1 ? d1 : 0.f;
              }
              pv[j] = q0; pv[j + 1] = q1;
              pk[sc * 16 + j] = p0; pk[sc * 16 + j + 1] = p1;
              part = fmaf(p0, d0, fmaf(p1, d1, part));
            }
            store_row16(pimg, trow, half * 32 + sc * 16, pv);
          }
          red[half * 128 + trow] = part;
        }
        team_bar(e);
        if (wact) {
          dl = red[trow] + red[128 + trow];
          if (DROP) dl *= keep_prob;
#pragma unroll
          for (int sc = 0; sc < 2; ++sc) {
            uint32_t rp[16];
            float dsv[16];
            tmem_ld16(tP + sc * 16, rp);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float d = __uint_as_float(rp[j]);
              if (DROP) d = ((kbits >> (sc * 16 + j)) & 1u) ? d : 0.f;
              dsv[j] = pk[sc * 16 + j] * (d - dl);
            }
            store_row16(dsimg, trow, half * 32 + sc * 16, dsv);
          }
        }
      }
      fence_async();
      tc_before();
      mbar_arrive(&pds_full[e]);

      if (c.qt == n_qt - 1) {
        // ---- dV_g / dK_g: lanes 0..63 = the keys of block g (quadrants 0, 1); staged in the P' image, then TMA stores
        mbar_wait(&dkv_full[e], n_dkv & 1, 9);
        ++n_dkv;
        tc_after();
        if (quad < 2) {
          uint32_t rv[32], rk[32];
          tmem_ld32(tm + 384 + half * 32 + lane_addr, rv);
          tmem_ld32(tm + 448 + half * 32 + lane_addr, rk);
          tmem_ld_wait();
          stage_out32(pimg, trow, half * 32, rv, 1.0f);
          stage_out32(pimg + BLK_B, trow, half * 32, rk, P.scale);
        }
        tc_before();
        fence_async();
        mbar_arrive(dkv_empty);
        team_bar(e);
        if (st_thread) {
          const int k0 = c.g * 64;
          const int n0p = min(max(P.kl.rows0p - k0, 0), 64);
          const bool seg1_here = P.kl.rows1 > 0 && P.kl.rows0p >= k0 && P.kl.rows0p < k0 + 64;
          if (n0p == 64) { tma3_store(&M.dv0f, pimg, h * 64, k0, p); tma3_store(&M.dk0f, pimg + BLK_B, h * 64, k0, p); }
          else if (n0p > 0) { tma3_store(&M.dv0t, pimg, h * 64, k0, p); tma3_store(&M.dk0t, pimg + BLK_B, h * 64, k0, p); }
          if (seg1_here) {
            const int off = (P.kl.rows0p - k0) * 128;
            tma3_store(&M.dv1, pimg + off, h * 64, 0, p);
            tma3_store(&M.dk1, pimg + BLK_B + off, h * 64, 0, p);
          }
          tma_store_commit();
        }
      }
      if (c.g == nkb - 1) {
        // ---- dQ_qt: staged in the dS image
        mbar_wait(&dq_full[e], n_dq & 1, 10);
        ++n_dq;
        tc_after();
        if (wact) {
          uint32_t rq[32];
          tmem_ld32(tm + 256 + c.qt * 64 + half * 32 + lane_addr, rq);
          tmem_ld_wait();
          stage_out32(dsimg, trow, half * 32, rq, P.scale);
        }
        tc_before();
        fence_async();
        mbar_arrive(&dq_empty[c.qt]);
        if (!ONEBLK) mbar_arrive(&d_empty[qs]);
        team_bar(e);
        if (st_thread) {
          const int n0p = min(max(P.ql.rows0p - r0, 0), 128);
          const bool seg1_here = P.ql.rows1 > 0 && P.ql.rows0p >= r0 && P.ql.rows0p < r0 + 128;
          if (n0p == 128) tma3_store(&M.dq0f, dsimg, h * 64, r0, p);
          else if (n0p > 0) tma3_store(&M.dq0t, dsimg, h * 64, r0, p);
          if (seg1_here) tma3_store(&M.dq1, dsimg + (P.ql.rows0p - r0) * 128, h * 64, 0, p);
          tma_store_commit();
        }
      }
      next(c, nkb, n_qt, stride);
      if (c.item < P.items) next(c, nkb, n_qt, stride);
    }
    if (st_thread) tma_store_wait_all();
  }
  tc_before();
  __syncthreads();
  if (warp == 1) { tc_after(); tmem_dealloc(tm, 512); }
}

}  // namespace ws

bool attn_ws_bwd_supported(const AttnDev& a, int64_t ldctx, const void* ctx, int64_t lddctx) {
  using namespace ws;
  if (!attn_ws_supported(a, ldctx, ctx) || (lddctx % 8)) return false;
  const RowLay ql = row_lay(a.q), kl = row_lay(a.k);
  if (ql.total > 256) return false;                        // two query tiles: dQ accumulators of both live in TMEM
  return tiles_ok(kl, 64);
}

int attn_ws_bwd(const AttnDev& a, const void* ctx, int64_t ldctx, const void* dctx, int64_t lddctx, const float* lse,
                void* dq, void* dk, void* dv, cudaStream_t st) {
  using namespace ws;
  Params P;
  fill_params(a, &P);
  BwdMaps M;
  const int64_t HD = (int64_t)a.heads * 64;
  if (int r = role_maps(a.q, P.ql, a.heads, 128, &M.q0f, &M.q0t, &M.q1)) return r;
  if (int r = role_maps(a.k, P.kl, a.heads, 64, &M.k0f, &M.k0t, &M.k1)) return r;
  if (int r = role_maps(a.v, P.kl, a.heads, 64, &M.v0f, &M.v0t, &M.v1)) return r;
  // per-problem tensors [NP][L][HD] as two virtual row segments (rows [0, rows0) and [rows0, L) of every problem)
  auto virt = [&](const void* base, int64_t ld, int L, const RowLay& lay, int tile, CUtensorMap* f, CUtensorMap* t, CUtensorMap* s1) -> int {
    const int64_t gs = (int64_t)L * ld;
    const int tail = lay.rows0p % tile;
    if (int r = make_map3(f, base, HD, lay.rows0, a.NP, ld, gs, tile)) return r;
    if (int r = make_map3(t, base, HD, lay.rows0, a.NP, ld, gs, tail ? tail : 8)) return r;
    if (lay.rows1) return make_map3(s1, (const bf16*)base + (int64_t)lay.rows0 * ld, HD, lay.rows1, a.NP, ld, gs, lay.rows1p);
    *s1 = *f;
    return 0;
  };
  if (int r = virt(dctx, lddctx, a.Lq, P.ql, 128, &M.g0f, &M.g0t, &M.g1)) return r;
  if (int r = virt(dq, HD, a.Lq, P.ql, 128, &M.dq0f, &M.dq0t, &M.dq1)) return r;
  if (int r = virt(dk, HD, a.Lk, P.kl, 64, &M.dk0f, &M.dk0t, &M.dk1)) return r;
  if (int r = virt(dv, HD, a.Lk, P.kl, 64, &M.dv0f, &M.dv0t, &M.dv1)) return r;
  const int nkb = key_blocks(P.kl);
  if (nkb > 1) { if (int r = virt(ctx, ldctx, a.Lq, P.ql, 128, &M.o0f, &M.o0t, &M.o1)) return r; }
  else { M.o0f = M.g0f; M.o0t = M.g0t; M.o1 = M.g1; }
  const unsigned grid = (unsigned)std::min<int64_t>(P.items, sm_count());
#define WS_BWD(DR, OB)                                                                                              \
  {                                                                                                                 \
    auto kern = attn_ws_bwd_kernel<DR, OB>;                                                                         \
    FCMF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BwdSmem<OB>::kSmem)); \
    kern<<<grid, BW_THREADS, BwdSmem<OB>::kSmem, st>>>(M, P, nkb, (const bf16*)ctx, ldctx, lse);                    \
  }
  const bool drop = a.drop.p > 0.f;
  if (nkb == 1) { if (drop) WS_BWD(true, true) else WS_BWD(false, true) }
  else { if (drop) WS_BWD(true, false) else WS_BWD(false, false) }
#undef WS_BWD
  FCMF_LAUNCH_OK();
  return 0;
}

}  // namespace fcmf
