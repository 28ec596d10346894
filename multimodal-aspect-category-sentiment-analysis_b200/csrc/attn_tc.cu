// tcgen05 / TMEM folded attention for head_dim 64, bf16 (forward, dQ, dK/dV).
//
// One CTA = one (problem, head, 128-row tile). Everything is organised in 64-wide blocks, so every contraction is
// the same "block MMA": D[128 x 64] (+)= A[128 x 64] . B[64 x 64], issued as four tcgen05.mma (M128 N64 K16) by one
// thread, operands in shared memory in the 128-byte-swizzled row format (row r at r*128 B, 16-byte chunk c stored at
// chunk c ^ (r & 7)) that both TMA and the UMMA descriptors of gemm_tc.cu use:
//   * "row operand"  (K-major):  rows = M or N index, the 64 columns are the reduction  (Q.K^T, dO.V^T, K.Q^T, V.dO^T)
//   * "col operand"  (MN-major): rows = reduction index, the 64 columns are N            (P.V, dS.K, P^T.dO, dS^T.Q)
// the SAME shared-memory image serves as either, only the descriptor differs.
// Scores live in TMEM; thread t of each warpgroup owns TMEM lane t (= one query row, or one key row in the dK/dV
// kernel) and the two warpgroups split the 32-column chunks: tcgen05.ld, softmax arithmetic in registers, bf16
// probabilities written back to shared memory as the A operand of the second contraction. Tiles are staged with
// cp.async (16-byte, zero-fill for padded rows) so every global load of a tile is in flight at once; two CTAs share an
// SM (<= 112 KB of shared memory, <= 256 TMEM columns each), which overlaps one CTA's loads with the other's MMAs.
//
// Reference semantics: BertCoAttention/BertSelfAttention.forward (mm_modeling.py:193-266): scale before the mask
// add, additive -10000 mask on the keys, softmax over keys, context = P.V; backward = autograd of the same.
#include "common.cuh"
#include "attn.h"

namespace fcmf {

constexpr int ATC_THREADS = 256;                // 2 warpgroups: both own the 128 TMEM lanes, they split the column chunks
constexpr int ATC_TILE = 128;                   // rows per CTA tile (UMMA M)
constexpr int ATC_BLK = 64;                     // block width
constexpr uint32_t ATC_TILE_BYTES = ATC_TILE * 128;   // 16 KB
constexpr uint32_t ATC_BLK_BYTES = ATC_BLK * 128;     // 8 KB
constexpr float kLog2e = 1.44269504088896340736f;

// ------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t a_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void a_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a_smem_u32(bar)), "r"(count));
}
// suspend-time hint: the waiting threads sleep in hardware until the MMA's commit flips the phase (see gemm_tc.cu)
__device__ __forceinline__ bool a_mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  const uint32_t hint_ns = 100000;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.b32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(a_smem_u32(bar)), "r"(parity), "r"(hint_ns) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void a_mbar_wait(uint64_t* bar, uint32_t parity) {
  if (a_mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!a_mbar_try(bar, parity)) {
    if ((++spins & 0x3f) == 0 && clock64() - t0 > 6000000000LL) {
      printf("fcmf attn_tc: mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void a_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void a_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void a_tc_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void a_tc_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void a_tmem_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a_smem_u32(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void a_tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void a_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void a_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(a_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void a_tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint64_t a_desc(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;                       // 8-row groups are 1024 B apart
  d |= (uint64_t)1 << 46;                                 // descriptor version
  d |= (uint64_t)2 << 61;                                 // SWIZZLE_128B
  return d;
}
constexpr uint32_t a_idesc(int b_mn_major) {             // bf16 x bf16 -> f32, M = 128, N = 64
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(64 >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}
// D[128x64] (+)= A_rowop[128 x 64] . B_rowop[64 x 64]^T   (reduction over the 64 columns of both)
__device__ __forceinline__ void block_mma_rr(uint32_t tmem_d, uint32_t a_tile, uint32_t b_blk, bool accumulate) {
#pragma unroll
  for (int k = 0; k < 4; ++k)
    a_umma(tmem_d, a_desc(a_tile + k * 32, 16), a_desc(b_blk + k * 32, 16), a_idesc(0), (accumulate || k > 0) ? 1u : 0u);
}
// D[128x64] (+)= A_rowop[128 x 64] . B_colop[64 rows(reduction) x 64 (N)]; only the first `ksteps` 16-row steps
__device__ __forceinline__ void block_mma_rc(uint32_t tmem_d, uint32_t a_tile, uint32_t b_blk, bool accumulate, int ksteps) {
  for (int k = 0; k < ksteps; ++k)
    a_umma(tmem_d, a_desc(a_tile + k * 32, 16), a_desc(b_blk + k * 2048, 8192), a_idesc(1), (accumulate || k > 0) ? 1u : 0u);
}

// ------------------------------------------------------------------------------------------- tile staging
__device__ __forceinline__ uint32_t swz(int r, int chunk) { return (uint32_t)(r * 128 + ((chunk ^ (r & 7)) << 4)); }

__device__ __forceinline__ void cp_async16(uint8_t* dst, const void* src, bool valid) {
  const uint32_t n = valid ? 16u : 0u;                         // src-size 0 => the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(a_smem_u32(dst)), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---- row sources of one q/k/v operand for one (problem, head) ------------------------------------------------
// The first version recomputed 64-bit segment addresses (and re-loaded the group indices) for every 16-byte chunk:
// ncu attributed 26 % of the forward kernel's instructions and ~20 % of its stall samples to staging. Now: the group
// indices of the NEXT item are fetched while the current one computes (SegIdx), the two segment base pointers are
// formed once per item (RowSrc, the thread's 16-byte column chunk and the head offset folded in), and a thread's rows
// (r_t + 32 k) advance by one 32-bit offset add per chunk on the common path.
struct SegIdx { int g0, g1; };
__device__ __forceinline__ SegIdx seg_idx(const SegDev (&s)[2], int p) {
  SegIdx g;
  g.g0 = s[0].idx ? s[0].idx[p] : p;
  g.g1 = (s[1].rows && s[1].idx) ? s[1].idx[p] : p;
  return g;
}
struct RowSrc {
  const uint8_t* b0;          // row 0 of segment 0 (+ head, + this thread's chunk)
  const uint8_t* b1;          // row 0 of segment 1 (== b0 when the segment is absent)
  uint32_t ldb0, ldb1;        // row strides in BYTES
  int rows0;                  // rows in segment 0
};
__device__ __forceinline__ RowSrc row_src(const SegDev (&s)[2], SegIdx g, int h) {
  RowSrc r;
  const int64_t col = (int64_t)h * 64 + (threadIdx.x & 7) * 8;
  r.rows0 = s[0].rows;
  r.ldb0 = (uint32_t)s[0].ld * 2u;
  r.b0 = reinterpret_cast<const uint8_t*>(reinterpret_cast<const bf16*>(s[0].ptr) + (int64_t)g.g0 * s[0].rows * s[0].ld + col);
  if (s[1].rows) {
    r.ldb1 = (uint32_t)s[1].ld * 2u;
    r.b1 = reinterpret_cast<const uint8_t*>(reinterpret_cast<const bf16*>(s[1].ptr) + (int64_t)g.g1 * s[1].rows * s[1].ld + col);
  } else { r.ldb1 = r.ldb0; r.b1 = r.b0; }
  return r;
}
__device__ __forceinline__ RowSrc row_src_plain(const bf16* base, int64_t ld, int64_t prow0, int h) {   // [NP*L, ld] activation
  RowSrc r;
  r.rows0 = 0x7fffffff;
  r.ldb0 = r.ldb1 = (uint32_t)ld * 2u;
  r.b0 = r.b1 = reinterpret_cast<const uint8_t*>(base + prow0 * ld + (int64_t)h * 64 + (threadIdx.x & 7) * 8);
  return r;
}
// rows [row0, row0+tile_rows) of the operand -> 128B-swizzled tile (row r at r*128, chunk c at c ^ (r & 7)); rows >= L are 0
__device__ __forceinline__ void stage_rows(uint8_t* tile, const RowSrc& src, int row0, int tile_rows, int L) {
  const int r_t = threadIdx.x >> 3;                              // this thread's first tile row; its rows are r_t + 32 k
  uint8_t* dst = tile + r_t * 128 + ((((int)threadIdx.x & 7) ^ (r_t & 7)) << 4);      // (r & 7) is the same for all of them
  int gr = row0 + r_t;
  const int end = row0 + tile_rows;
  const int fast_end = min(min(src.rows0, L), end);              // rows of segment 0 that exist
  uint32_t off = (uint32_t)gr * src.ldb0;
  const uint32_t step = (uint32_t)(ATC_THREADS / 8) * src.ldb0;
  for (; gr < fast_end; gr += ATC_THREADS / 8, dst += (ATC_THREADS / 8) * 128, off += step)
    cp_async16(dst, src.b0 + off, true);
  for (; gr < end; gr += ATC_THREADS / 8, dst += (ATC_THREADS / 8) * 128) {            // segment 1 rows and the zero padding
    const bool ok = gr < L;
    const uint8_t* sp = gr < src.rows0 ? src.b0 + (uint32_t)gr * src.ldb0 : src.b1 + (uint32_t)(gr - src.rows0) * src.ldb1;
    cp_async16(dst, ok ? sp : src.b0, ok);
  }
}
// thread-owned row: write 32 consecutive bf16 (cols c0..c0+31 of a 64-col block) of row r into a swizzled tile
__device__ __forceinline__ void store_row32(uint8_t* tile, int r, int c0, const float (&v)[32]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint4 w;
    __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&w);
#pragma unroll
    for (int j = 0; j < 4; ++j) hh[j] = __floats2bfloat162_rn(v[g * 8 + 2 * j], v[g * 8 + 2 * j + 1]);
    *reinterpret_cast<uint4*>(tile + swz(r, (c0 >> 3) + g)) = w;
  }
}
// 32 fp32 TMEM values * scale -> 32 bf16 to global (64 contiguous bytes)
__device__ __forceinline__ void store_out32(bf16* o, const uint32_t (&r)[32], float sc) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint4 w;
    __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&w);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 t = __fmul2_rn(make_float2(__uint_as_float(r[g * 8 + 2 * j]), __uint_as_float(r[g * 8 + 2 * j + 1])), make_float2(sc, sc));
      hh[j] = __floats2bfloat162_rn(t.x, t.y);
    }
    *reinterpret_cast<uint4*>(o + g * 8) = w;
  }
}

struct AtcShared {
  uint64_t bar;
  uint32_t tmem;
};

// 1024-byte alignment of the dynamic shared memory WITHOUT laundering the pointer through an integer: base + offset keeps
// the shared address space, so the compiler emits LDS / STS (the uintptr_t round trip made every mask load and P store a
// generic LD / ST, tracked on the long scoreboard).
__device__ __forceinline__ uint8_t* align1k(uint8_t* p) {
  return p + ((1024u - (a_smem_u32(p) & 1023u)) & 1023u);
}

// One-time CTA setup of the persistent kernels: mbarrier, TMEM columns. Returns the TMEM base address.
__device__ __forceinline__ uint32_t atc_setup(AtcShared* sh, uint32_t cols) {
  if (threadIdx.x == 0) { a_mbar_init(&sh->bar, 1); a_fence_init(); }
  if ((threadIdx.x >> 5) == 0) a_tmem_alloc(&sh->tmem, cols);
  a_tc_before();
  __syncthreads();
  a_tc_after();
  return sh->tmem;
}
__device__ __forceinline__ void atc_teardown(uint32_t tm, uint32_t cols) {
  a_tc_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) { a_tc_after(); a_tmem_dealloc(tm, cols); }
}
// drop the probabilities of columns col0 .. col0+31 of the row with seed `rseed` (dropout AFTER the softmax sum)
__device__ __forceinline__ void drop_row32(float (&v)[32], uint32_t rseed, uint32_t col0, uint32_t thr16) {
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    const uint32_t hsh = drop_pair(rseed, col0 + j);
    v[j] = drop_keep_lo(hsh, thr16) ? v[j] : 0.f;
    v[j + 1] = drop_keep_hi(hsh, thr16) ? v[j + 1] : 0.f;
  }
}

// Packed fp32x2 arithmetic (FFMA2 / FADD2 / FMUL2, sm_100): two lanes' worth of IEEE fp32 operations per issue slot. The
// softmax loops are issue-limited (ncu: 34-42 % of slots, 16-24 warps per SM), and every operation here is the same
// round-to-nearest fp32 operation as its scalar form, so results do not change.
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 u2f2(uint32_t a, uint32_t b) { return make_float2(__uint_as_float(a), __uint_as_float(b)); }

// All three kernels are PERSISTENT: a CTA walks over work items (item = blockIdx.x + n * gridDim.x) with its mbarrier and
// TMEM columns set up once; what an item shares between its tiles (K/V and the mask for the query-tiled kernels) is
// staged once per item; warps whose 32 tile rows lie entirely beyond the sequence end skip the softmax arithmetic
// (170 query rows = 128 + 42: two of the second tile's four row-warps are idle; 49 keys in a 128-row key tile: two of
// four) -- the first version spent the same issue slots on padding as on data (ncu: issue-bound, ~50 % of slots).

// ------------------------------------------------------------------------------------------- forward
// item = (problem, head); inner loop over the 128-row query tiles.
// smem: [Q tile 16K = P block 0 (Q is dead once S is in TMEM)][K blocks NKB*8K][V blocks NKB*8K][P blocks 1.. (NKB-1)*16K]
//       [mask NKB*64 f32][red 2*2*128 f32][AtcShared]
template <int NKB, bool DROP>
__global__ void __launch_bounds__(ATC_THREADS, (NKB == 1 ? 3 : 2))
attn_tc_fwd_kernel(AttnDev a, bf16* __restrict__ ctx, int64_t ldctx, float* __restrict__ lse, int mtiles, int items) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = align1k(raw);
  uint8_t* Qs = sm;
  uint8_t* Ks = Qs + ATC_TILE_BYTES;
  uint8_t* Vs = Ks + NKB * ATC_BLK_BYTES;
  uint8_t* Px = Vs + NKB * ATC_BLK_BYTES;
  float* msk = reinterpret_cast<float*>(Px + (NKB - 1) * ATC_TILE_BYTES);
  float* red = msk + NKB * 64;                                  // [2 passes][2 groups][128 rows]
  AtcShared* sh = reinterpret_cast<AtcShared*>(red + 512);
  constexpr uint32_t kCols = (NKB + 1) * 64 <= 128 ? 128 : ((NKB + 1) * 64 <= 256 ? 256 : 512);
  auto Pblk = [&](int b) -> uint8_t* { return b == 0 ? Qs : Px + (b - 1) * ATC_TILE_BYTES; };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = warp >> 2, trow = (warp & 3) * 32 + lane;     // TMEM lane / tile row owned by this thread
  const int Lq = a.Lq, Lk = a.Lk;
  const float scale2 = a.scale * kLog2e;                        // scores are handled in the log2 domain: one FFMA + one MUFU.EX2 each
  const uint32_t tm = atc_setup(sh, kCols);
  const uint32_t tS = tm, tO = tm + NKB * 64;
  const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
  DropCfg dc;
  if (DROP) dc = make_drop(a.drop);
  uint32_t parity = 0;
  SegIdx gq, gk, gv;                                            // group indices, fetched one item ahead
  if ((int)blockIdx.x < items) { const int p0 = blockIdx.x / a.heads; gq = seg_idx(a.q, p0); gk = seg_idx(a.k, p0); gv = seg_idx(a.v, p0); }

#pragma unroll 1
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int p = item / a.heads, h = item - p * a.heads;
    const RowSrc qsrc = row_src(a.q, gq, h);
    // every reader of the previous item's K/V/mask (its MMAs and pass-2 loops) has finished: see the waits below
    stage_rows(Ks, row_src(a.k, gk, h), 0, NKB * 64, Lk);
    stage_rows(Vs, row_src(a.v, gv, h), 0, NKB * 64, Lk);
    if (item + (int)gridDim.x < items) {
      const int pn = (item + (int)gridDim.x) / a.heads;
      gq = seg_idx(a.q, pn); gk = seg_idx(a.k, pn); gv = seg_idx(a.v, pn);
    }
    const float* madd = a.mask_add ? a.mask_add + (int64_t)(p / a.mask_div) * a.ld_mask : nullptr;
    for (int j = tid; j < NKB * 64; j += ATC_THREADS) msk[j] = j < Lk ? (madd ? madd[j] * kLog2e : 0.f) : -INFINITY;   // log2 domain

#pragma unroll 1
    for (int mt = 0; mt < mtiles; ++mt) {
      const int row0 = mt * ATC_TILE;
      const int row = row0 + trow;
      const bool wact = row0 + (warp & 3) * 32 < Lq;            // warp-uniform: any of this warp's 32 rows is a real query
      stage_rows(Qs, qsrc, row0, ATC_TILE, Lq);
      cp_async_commit();
      cp_async_wait_all();
      a_fence_async();
      a_tc_before();
      __syncthreads();
      a_tc_after();
      if (tid == 0) {
#pragma unroll
        for (int b = 0; b < NKB; ++b) block_mma_rr(tS + b * 64, a_smem_u32(Qs), a_smem_u32(Ks + b * ATC_BLK_BYTES), false);
        a_commit(&sh->bar);
      }
      a_mbar_wait(&sh->bar, parity); parity ^= 1;
      a_tc_after();

      float mx = -INFINITY;
      if (wact) {
#pragma unroll 1
        for (int c = grp; c < NKB * 2; c += 2) {
          uint32_t r[32];
          a_tmem_ld32(tS + lane_addr + c * 32, r);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 m4 = *reinterpret_cast<const float4*>(msk + c * 32 + j);
            const float2 a = __ffma2_rn(u2f2(r[j], r[j + 1]), f2(scale2, scale2), f2(m4.x, m4.y));
            const float2 b = __ffma2_rn(u2f2(r[j + 2], r[j + 3]), f2(scale2, scale2), f2(m4.z, m4.w));
            mx = fmaxf(fmaxf(mx, fmaxf(a.x, a.y)), fmaxf(b.x, b.y));
          }
        }
        red[grp * 128 + trow] = mx;
      }
      // the two warps that share a lane quarter (w, w + 4) exchange their row maxima: a 64-thread named barrier instead of a
      // CTA-wide one (nothing else is shared between the passes; idle pairs skip it together -- wact is uniform per pair)
      if (wact) asm volatile("bar.sync %0, 64;" ::"r"(1 + (warp & 3)) : "memory");
      float sum = 0.f;
      if (wact) {
        mx = fmaxf(red[trow], red[128 + trow]);                 // finite: at least key 0 is real and its mask is finite
        uint32_t rseed = 0;
        if (DROP) rseed = drop_rowseed(dc.seed, attn_drop_row(a, p, h, row));
        float2 sum2 = f2(0.f, 0.f);
#pragma unroll 1
        for (int c = grp; c < NKB * 2; c += 2) {
          uint32_t r[32];
          float v[32];
          a_tmem_ld32(tS + lane_addr + c * 32, r);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 m4 = *reinterpret_cast<const float4*>(msk + c * 32 + j);
            const float2 a = __fadd2_rn(__ffma2_rn(u2f2(r[j], r[j + 1]), f2(scale2, scale2), f2(m4.x, m4.y)), f2(-mx, -mx));
            const float2 b = __fadd2_rn(__ffma2_rn(u2f2(r[j + 2], r[j + 3]), f2(scale2, scale2), f2(m4.z, m4.w)), f2(-mx, -mx));
            v[j] = ex2_approx(a.x); v[j + 1] = ex2_approx(a.y);
            v[j + 2] = ex2_approx(b.x); v[j + 3] = ex2_approx(b.y);
            sum2 = __fadd2_rn(sum2, __fadd2_rn(f2(v[j], v[j + 1]), f2(v[j + 2], v[j + 3])));
          }
          if (DROP) drop_row32(v, rseed, (uint32_t)(c * 32), dc.thr16);    // the denominator keeps the dropped terms
          store_row32(Pblk(c >> 1), trow, (c & 1) * 32, v);
        }
        sum = sum2.x + sum2.y;
        red[256 + grp * 128 + trow] = sum;
      }
      a_fence_async();
      a_tc_before();
      __syncthreads();
      a_tc_after();
      if (tid == 0) {
        const int ksteps_total = (Lk + 15) >> 4;
#pragma unroll
        for (int b = 0; b < NKB; ++b) {
          const int ks = min(4, ksteps_total - b * 4);
          if (ks > 0) block_mma_rc(tO, a_smem_u32(Pblk(b)), a_smem_u32(Vs + b * ATC_BLK_BYTES), b > 0, ks);
        }
        a_commit(&sh->bar);
      }
      a_mbar_wait(&sh->bar, parity); parity ^= 1;              // P.V retired: Q/P, K, V and the mask may be overwritten
      a_tc_after();
      if (wact) {
        sum = red[256 + trow] + red[256 + 128 + trow];
        uint32_t r[32];
        a_tmem_ld32(tO + lane_addr + grp * 32, r);
        const float osc = (DROP ? dc.inv_keep : 1.0f) / sum;
        if (row < Lq) {
          store_out32(ctx + ((int64_t)p * Lq + row) * ldctx + (int64_t)h * 64 + grp * 32, r, osc);
          if (grp == 0 && lse) lse[((int64_t)p * a.heads + h) * Lq + row] = (mx + __log2f(sum)) * 0.69314718055994530942f;   // back to natural log
        }
      }
      a_tc_before();                                            // orders this tile's tcgen05.ld before the next tile's MMAs (next __syncthreads)
    }
  }
  atc_teardown(tm, kCols);
}

// ------------------------------------------------------------------------------------------- dQ (+ delta)
// item = (problem, head); inner loop over the 128-row query tiles, K/V/mask staged once per item.
// smem: [Q 16K][dO 16K][K NKB*8K][V NKB*8K][dS 16K][mask NKB*64 f32][dpart 256 f32][AtcShared]; TMEM: S 64 | dP 64 | dQ 64
template <int NKB, bool DROP>
__global__ void __launch_bounds__(ATC_THREADS, 2)
attn_tc_dq_kernel(AttnDev a, const bf16* __restrict__ ctx, int64_t ldctx, const bf16* __restrict__ dctx, int64_t lddctx,
                  const float* __restrict__ lse, bf16* __restrict__ dq, float* __restrict__ delta, int mtiles, int items) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = align1k(raw);
  uint8_t* Qs = sm;
  uint8_t* Gs = Qs + ATC_TILE_BYTES;
  uint8_t* Ks = Gs + ATC_TILE_BYTES;
  uint8_t* Vs = Ks + NKB * ATC_BLK_BYTES;
  uint8_t* Ds = Vs + NKB * ATC_BLK_BYTES;
  float* msk = reinterpret_cast<float*>(Ds + ATC_TILE_BYTES);
  float* dpart = msk + NKB * 64;                                 // [2 groups][128 rows] halves of delta
  AtcShared* sh = reinterpret_cast<AtcShared*>(dpart + 256);
  constexpr uint32_t kCols = 256;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = warp >> 2, trow = (warp & 3) * 32 + lane;
  const int Lq = a.Lq, Lk = a.Lk, HD = a.heads * 64;
  const float scale2 = a.scale * kLog2e;
  const int nblk = (Lk + 63) >> 6;
  const uint32_t tm = atc_setup(sh, kCols);
  const uint32_t tS = tm, tP = tm + 64, tQ = tm + 128;
  const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
  DropCfg dc;
  if (DROP) dc = make_drop(a.drop);
  uint32_t parity = 0;
  SegIdx gq, gk, gv;                                             // group indices, fetched one item ahead
  if ((int)blockIdx.x < items) { const int p0 = blockIdx.x / a.heads; gq = seg_idx(a.q, p0); gk = seg_idx(a.k, p0); gv = seg_idx(a.v, p0); }

#pragma unroll 1
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int p = item / a.heads, h = item - p * a.heads;
    const RowSrc qsrc = row_src(a.q, gq, h);
    stage_rows(Ks, row_src(a.k, gk, h), 0, NKB * 64, Lk);
    stage_rows(Vs, row_src(a.v, gv, h), 0, NKB * 64, Lk);
    if (item + (int)gridDim.x < items) {
      const int pn = (item + (int)gridDim.x) / a.heads;
      gq = seg_idx(a.q, pn); gk = seg_idx(a.k, pn); gv = seg_idx(a.v, pn);
    }
    const float* madd = a.mask_add ? a.mask_add + (int64_t)(p / a.mask_div) * a.ld_mask : nullptr;
    for (int j = tid; j < NKB * 64; j += ATC_THREADS) msk[j] = j < Lk ? (madd ? madd[j] * kLog2e : 0.f) : -INFINITY;   // log2 domain

#pragma unroll 1
    for (int mt = 0; mt < mtiles; ++mt) {
      const int row0 = mt * ATC_TILE;
      const int row = row0 + trow;
      const bool wact = row0 + (warp & 3) * 32 < Lq;
      stage_rows(Qs, qsrc, row0, ATC_TILE, Lq);
      stage_rows(Gs, row_src_plain(dctx, lddctx, (int64_t)p * Lq, h), row0, ATC_TILE, Lq);
      // the O tile rides in the dS buffer until the first block (its last reader, the previous tile's dQ MMA, retired
      // before that tile's epilogue): delta is then a shared-memory dot product instead of 8 strided global loads per
      // thread (ncu: 18 % of the kernel's stall samples sat on those loads)
      stage_rows(Ds, row_src_plain(ctx, ldctx, (int64_t)p * Lq, h), row0, ATC_TILE, Lq);
      cp_async_commit();
      const int64_t stat = ((int64_t)p * a.heads + h) * Lq + row;
      float l2 = 0.f;
      if (row < Lq) l2 = lse[stat] * kLog2e;                     // in flight while the tiles land
      cp_async_wait_all();
      a_fence_async();
      a_tc_before();
      __syncthreads();
      a_tc_after();
      {
        // delta_i = dO_i . O_i : two threads per row (one per warpgroup), 32 columns each; padded rows are zero-filled
        float part = 0.f;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint4 wo = *reinterpret_cast<const uint4*>(Ds + swz(trow, grp * 4 + g));
          const uint4 wg = *reinterpret_cast<const uint4*>(Gs + swz(trow, grp * 4 + g));
          const __nv_bfloat162* ho = reinterpret_cast<const __nv_bfloat162*>(&wo);
          const __nv_bfloat162* hg = reinterpret_cast<const __nv_bfloat162*>(&wg);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 fo = __bfloat1622float2(ho[j]), fg = __bfloat1622float2(hg[j]);
            part = fmaf(fo.x, fg.x, fmaf(fo.y, fg.y, part));
          }
        }
        dpart[grp * 128 + trow] = part;
      }
      __syncthreads();                                           // halves exchanged; every read of the O tile (dS buffer) is done
      const float dl = dpart[trow] + dpart[128 + trow];
      if (row < Lq && grp == 0) delta[stat] = dl;
      uint32_t rseed = 0;
      float keep_sc = 1.0f;
      if (DROP) { rseed = drop_rowseed(dc.seed, attn_drop_row(a, p, h, row)); keep_sc = dc.inv_keep; }

#pragma unroll 1
      for (int b = 0; b < nblk; ++b) {
        if (tid == 0) {
          block_mma_rr(tS, a_smem_u32(Qs), a_smem_u32(Ks + b * ATC_BLK_BYTES), false);      // S_b  = Q . K_b^T
          block_mma_rr(tP, a_smem_u32(Gs), a_smem_u32(Vs + b * ATC_BLK_BYTES), false);      // dP_b = dO . V_b^T
          a_commit(&sh->bar);
        }
        a_mbar_wait(&sh->bar, parity); parity ^= 1;       // also: the previous block's dQ MMA (reads dS) has retired
        a_tc_after();
        if (wact) {                                        // rows of idle warps feed only dQ rows that are never stored
          uint32_t rs[32], rp[32];
          float v[32];
          a_tmem_ld32(tS + lane_addr + grp * 32, rs);
          a_tmem_ld32(tP + lane_addr + grp * 32, rp);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 m4 = *reinterpret_cast<const float4*>(msk + b * 64 + grp * 32 + j);
            const float mm[4] = {m4.x, m4.y, m4.z, m4.w};
            uint32_t h0 = 0, h1 = 0;
            if (DROP) { h0 = drop_pair(rseed, (uint32_t)(b * 64 + grp * 32 + j)); h1 = drop_pair(rseed, (uint32_t)(b * 64 + grp * 32 + j + 2)); }
#pragma unroll
            for (int u = 0; u < 4; u += 2) {
              const float2 x = __fadd2_rn(__ffma2_rn(u2f2(rs[j + u], rs[j + u + 1]), f2(scale2, scale2), f2(mm[u], mm[u + 1])), f2(-l2, -l2));
              const float2 pj = f2(ex2_approx(x.x), ex2_approx(x.y));
              float2 dp = u2f2(rp[j + u], rp[j + u + 1]);
              if (DROP) {
                const uint32_t hh = u < 2 ? h0 : h1;
                dp = __fmul2_rn(dp, f2(drop_keep_lo(hh, dc.thr16) ? keep_sc : 0.f, drop_keep_hi(hh, dc.thr16) ? keep_sc : 0.f));   // dP = keep/(1-p) * (dO . v_j)
              }
              const float2 ds = __fmul2_rn(pj, __fadd2_rn(dp, f2(-dl, -dl)));
              v[j + u] = ds.x; v[j + u + 1] = ds.y;
            }
          }
          store_row32(Ds, trow, grp * 32, v);
        }
        a_fence_async();
        a_tc_before();
        __syncthreads();
        a_tc_after();
        if (tid == 0) {
          const int ks = min(4, ((Lk + 15) >> 4) - b * 4);
          block_mma_rc(tQ, a_smem_u32(Ds), a_smem_u32(Ks + b * ATC_BLK_BYTES), b > 0, ks);  // dQ += dS_b . K_b
          if (b == nblk - 1) a_commit(&sh->bar);
        }
      }
      a_mbar_wait(&sh->bar, parity); parity ^= 1;
      a_tc_after();
      if (wact) {
        uint32_t r[32];
        a_tmem_ld32(tQ + lane_addr + grp * 32, r);
        if (row < Lq) store_out32(dq + ((int64_t)p * Lq + row) * HD + (int64_t)h * 64 + grp * 32, r, a.scale);
      }
      a_tc_before();
    }
  }
  atc_teardown(tm, kCols);
}

// ------------------------------------------------------------------------------------------- dK, dV
// item = (problem, head, 128-row KEY tile); query blocks of 64 stream through a 2-stage ring (prefetched with cp.async).
// smem: [K 16K][V 16K][ring 2 x (Q 8K | dO 8K)][P^T 16K][dS^T 16K][lse NQB*64][delta NQB*64][row seeds NQB*64][AtcShared]
// TMEM: S^T 64 | dP^T 64 | dV 64 | dK 64
template <int NQB, bool DROP>
__global__ void __launch_bounds__(ATC_THREADS, 2)
attn_tc_dkv_kernel(AttnDev a, const bf16* __restrict__ dctx, int64_t lddctx, const float* __restrict__ lse,
                   const float* __restrict__ delta, bf16* __restrict__ dk, bf16* __restrict__ dv, int ktiles, int items) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = align1k(raw);
  uint8_t* Ks = sm;
  uint8_t* Vs = Ks + ATC_TILE_BYTES;
  uint8_t* ring = Vs + ATC_TILE_BYTES;                           // stage s: Q block at ring + s*16K, dO block 8K later
  uint8_t* Pt = ring + 2 * ATC_TILE_BYTES;
  uint8_t* St = Pt + ATC_TILE_BYTES;
  float* ls = reinterpret_cast<float*>(St + ATC_TILE_BYTES);
  float* dls = ls + NQB * 64;
  uint32_t* rsd = reinterpret_cast<uint32_t*>(dls + NQB * 64);   // dropout row seed of every query row
  AtcShared* sh = reinterpret_cast<AtcShared*>(rsd + NQB * 64);
  constexpr uint32_t kCols = 256;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = warp >> 2, trow = (warp & 3) * 32 + lane;
  const int Lq = a.Lq, Lk = a.Lk, HD = a.heads * 64;
  const int nblk = (Lq + 63) >> 6;
  const float scale2 = a.scale * kLog2e;
  const uint32_t tm = atc_setup(sh, kCols);
  const uint32_t tS = tm, tP = tm + 64, tV = tm + 128, tK = tm + 192;
  const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
  DropCfg dc;
  if (DROP) dc = make_drop(a.drop);
  uint32_t parity = 0;
  SegIdx gq, gk, gv;                                             // group indices, fetched one item ahead
  if ((int)blockIdx.x < items) { const int p0 = (blockIdx.x / ktiles) / a.heads; gq = seg_idx(a.q, p0); gk = seg_idx(a.k, p0); gv = seg_idx(a.v, p0); }

#pragma unroll 1
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int kt = item % ktiles;
    const int ph = item / ktiles;
    const int p = ph / a.heads, h = ph - p * a.heads;
    const RowSrc qsrc = row_src(a.q, gq, h);
    const RowSrc gsrc = row_src_plain(dctx, lddctx, (int64_t)p * Lq, h);
    const int key0 = kt * ATC_TILE;
    const int key = key0 + trow;
    const bool wact = key0 + (warp & 3) * 32 < Lk;               // warp-uniform: any of this warp's 32 rows is a real key
    stage_rows(Ks, row_src(a.k, gk, h), key0, ATC_TILE, Lk);
    stage_rows(Vs, row_src(a.v, gv, h), key0, ATC_TILE, Lk);
    stage_rows(ring, qsrc, 0, 64, Lq);
    stage_rows(ring + ATC_BLK_BYTES, gsrc, 0, 64, Lq);
    cp_async_commit();
    if (item + (int)gridDim.x < items) {
      const int pn = ((item + (int)gridDim.x) / ktiles) / a.heads;
      gq = seg_idx(a.q, pn); gk = seg_idx(a.k, pn); gv = seg_idx(a.v, pn);
    }
    const int64_t stat0 = ((int64_t)p * a.heads + h) * Lq;
    for (int i = tid; i < NQB * 64; i += ATC_THREADS) {
      ls[i] = i < Lq ? lse[stat0 + i] * kLog2e : INFINITY;      // log2 domain; exp2(s - inf) = 0 for the padded queries
      dls[i] = i < Lq ? delta[stat0 + i] : 0.f;
      if (DROP) rsd[i] = drop_rowseed(dc.seed, attn_drop_row(a, p, h, i));
    }
    const float* madd = a.mask_add ? a.mask_add + (int64_t)(p / a.mask_div) * a.ld_mask : nullptr;
    const float mk = (key < Lk && madd) ? madd[key] * kLog2e : 0.f;
    const uint32_t kpair = (uint32_t)key >> 1, kshift = ((uint32_t)key & 1u) * 16u;
    cp_async_wait_all();
    a_fence_async();
    a_tc_before();
    __syncthreads();
    a_tc_after();

#pragma unroll 1
    for (int b = 0; b < nblk; ++b) {
      uint8_t* Qb = ring + (b & 1) * ATC_TILE_BYTES;
      uint8_t* Gb = Qb + ATC_BLK_BYTES;
      if (tid == 0) {
        block_mma_rr(tS, a_smem_u32(Ks), a_smem_u32(Qb), false);                          // S^T_b  = K . Q_b^T
        block_mma_rr(tP, a_smem_u32(Vs), a_smem_u32(Gb), false);                          // dP^T_b = V . dO_b^T
        a_commit(&sh->bar);
      }
      a_mbar_wait(&sh->bar, parity); parity ^= 1;       // every earlier MMA has retired: the other ring stage and P^T/dS^T are free
      a_tc_after();
      if (b + 1 < nblk) {                                // prefetch the next query block while this one is processed
        uint8_t* Qn = ring + ((b + 1) & 1) * ATC_TILE_BYTES;
        stage_rows(Qn, qsrc, (b + 1) * 64, 64, Lq);
        stage_rows(Qn + ATC_BLK_BYTES, gsrc, (b + 1) * 64, 64, Lq);
      }
      cp_async_commit();
      if (wact) {                                        // rows of idle warps feed only dK/dV rows that are never stored
        uint32_t rs[32], rp[32];
        float pv[32], dsv[32];
        a_tmem_ld32(tS + lane_addr + grp * 32, rs);
        a_tmem_ld32(tP + lane_addr + grp * 32, rp);
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const int qi = b * 64 + grp * 32 + j;
          const float2 l2 = *reinterpret_cast<const float2*>(ls + qi), d2 = *reinterpret_cast<const float2*>(dls + qi);
          const float2 x = __fadd2_rn(__ffma2_rn(u2f2(rs[j], rs[j + 1]), f2(scale2, scale2), f2(mk, mk)), f2(-l2.x, -l2.y));
          const float2 pj = f2(ex2_approx(x.x), ex2_approx(x.y));                 // 0 for padded queries (ls = +inf)
          float2 pd = pj, dp = u2f2(rp[j], rp[j + 1]);
          if (DROP) {
            const bool k0 = ((mix32(rsd[qi] + kpair) >> kshift) & 0xffffu) >= dc.thr16;
            const bool k1 = ((mix32(rsd[qi + 1] + kpair) >> kshift) & 0xffffu) >= dc.thr16;
            const float2 ks = f2(k0 ? dc.inv_keep : 0.f, k1 ? dc.inv_keep : 0.f);
            pd = __fmul2_rn(pj, ks);                       // dropped probability: dV = P_drop^T . dO
            dp = __fmul2_rn(dp, ks);
          }
          const float2 ds = __fmul2_rn(pj, __fadd2_rn(dp, f2(-d2.x, -d2.y)));
          pv[j] = pd.x; pv[j + 1] = pd.y;
          dsv[j] = ds.x; dsv[j + 1] = ds.y;
        }
        store_row32(Pt, trow, grp * 32, pv);
        store_row32(St, trow, grp * 32, dsv);
      }
      cp_async_wait_all();
      a_fence_async();
      a_tc_before();
      __syncthreads();
      a_tc_after();
      if (tid == 0) {
        const int ks = min(4, ((Lq + 15) >> 4) - b * 4);
        block_mma_rc(tV, a_smem_u32(Pt), a_smem_u32(Gb), b > 0, ks);                      // dV += P^T_b . dO_b
        block_mma_rc(tK, a_smem_u32(St), a_smem_u32(Qb), b > 0, ks);                      // dK += dS^T_b . Q_b
        if (b == nblk - 1) a_commit(&sh->bar);
      }
    }
    a_mbar_wait(&sh->bar, parity); parity ^= 1;          // all MMAs retired: K, V, the ring and the statistics may be overwritten
    a_tc_after();
    if (wact) {
#pragma unroll 1
      for (int c = grp; c < 4; c += 2) {                 // chunks 0,1 = dV columns, 2,3 = dK columns
        uint32_t r[32];
        a_tmem_ld32((c < 2 ? tV : tK) + lane_addr + (c & 1) * 32, r);
        if (key < Lk)
          store_out32((c < 2 ? dv : dk) + ((int64_t)p * Lk + key) * HD + (int64_t)h * 64 + (c & 1) * 32, r, c < 2 ? 1.0f : a.scale);
      }
    }
    a_tc_before();
  }
  atc_teardown(tm, kCols);
}

// ------------------------------------------------------------------------------------------- host dispatch
static bool seg_ok(const SegDev& s) {
  if (!s.ptr || s.rows == 0) return true;
  return (reinterpret_cast<uintptr_t>(s.ptr) & 15u) == 0 && (s.ld % 8) == 0;
}

bool attn_tc_supported(const AttnDev& a, int64_t ldctx, const void* ctx) {
  if (a.dh != 64 || a.bias != nullptr || a.causal) return false;
  if (a.Lq < 16 || a.Lk < 1 || a.Lk > 320 || a.Lq > 320) return false;      // NKB, NQB <= 5
  if ((ldctx % 8) || (reinterpret_cast<uintptr_t>(ctx) & 15u)) return false;
  for (int s = 0; s < 2; ++s)
    if (!seg_ok(a.q[s]) || !seg_ok(a.k[s]) || !seg_ok(a.v[s])) return false;
  return (int64_t)a.NP * a.heads * ((a.Lq + 127) / 128) < (1LL << 31);
}

template <typename K>
static int set_smem_tc(K kernel, size_t bytes) {
  FCMF_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}
// persistent grid: every SM gets `per_sm` resident CTAs (fewer when there is less work)
static unsigned persist_grid(int64_t items, int per_sm) {
  const int64_t cap = (int64_t)sm_count() * per_sm;
  return (unsigned)(items < cap ? items : cap);
}

#define ATC_DISPATCH(nb, CALL) \
  switch (nb) { case 1: { CALL(1); } break; case 2: { CALL(2); } break; case 3: { CALL(3); } break; \
                case 4: { CALL(4); } break; default: { CALL(5); } break; }

int attn_tc_fwd(const AttnDev& a, void* ctx, int64_t ldctx, float* lse, cudaStream_t st) {
  const int nkb = (a.Lk + 63) / 64, mtiles = (a.Lq + 127) / 128;
  const int items = a.NP * a.heads;
  const bool drop = a.drop.p > 0.f;
#define LAUNCH(NB, DR)                                                                                        \
  if (int r = set_smem_tc(attn_tc_fwd_kernel<NB, DR>, smem)) return r;                                        \
  attn_tc_fwd_kernel<NB, DR><<<grid, ATC_THREADS, smem, st>>>(a, (bf16*)ctx, ldctx, lse, mtiles, items);
#define CALL(NB)                                                                                              \
  const size_t smem = 1024 + (size_t)NB * (2 * ATC_BLK_BYTES + ATC_TILE_BYTES + 256) + 2048 + 64;            \
  const unsigned grid = persist_grid(items, NB == 1 ? 3 : (NB <= 3 ? 2 : 1));   /* NB >= 4: 512 TMEM columns, one CTA per SM */                                                 \
  if (drop) { LAUNCH(NB, true) } else { LAUNCH(NB, false) }
  ATC_DISPATCH(nkb, CALL)
#undef CALL
#undef LAUNCH
  FCMF_LAUNCH_OK();
  return 0;
}

int attn_tc_bwd(const AttnDev& a, const void* ctx, int64_t ldctx, const void* dctx, int64_t lddctx, const float* lse,
                float* delta, void* dq, void* dk, void* dv, cudaStream_t st) {
  const int nkb = (a.Lk + 63) / 64, nqb = (a.Lq + 63) / 64;
  const int mtiles = (a.Lq + 127) / 128, ktiles = (a.Lk + 127) / 128;
  const bool drop = a.drop.p > 0.f;
  {
    const int items = a.NP * a.heads;
    const unsigned grid = persist_grid(items, 2);
#define LAUNCH(NB, DR)                                                                                        \
    if (int r = set_smem_tc(attn_tc_dq_kernel<NB, DR>, smem)) return r;                                       \
    attn_tc_dq_kernel<NB, DR><<<grid, ATC_THREADS, smem, st>>>(a, (const bf16*)ctx, ldctx, (const bf16*)dctx, lddctx, lse, (bf16*)dq, delta, mtiles, items);
#define CALL(NB)                                                                                              \
    const size_t smem = 1024 + 3 * ATC_TILE_BYTES + (size_t)NB * (2 * ATC_BLK_BYTES + 256) + 1024 + 64;      \
    if (drop) { LAUNCH(NB, true) } else { LAUNCH(NB, false) }
    ATC_DISPATCH(nkb, CALL)
#undef CALL
#undef LAUNCH
    FCMF_LAUNCH_OK();
  }
  {
    const int items = a.NP * a.heads * ktiles;
    const unsigned grid = persist_grid(items, 2);
#define LAUNCH(NB, DR)                                                                                        \
    if (int r = set_smem_tc(attn_tc_dkv_kernel<NB, DR>, smem)) return r;                                      \
    attn_tc_dkv_kernel<NB, DR><<<grid, ATC_THREADS, smem, st>>>(a, (const bf16*)dctx, lddctx, lse, delta, (bf16*)dk, (bf16*)dv, ktiles, items);
#define CALL(NB)                                                                                              \
    const size_t smem = 1024 + 6 * ATC_TILE_BYTES + (size_t)NB * 768 + 64;                                   \
    if (drop) { LAUNCH(NB, true) } else { LAUNCH(NB, false) }
    ATC_DISPATCH(nqb, CALL)
#undef CALL
#undef LAUNCH
    FCMF_LAUNCH_OK();
  }
  return 0;
}

}  // namespace fcmf
