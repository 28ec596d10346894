// Warp-specialised, TMA-fed tcgen05 / TMEM folded attention for head_dim 64, bf16 (forward + ONE backward kernel).
//
// Replaces the cp.async kernels of attn_tc.cu for the shapes of the fusion path (padded key length <= 192). Round-1
// profile of those kernels: 8-17 % of DRAM peak, issue slots 34-42 % busy, every MMA followed by a CTA-wide wait -- the
// loads, the MMAs and the softmax of one CTA never overlapped. Here one persistent CTA per SM runs three roles:
//   warp 0      producer: TMA (cp.async.bulk.tensor.3d, SWIZZLE_128B) of Q / K / V (/ dO / O) tiles into shared-memory
//               rings, one elected lane; the other lanes build the additive key mask of the item
//   warp 1      MMA issuer: one thread issues every tcgen05.mma, completion is signalled with tcgen05.commit -> mbarrier
//   warps 2..9  two softmax warpgroups that work on ALTERNATE tiles, each with its own TMEM columns and P buffers, so
//               the MMAs of one tile run under the exponentials of the other
// The index indirection of the folded problems (problem p -> group idx[p] of a [groups, rows, cols] tensor) is the third
// coordinate of a 3-D tensor map; the box is clipped at the group's row count (zero fill), which also pads the tiles.
// Row segments: an operand is up to two row segments (text rows + ROI rows); segment 1 is placed behind segment 0 at the
// next multiple of 8 rows (a 1024-byte boundary of the 128-byte-swizzled tile), so padded positions differ from logical
// rows by a constant gap (RowLay).
//
// Reference semantics: BertCoAttention/BertSelfAttention.forward (mm_modeling.py:193-266): scale before the mask add,
// additive -10000 mask on the keys, softmax over keys, dropout on the probabilities, context = P.V; backward = autograd
// of the same.
#pragma once
#include "common.cuh"
#include "attn.h"

#include <cuda.h>
#include <algorithm>
#include <mutex>

namespace fcmf {
namespace ws {

constexpr int THREADS = 320;
constexpr uint32_t TILE_B = 128 * 128;          // [128 rows x 64 bf16] swizzled tile
constexpr uint32_t BLK_B = 64 * 128;            // [64 rows x 64 bf16]
constexpr float kLog2e = 1.44269504088896340736f;
constexpr float kLn2 = 0.69314718055994530942f;

// ------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  const uint32_t hint_ns = 100000;             // sleep in hardware until the phase flips (see gemm_tc.cu)
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.b32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(s32(bar)), "r"(parity), "r"(hint_ns) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug becomes a trap (launch failure) instead of a hung GPU. FCMF_WS_NO_TRAP lifts the bound
// (compute-sanitizer slows the kernels by orders of magnitude).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try(bar, parity)) {
#ifndef FCMF_WS_NO_TRAP
    if ((++spins & 0x3f) == 0 && clock64() - t0 > 6000000000LL) {
      printf("fcmf attn_ws: mbarrier wait timed out (tag %d block %d thread %d parity %u)\n", tag, (int)blockIdx.x, (int)threadIdx.x, parity);
      __trap();
    }
#endif
  }
}
__device__ __forceinline__ void fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tma3(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(s32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(s32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma3_store(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(s32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void wg_bar(int w) { asm volatile("bar.sync %0, 128;" ::"r"(w + 1) : "memory"); }   // one softmax warpgroup
__device__ __forceinline__ void prefetch_map(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// SWIZZLE_128B shared-memory matrix descriptor (version 1): start address, leading / stride byte offsets in 16-byte units
__device__ __forceinline__ uint64_t sdesc(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;                      // 8-row groups are 1024 B apart
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: bf16 x bf16 -> fp32, M = 128, N, operand majors (bit 15: A is MN-major, bit 16: B)
__host__ __device__ constexpr uint32_t idesc(int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ uint32_t swz(int r, int chunk) { return (uint32_t)(r * 128 + ((chunk ^ (r & 7)) << 4)); }

// thread-owned row: 32 consecutive bf16 (cols c0..c0+31 of a 64-col block) of row r into a swizzled [128 x 64] image
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
// 16 consecutive bf16 (cols c0..c0+15, c0 a multiple of 16) of row r into a swizzled image
__device__ __forceinline__ void store_row16(uint8_t* tile, int r, int c0, const float (&v)[16]) {
#pragma unroll
  for (int g = 0; g < 2; ++g) {
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
    for (int j = 0; j < 4; ++j)
      hh[j] = __floats2bfloat162_rn(__uint_as_float(r[g * 8 + 2 * j]) * sc, __uint_as_float(r[g * 8 + 2 * j + 1]) * sc);
    *reinterpret_cast<uint4*>(o + g * 8) = w;
  }
}

// 32 fp32 TMEM values * scale -> 32 bf16 of row r (cols c0..c0+31) of a swizzled [128 x 64] staging image
__device__ __forceinline__ void stage_out32(uint8_t* tile, int r, int c0, const uint32_t (&v)[32], float sc) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint4 w;
    __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&w);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      hh[j] = __floats2bfloat162_rn(__uint_as_float(v[g * 8 + 2 * j]) * sc, __uint_as_float(v[g * 8 + 2 * j + 1]) * sc);
    *reinterpret_cast<uint4*>(tile + swz(r, (c0 >> 3) + g)) = w;
  }
}

// packed fp32x2 arithmetic (FFMA2 / FADD2 / FMUL2): the same fp32 operations, two per instruction
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 u2f2(uint32_t a, uint32_t b) { return make_float2(__uint_as_float(a), __uint_as_float(b)); }

// ------------------------------------------------------------------------------------------- padded row layout
// Segment 0 occupies padded positions [0, rows0), zero fill up to rows0p = round_up(rows0, 8); segment 1 occupies
// [rows0p, rows0p + rows1). `total` = one past the last real position.
struct RowLay {
  int rows0, rows0p, rows1, rows1p, total;
};
__device__ __forceinline__ int logical_row(const RowLay& L, int x) {
  if (x < L.rows0) return x;
  if (x >= L.rows0p && x < L.rows0p + L.rows1) return x - (L.rows0p - L.rows0);
  return -1;
}

struct Params {
  RowLay ql, kl;
  int heads, items, n_qt, Lq, Lk;
  const int32_t *qidx0, *qidx1, *kidx0, *kidx1, *vidx0, *vidx1;
  const float* mask_add; int64_t ld_mask; int mask_div;
  float scale;
  fcmf_dropout drop;
};

// ------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}
// The encode call is a DRIVER API: it needs a current context on the calling thread. A thread whose first CUDA work is one of
// our calls (autograd's backward thread when an attention backward is the first op it runs) has none yet -> error 201.
// One runtime call binds the primary context.
static inline void bind_context() {
  static thread_local bool bound = false;
  if (!bound) { cudaFree(nullptr); bound = true; }
}
// 3-D bf16 tensor [groups][rows][cols] (row stride ld elements, group stride gstride elements); box = [64 cols][box_rows][1]
static inline int make_map3(CUtensorMap* map, const void* ptr, int64_t cols, int64_t rows, int64_t groups, int64_t ld, int64_t gstride, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(FCMF_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  bind_context();
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)groups};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)gstride * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(FCMF_ERR_CUDA, "cuTensorMapEncodeTiled (3-D) failed (%d) cols=%lld rows=%lld groups=%lld ld=%lld gstride=%lld box_rows=%d",
                (int)r, (long long)cols, (long long)rows, (long long)groups, (long long)ld, (long long)gstride, box_rows);
  return 0;
}

static inline int up8(int x) { return (x + 7) & ~7; }
static inline RowLay row_lay(const SegDev (&s)[2]) {
  RowLay L;
  L.rows0 = s[0].rows; L.rows0p = up8(s[0].rows);
  L.rows1 = s[1].rows; L.rows1p = up8(s[1].rows);
  L.total = L.rows1 ? L.rows0p + L.rows1 : L.rows0;
  return L;
}
static inline bool seg_ok(const SegDev& s) {
  if (!s.ptr || s.rows == 0) return true;
  return (reinterpret_cast<uintptr_t>(s.ptr) & 15u) == 0 && (s.ld % 8) == 0 && s.groups > 0 && s.rows <= 256;
}
static inline bool tiles_ok(const RowLay& L, int tile) {          // segment 1 must not straddle a tile boundary
  if (L.rows1 == 0) return true;
  return (L.rows0p % tile) + L.rows1p <= tile;
}
static inline int key_blocks(const RowLay& L) { return ((L.rows1 ? L.rows0p + L.rows1p : L.rows0p) + 63) / 64; }

static inline void fill_params(const AttnDev& a, Params* P) {
  P->ql = row_lay(a.q); P->kl = row_lay(a.k);
  P->heads = a.heads; P->items = a.NP * a.heads; P->Lq = a.Lq; P->Lk = a.Lk;
  P->n_qt = (P->ql.total + 127) / 128;
  P->qidx0 = a.q[0].idx; P->qidx1 = a.q[1].idx; P->kidx0 = a.k[0].idx; P->kidx1 = a.k[1].idx;
  P->vidx0 = a.v[0].idx; P->vidx1 = a.v[1].idx;
  P->mask_add = a.mask_add; P->ld_mask = a.ld_mask; P->mask_div = a.mask_div;
  P->scale = a.scale; P->drop = a.drop;
}
// maps of one operand role: segment 0 with a full-tile box (`tile` rows) and a tail box, segment 1
static inline int role_maps(const SegDev (&s)[2], const RowLay& L, int heads, int tile, CUtensorMap* full, CUtensorMap* tail, CUtensorMap* seg1) {
  const int64_t cols = (int64_t)heads * 64;
  const int tail_rows = L.rows0p % tile;
  if (full) { if (int r = make_map3(full, s[0].ptr, cols, s[0].rows, s[0].groups, s[0].ld, (int64_t)s[0].rows * s[0].ld, tile)) return r; }
  if (tail) { if (int r = make_map3(tail, s[0].ptr, cols, s[0].rows, s[0].groups, s[0].ld, (int64_t)s[0].rows * s[0].ld, tail_rows ? tail_rows : 8)) return r; }
  if (seg1) {
    if (L.rows1) { if (int r = make_map3(seg1, s[1].ptr, cols, s[1].rows, s[1].groups, s[1].ld, (int64_t)s[1].rows * s[1].ld, L.rows1p)) return r; }
    else *seg1 = *(full ? full : tail);
  }
  return 0;
}
}  // namespace ws
}  // namespace fcmf
