// tcgen05 / TMEM / TMA GEMM engine for sm_100a (bf16 operands, fp32 accumulation in tensor memory).
//
//   * persistent, warp-specialised CTA: warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread) and TMEM
//     owner, warps 2..9 = epilogue (two warps per TMEM lane quarter);
//   * operands staged by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) into a shared-memory ring and consumed by
//     tcgen05.mma.kind::f16 straight from shared memory through matrix descriptors;
//   * accumulators live in TMEM, double-buffered: the epilogue of tile i overlaps the main loop of tile i+1;
//   * bf16 epilogue: tcgen05.ld -> bias / erf-GELU / tanh / dGELU in registers -> 128-byte-swizzled shared-memory
//     slab of 128 rows x 64 columns -> ONE TMA store per slab (cp.async.bulk.tensor store). A thread owns a TMEM lane,
//     i.e. an output ROW, so direct global stores would issue 16-byte requests to 32 different rows per instruction
//     (measured: the L2 request rate, not DRAM, bounded the N=3072/K=768 GEMMs at 430 TFLOP/s); the TMA store moves
//     whole 128-byte lines and also clips the M/N tails. The dGELU epilogue's auxiliary input comes in the same way
//     (TMA load of the matching slab);
//   * both operand majors: K-major (forward and input-gradient GEMMs) and MN-major (weight-gradient GEMM, whose
//     reduction runs over the rows of two row-major activation matrices), split along the reduction with fp32
//     atomics for the weight gradient;
//   * two kernels: a CTA-pair kernel (cta_group::2, 256 x 256 tile per pair, half the L2->SM operand traffic per
//     FLOP) for large problems and a single-CTA kernel (128 x 256 / 128 x 128) for everything else.
//
// Descriptor encodings follow the PTX ISA "tcgen05 matrix descriptor" / "instruction descriptor" tables
// (same bit layout as cute::UMMA::SmemDescriptor / InstrDescriptor).
#include "common.cuh"
#include "gemm.h"

#include <cuda.h>
#include <mutex>
#include <stdlib.h>

namespace fcmf {

// ------------------------------------------------------------------------------------------- tile configuration
constexpr int TC_BM = 128;
constexpr int TC_BK = 64;                       // 64 bf16 = 128 bytes = one SWIZZLE_128B row
constexpr int TC_UMMA_K = 16;
constexpr int TC_THREADS = 320;                 // 10 warps: TMA, MMA, 8 epilogue
constexpr int TC_EPI_WARP0 = 2;                 // warps 2..9: two per TMEM lane quarter, one 32-column half of a slab each
constexpr int TC_SLAB_BYTES = TC_BM * 128;      // 128 rows x 64 bf16 columns
constexpr int TC_EPI_SMEM = 4 * TC_SLAB_BYTES;  // 2 buffers x (D slab + aux slab)
constexpr int TC_CTRL_BYTES = 512;              // barriers + TMEM slot

template <int BN> struct TcCfg {
  static constexpr int kStageA = TC_BM * TC_BK * 2;                 // 16 KB
  static constexpr int kStageB = BN * TC_BK * 2;                    // 16 / 32 KB
  static constexpr int kStageBytes = kStageA + kStageB;
  static constexpr int kStages = BN == 256 ? 3 : 4;                 // 144 / 128 KB of operand ring
  static constexpr int kTmemCols = 2 * BN;                          // double-buffered accumulator (256 / 512)
  static constexpr int kSmemBytes = kStages * kStageBytes + TC_EPI_SMEM + 1024 /*align*/ + TC_CTRL_BYTES;
};

// ------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait with a suspend-time hint: the thread SLEEPS in hardware until the phase completes (or the hint expires)
// instead of spinning. Without the hint the default time limit is so short that the producer / MMA / epilogue waits
// became hot loops: ncu attributed 62 % of all executed instructions of the GELU GEMM to this wait, issue slots stolen
// from the epilogue warps that share the scheduler (measured: GELU GEMM 2.21 -> 1.91 ms, dGELU 2.47 -> 2.14 ms).
constexpr uint32_t kMbarSuspendNs = 100000;
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(kMbarSuspendNs) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug becomes a trap (launch failure) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3f) == 0 && clock64() - t0 > 6000000000LL) {    // ~3-4 s
      printf("fcmf gemm_tc: mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // the 8 epilogue warps

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c_inner, int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c_inner, int c_outer) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c_inner), "r"(c_outer) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_out, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_out)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// mbarrier arrives once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (SWIZZLE_128B, version 1). Offsets are in bytes; >>4 encodes 16-byte units.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);                 // [0,14)  start address
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;        // [16,30) leading byte offset
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;        // [32,46) stride byte offset
  d |= (uint64_t)1 << 46;                                  // [46,48) descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                                  // [61,64) SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: bf16 x bf16 -> fp32, M x N tile, operand majors.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) /*D=f32*/ | (1u << 7) /*A=bf16*/ | (1u << 10) /*B=bf16*/ | ((uint32_t)a_mn_major << 15) |
         ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------- parameters
struct TcParams {
  // logical GEMM: D[Mg, Ng] = sum_k A(m,k) B(n,k), k < Kg
  int64_t Mg, Ng, Kg;
  int m_tiles, n_tiles, k_blocks, splits;      // work unit = (tile, split)
  const float* bias;
  float* Df; int64_t lddf;                     // fp32 output (weight gradient)
  int epi;
  int f32_mode;                                // 0 = bf16 epilogue through TMA stores, 1 = fp32 store, 2 = fp32 atomic add
  int has_aux;                                 // bf16 mode: aux tensor map is valid (GELU: output, dGELU: input)
};

struct EpiCtx {                                // per-thread view of the epilogue's shared resources
  uint8_t* stage;                              // 4 slabs: [buf][D | aux]
  uint64_t* xbar;                              // [2] aux-slab-landed barriers (dGELU)
  uint32_t slab_it;                            // slabs processed so far by this CTA (buffer = slab_it & 1)
  int quarter, half, lane;
  bool store_thread;                           // the one thread that issues TMA stores / aux loads
};

__device__ __forceinline__ uint32_t slab_off(int row, int chunk) {          // 128B-swizzled [128 x 64] bf16 slab
  return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4));
}

// fp32 output of one thread's 32 accumulator columns (weight gradient: plain store or split-K atomics)
__device__ __forceinline__ void epilogue_f32(const TcParams& P, int64_t m, int64_t n0, const uint32_t (&r)[32]) {
  float* dp = P.Df + m * P.lddf + n0;
  const int ncols = (int)min((int64_t)32, P.Ng - n0);
  if (P.f32_mode == 1) {
    if (ncols == 32 && (P.lddf & 3) == 0) {                      // 8 x 16-byte stores (the thread's 128 contiguous bytes)
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(dp + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) if (j < ncols) dp[j] = __uint_as_float(r[j]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) if (j < ncols) atomicAdd(dp + j, __uint_as_float(r[j]));
  }
}

// Epilogue of one accumulator tile [128 rows x BN columns] held in TMEM at `taddr_tile` (lane bits not yet added).
// Called by all 8 epilogue warps. (m0, n_tile0) = global coordinates of the tile's first row / column.
template <int BN>
__device__ __forceinline__ void epilogue_tile(const TcParams& P, const CUtensorMap* tmD, const CUtensorMap* tmX, EpiCtx& E,
                                              uint32_t taddr_tile, int m0, int64_t n_tile0, bool has_work) {
  const uint32_t taddr = taddr_tile + ((uint32_t)(E.quarter * 32) << 16);
  const int row = E.quarter * 32 + E.lane;                      // row inside the CTA tile == TMEM lane
  if (P.f32_mode != 0) {
    const int64_t m = (int64_t)m0 + row;
#pragma unroll 1
    for (int c = E.half; c < BN / 32; c += 2) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(taddr + c * 32, r);
      tmem_ld_wait();
      const int64_t n0 = n_tile0 + c * 32;
      if (m < P.Mg && n0 < P.Ng && has_work) epilogue_f32(P, m, n0, r);
    }
    return;
  }
#pragma unroll 1
  for (int s = 0; s < BN / 64; ++s) {
    const int64_t n_slab = n_tile0 + s * 64;
    if (n_slab >= P.Ng) break;                                  // uniform across the CTA
    const uint32_t buf = E.slab_it & 1;
    uint8_t* sD = E.stage + buf * (2 * TC_SLAB_BYTES);
    uint8_t* sX = sD + TC_SLAB_BYTES;
    if (E.store_thread) tma_store_wait_read<1>();               // the stores that read this buffer (2 slabs ago) are done
    epi_bar_sync();                                             // (A) staging buffer is free
    if (E.store_thread && P.epi == FCMF_EPI_DGELU && s + 1 < BN / 64 && n_slab + 64 < P.Ng) {
      // aux slab of the NEXT column block, one slab ahead (this slab's was issued a slab ago / before the accumulator
      // wait): the HBM latency of the load is no longer on the epilogue's critical path. Its buffer's last readers
      // (slab s-1) are behind barrier (A).
      const uint32_t nb = buf ^ 1;
      mbar_expect_tx(&E.xbar[nb], TC_SLAB_BYTES);
      tma_load_2d(E.stage + nb * (2 * TC_SLAB_BYTES) + TC_SLAB_BYTES, tmX, &E.xbar[nb], (int)(n_slab + 64), m0);
    }
    uint32_t r[32];
    tmem_ld_32x32b_x32(taddr + s * 64 + E.half * 32, r);
    tmem_ld_wait();
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    const int64_t n0 = n_slab + E.half * 32;
    if (P.bias) {
      if (n0 + 32 <= P.Ng) {                                    // whole chunk in range: 8 x 16-byte loads (n0 % 32 == 0)
        const float4* b4 = reinterpret_cast<const float4*>(P.bias + n0);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 t = __ldg(b4 + q);
          const float2 lo = __fadd2_rn(make_float2(v[4 * q], v[4 * q + 1]), make_float2(t.x, t.y));
          const float2 hi = __fadd2_rn(make_float2(v[4 * q + 2], v[4 * q + 3]), make_float2(t.z, t.w));
          v[4 * q] = lo.x; v[4 * q + 1] = lo.y; v[4 * q + 2] = hi.x; v[4 * q + 3] = hi.y;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) if (n0 + j < P.Ng) v[j] += __ldg(P.bias + n0 + j);
      }
    }
    if (P.epi == FCMF_EPI_GELU) {
      if (P.has_aux) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 w; __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&w);
#pragma unroll
          for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[g * 8 + 2 * j], v[g * 8 + 2 * j + 1]);
          *reinterpret_cast<uint4*>(sX + slab_off(row, E.half * 4 + g)) = w;
        }
      }
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float2 gq = gelu_erf_fast2(make_float2(v[j], v[j + 1]));
        v[j] = gq.x; v[j + 1] = gq.y;
      }
    } else if (P.epi == FCMF_EPI_TANH) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = tanh_approx(v[j]);
    } else if (P.epi == FCMF_EPI_DGELU) {
      mbar_wait(&E.xbar[buf], (E.slab_it >> 1) & 1);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint4 w = *reinterpret_cast<const uint4*>(sX + slab_off(row, E.half * 4 + g));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&w);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 d = __fmul2_rn(make_float2(v[g * 8 + 2 * j], v[g * 8 + 2 * j + 1]), gelu_erf_grad_fast2(__bfloat1622float2(h[j])));
          v[g * 8 + 2 * j] = d.x; v[g * 8 + 2 * j + 1] = d.y;
        }
      }
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      uint4 w; __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&w);
#pragma unroll
      for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[g * 8 + 2 * j], v[g * 8 + 2 * j + 1]);
      *reinterpret_cast<uint4*>(sD + slab_off(row, E.half * 4 + g)) = w;
    }
    fence_proxy_async();                                        // generic-proxy writes -> visible to the TMA store
    epi_bar_sync();                                             // (B) slab complete
    if (E.store_thread) {
      tma_store_2d(tmD, sD, (int)n_slab, m0);                   // clipped at the M / N edges by the tensor map
      if (P.epi == FCMF_EPI_GELU && P.has_aux) tma_store_2d(tmX, sX, (int)n_slab, m0);
      tma_store_commit();
    }
    ++E.slab_it;
  }
}

// dGELU: request the aux slab of the tile's first column block BEFORE waiting for the accumulator (called by all
// epilogue warps, one thread acts). The buffer (slab_it & 1) was last read two slabs ago.
__device__ __forceinline__ void epilogue_prefetch_aux(const TcParams& P, const CUtensorMap* tmX, EpiCtx& E, int m0, int64_t n_tile0) {
  if (P.f32_mode == 0 && P.epi == FCMF_EPI_DGELU && E.store_thread && n_tile0 < P.Ng) {
    const uint32_t buf = E.slab_it & 1;
    mbar_expect_tx(&E.xbar[buf], TC_SLAB_BYTES);
    tma_load_2d(E.stage + buf * (2 * TC_SLAB_BYTES) + TC_SLAB_BYTES, tmX, &E.xbar[buf], (int)n_tile0, m0);
  }
}

// ------------------------------------------------------------------------------------------- single-CTA kernel
template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmX, const TcParams P) {
  using Cfg = TcCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // base + offset keeps the shared address space (LDS/STS, not generic LD/ST)
  uint8_t* ring = smem;
  uint8_t* stage_epi = smem + Cfg::kStages * Cfg::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stage_epi + TC_EPI_SMEM);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* tfull_bar = empty_bar + 8;                     // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                    // [2]
  uint64_t* x_bar = tempty_bar + 2;                        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(x_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (P.f32_mode == 0) { tma_prefetch_desc(&tmD); if (P.has_aux) tma_prefetch_desc(&tmX); }
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 8); mbar_init(&x_bar[b], 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t units = (int64_t)P.m_tiles * P.n_tiles * P.splits;
  const int kb_per_split = (P.k_blocks + P.splits - 1) / P.splits;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int64_t u = blockIdx.x; u < units; u += gridDim.x) {
        const int split = (int)(u % P.splits);
        const int64_t tile = u / P.splits;
        const int n_blk = (int)(tile % P.n_tiles), m_blk = (int)(tile / P.n_tiles);
        const int kb0 = split * kb_per_split;
        const int kb1 = min(P.k_blocks, kb0 + kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = ring + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kStageA;
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          if (!A_MN) {
            tma_load_2d(sa, &tmA, &full_bar[stage], kb * TC_BK, m_blk * TC_BM);          // box [64 k][128 rows]
          } else {
#pragma unroll
            for (int c = 0; c < TC_BM / 64; ++c)                                           // box [64 m][64 k-rows]
              tma_load_2d(sa + c * (TC_BK * 128), &tmA, &full_bar[stage], m_blk * TC_BM + c * 64, kb * TC_BK);
          }
          if (!B_MN) {
            tma_load_2d(sb, &tmB, &full_bar[stage], kb * TC_BK, n_blk * BN);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              tma_load_2d(sb + c * (TC_BK * 128), &tmB, &full_bar[stage], n_blk * BN + c * 64, kb * TC_BK);
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(TC_BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      // K-major: 8-row groups are 1024 B apart (SBO); LBO unused (encoded 1).  MN-major: 64-element column
      // blocks are BK*128 B apart (LBO), 8-k-row groups 1024 B apart (SBO).
      constexpr uint32_t a_lbo = A_MN ? TC_BK * 128 : 16, b_lbo = B_MN ? TC_BK * 128 : 16;
      constexpr uint32_t a_kstep = A_MN ? TC_UMMA_K * 128 : TC_UMMA_K * 2;
      constexpr uint32_t b_kstep = B_MN ? TC_UMMA_K * 128 : TC_UMMA_K * 2;
      int stage = 0; uint32_t phase = 0;
      uint32_t it = 0;
      for (int64_t u = blockIdx.x; u < units; u += gridDim.x, ++it) {
        const int split = (int)(u % P.splits);
        const int kb0 = split * kb_per_split;
        const int kb1 = min(P.k_blocks, kb0 + kb_per_split);
        const uint32_t buf = it & 1, use = it >> 1;
        mbar_wait(&tempty_bar[buf], (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(ring + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kStageA;
#pragma unroll
          for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
            const uint64_t ad = make_smem_desc(sa + k * a_kstep, a_lbo, 1024);
            const uint64_t bd = make_smem_desc(sb + k * b_kstep, b_lbo, 1024);
            umma_f16(tmem_d, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);                  // frees the smem slot when these MMAs retire
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[buf]);                      // accumulator complete
      }
    }
  } else {
    // ===================================================================== epilogue (8 warps)
    EpiCtx E;
    E.stage = stage_epi; E.xbar = x_bar; E.slab_it = 0;
    E.quarter = warp & 3; E.half = (warp - TC_EPI_WARP0) >> 2; E.lane = lane;
    E.store_thread = (warp == TC_EPI_WARP0 && lane == 0);
    uint32_t it = 0;
    for (int64_t u = blockIdx.x; u < units; u += gridDim.x, ++it) {
      const int split = (int)(u % P.splits);
      const int64_t tile = u / P.splits;
      const int n_blk = (int)(tile % P.n_tiles), m_blk = (int)(tile / P.n_tiles);
      const bool has_work = split * kb_per_split < P.k_blocks;   // an empty split contributes nothing
      const uint32_t buf = it & 1, use = it >> 1;
      epilogue_prefetch_aux(P, &tmX, E, m_blk * TC_BM, (int64_t)n_blk * BN);
      mbar_wait(&tfull_bar[buf], use & 1);
      tc_fence_after();
      epilogue_tile<BN>(P, &tmD, &tmX, E, tmem_base + buf * BN, m_blk * TC_BM, (int64_t)n_blk * BN, has_work);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);
    }
    if (E.store_thread) tma_store_wait_all();               // shared memory must outlive the bulk stores
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, Cfg::kTmemCols); }
}

// ------------------------------------------------------------------------------------------- 2-CTA (cta_group::2) kernel
// A CTA pair (cluster of 2 on one TPC) owns a 256 x 256 output tile: each CTA stages its own 128 rows of A and its
// own 128-row half of B (32 KB per stage instead of 48 KB for the same MMA work), the leader CTA issues
// tcgen05.mma.cta_group::2 (M = 256) which reads both CTAs' shared memory and writes 128 accumulator lanes into each
// CTA's TMEM; each CTA runs the epilogue of its own 128 rows.
constexpr int TC2_BN = 256;
constexpr int TC2_STAGE_A = TC_BM * TC_BK * 2;        // 16 KB: this CTA's 128 rows of A
constexpr int TC2_STAGE_B = 128 * TC_BK * 2;          // 16 KB: this CTA's half of the 256 B rows
constexpr int TC2_STAGE = TC2_STAGE_A + TC2_STAGE_B;
constexpr int TC2_STAGES = 4;
constexpr int TC2_SMEM = TC2_STAGES * TC2_STAGE + TC_EPI_SMEM + 1024 + TC_CTRL_BYTES;
constexpr uint32_t TC2_PEER_MASK = 0xFEFFFFFFu;       // shared::cluster address of the same offset in the even (leader) CTA

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c_inner, int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & TC2_PEER_MASK), "r"(c_inner), "r"(c_outer)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_out, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_out)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  const uint32_t z = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {      // arrives on `bar` at the same offset in BOTH CTAs
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar, uint32_t cta) {   // arrive on `bar` of CTA `cta` of the cluster
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}

template <bool A_MN, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmX, const TcParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // base + offset keeps the shared address space (LDS/STS, not generic LD/ST)
  uint8_t* ring = smem;
  uint8_t* stage_epi = smem + TC2_STAGES * TC2_STAGE;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stage_epi + TC_EPI_SMEM);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* tfull_bar = empty_bar + 8;                    // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                   // [2], only the leader copy is used (16 arrivals: 8 warps x 2 CTAs)
  uint64_t* x_bar = tempty_bar + 2;                       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(x_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (P.f32_mode == 0) { tma_prefetch_desc(&tmD); if (P.has_aux) tma_prefetch_desc(&tmX); }
    for (int s = 0; s < TC2_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 16); mbar_init(&x_bar[b], 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t units = (int64_t)P.m_tiles * P.n_tiles * P.splits;      // tiles are 256 x 256 here
  const int kb_per_split = (P.k_blocks + P.splits - 1) / P.splits;
  const int64_t cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int64_t u = cluster_id; u < units; u += n_clusters) {
        const int split = (int)(u % P.splits);
        const int64_t tile = u / P.splits;
        const int n_blk = (int)(tile % P.n_tiles), m_blk = (int)(tile / P.n_tiles);
        const int kb0 = split * kb_per_split;
        const int kb1 = min(P.k_blocks, kb0 + kb_per_split);
        const int m0 = m_blk * 256 + (int)rank * 128, n0 = n_blk * 256 + (int)rank * 128;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = ring + stage * TC2_STAGE;
          uint8_t* sb = sa + TC2_STAGE_A;
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * TC2_STAGE);      // bytes of BOTH CTAs land on the leader's barrier
          if (!A_MN) {
            tma_load_2d_2sm(sa, &tmA, &full_bar[stage], kb * TC_BK, m0);
          } else {
#pragma unroll
            for (int c = 0; c < 2; ++c) tma_load_2d_2sm(sa + c * (TC_BK * 128), &tmA, &full_bar[stage], m0 + c * 64, kb * TC_BK);
          }
          if (!B_MN) {
            tma_load_2d_2sm(sb, &tmB, &full_bar[stage], kb * TC_BK, n0);
          } else {
#pragma unroll
            for (int c = 0; c < 2; ++c) tma_load_2d_2sm(sb + c * (TC_BK * 128), &tmB, &full_bar[stage], n0 + c * 64, kb * TC_BK);
          }
          if (++stage == TC2_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = make_idesc(256, TC2_BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      constexpr uint32_t a_lbo = A_MN ? TC_BK * 128 : 16, b_lbo = B_MN ? TC_BK * 128 : 16;
      constexpr uint32_t a_kstep = A_MN ? TC_UMMA_K * 128 : TC_UMMA_K * 2;
      constexpr uint32_t b_kstep = B_MN ? TC_UMMA_K * 128 : TC_UMMA_K * 2;
      int stage = 0; uint32_t phase = 0;
      uint32_t it = 0;
      for (int64_t u = cluster_id; u < units; u += n_clusters, ++it) {
        const int split = (int)(u % P.splits);
        const int kb0 = split * kb_per_split;
        const int kb1 = min(P.k_blocks, kb0 + kb_per_split);
        const uint32_t buf = it & 1, use = it >> 1;
        mbar_wait(&tempty_bar[buf], (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * TC2_BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(ring + stage * TC2_STAGE);
          const uint32_t sb = sa + TC2_STAGE_A;
#pragma unroll
          for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
            const uint64_t ad = make_smem_desc(sa + k * a_kstep, a_lbo, 1024);
            const uint64_t bd = make_smem_desc(sb + k * b_kstep, b_lbo, 1024);
            umma_f16_2sm(tmem_d, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit_2sm(&empty_bar[stage]);
          if (++stage == TC2_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm(&tfull_bar[buf]);
      }
    }
  } else {
    EpiCtx E;
    E.stage = stage_epi; E.xbar = x_bar; E.slab_it = 0;
    E.quarter = warp & 3; E.half = (warp - TC_EPI_WARP0) >> 2; E.lane = lane;
    E.store_thread = (warp == TC_EPI_WARP0 && lane == 0);
    uint32_t it = 0;
    for (int64_t u = cluster_id; u < units; u += n_clusters, ++it) {
      const int split = (int)(u % P.splits);
      const int64_t tile = u / P.splits;
      const int n_blk = (int)(tile % P.n_tiles), m_blk = (int)(tile / P.n_tiles);
      const bool has_work = split * kb_per_split < P.k_blocks;
      const uint32_t buf = it & 1, use = it >> 1;
      epilogue_prefetch_aux(P, &tmX, E, m_blk * 256 + (int)rank * 128, (int64_t)n_blk * TC2_BN);
      mbar_wait(&tfull_bar[buf], use & 1);
      tc_fence_after();
      epilogue_tile<TC2_BN>(P, &tmD, &tmX, E, tmem_base + buf * TC2_BN, m_blk * 256 + (int)rank * 128,
                            (int64_t)n_blk * TC2_BN, has_work);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cta(&tempty_bar[buf], 0);
    }
    if (E.store_thread) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                     // the peer's smem/TMEM/barriers stay valid until both are done
  if (warp == 1) { tc_fence_after(); tmem_dealloc_2sm(tmem_base, 512); }
}

// ------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 2-D bf16 row-major [rows, cols] with row stride ld (elements); box = [box_cols (inner), box_rows].
static int make_map(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(FCMF_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  {   // driver API: the calling thread needs a current context (autograd's backward thread may not have one yet)
    static thread_local bool bound = false;
    if (!bound) { cudaFree(nullptr); bound = true; }
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(FCMF_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld box=%dx%d", (int)r,
                (long long)rows, (long long)cols, (long long)ld, box_cols, box_rows);
  return 0;
}

static bool ok16(const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

bool gemm_tc_supported_tn(int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldd, int64_t ldaux,
                          const void* A, const void* B, const void* D, const void* aux) {
  if (M <= 0 || N <= 0 || K <= 0) return false;
  if (M >= (1LL << 31) || N >= (1LL << 31) || K >= (1LL << 31)) return false;
  if ((N % 8) || (K % 8) || (lda % 8) || (ldb % 8) || (ldd % 8) || (aux && (ldaux % 8))) return false;
  return ok16(A) && ok16(B) && ok16(D) && ok16(aux);
}

bool gemm_tc_supported_wgrad(int64_t M, int64_t N, int64_t K, int64_t lddy, int64_t ldx, const void* dY, const void* X) {
  if (M <= 0 || N <= 0 || K <= 0) return false;
  if (M >= (1LL << 31) || N >= (1LL << 31) || K >= (1LL << 31)) return false;
  if ((N % 8) || (K % 8) || (lddy % 8) || (ldx % 8)) return false;
  return ok16(dY) && ok16(X);
}

struct Maps { CUtensorMap a, b, d, x; };

template <int BN, bool A_MN, bool B_MN>
static int launch(const Maps& m, const TcParams& P, cudaStream_t st) {
  using Cfg = TcCfg<BN>;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN>;
  static bool attr_set[64] = {false};                       // per instantiation, per device
  int dev = 0;
  FCMF_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    FCMF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const int64_t units = (int64_t)P.m_tiles * P.n_tiles * P.splits;
  const int grid = (int)(units < sm_count() ? units : sm_count());
  kern<<<grid, TC_THREADS, Cfg::kSmemBytes, st>>>(m.a, m.b, m.d, m.x, P);
  FCMF_LAUNCH_OK();
  return 0;
}

// FCMF_GEMM_2CTA=0 disables the cta_group::2 kernel (debug / A-B measurements)
static bool use_2cta() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("FCMF_GEMM_2CTA"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

template <bool A_MN, bool B_MN>
static int launch2(const Maps& m, const TcParams& P, cudaStream_t st) {
  auto kern = gemm_tc2_kernel<A_MN, B_MN>;
  static bool attr_set[64] = {false};
  int dev = 0;
  FCMF_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    FCMF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TC2_SMEM));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const int64_t units = (int64_t)P.m_tiles * P.n_tiles * P.splits;
  const int64_t pairs = sm_count() / 2;
  const int grid = 2 * (int)(units < pairs ? units : pairs);
  kern<<<grid, TC_THREADS, TC2_SMEM, st>>>(m.a, m.b, m.d, m.x, P);
  FCMF_LAUNCH_OK();
  return 0;
}

int gemm_tc_tn(const void* A, int64_t lda, const void* B, int64_t ldb, const float* bias, void* D, int64_t ldd,
               void* aux, int64_t ldaux, int64_t M, int64_t N, int64_t K, int epi, cudaStream_t st) {
  TcParams P{};
  P.Mg = M; P.Ng = N; P.Kg = K;
  P.k_blocks = (int)((K + TC_BK - 1) / TC_BK);
  P.splits = 1;
  P.bias = bias; P.epi = epi; P.f32_mode = 0;
  P.has_aux = (aux != nullptr && (epi == FCMF_EPI_GELU || epi == FCMF_EPI_DGELU)) ? 1 : 0;
  Maps m;
  if (int r = make_map(&m.d, D, M, N, ldd, 64, TC_BM)) return r;                    // output slabs: box [64 cols][128 rows]
  if (P.has_aux) { if (int r = make_map(&m.x, aux, M, N, ldaux, 64, TC_BM)) return r; }
  else m.x = m.d;
  if (use_2cta() && N % 256 == 0 && ((M + 255) / 256) * (N / 256) >= sm_count() / 2) {
    P.m_tiles = (int)((M + 255) / 256);
    P.n_tiles = (int)(N / 256);
    if (int r = make_map(&m.a, A, M, K, lda, TC_BK, 128)) return r;
    if (int r = make_map(&m.b, B, N, K, ldb, TC_BK, 128)) return r;
    return launch2<false, false>(m, P, st);
  }
  P.m_tiles = (int)((M + TC_BM - 1) / TC_BM);
  // 256-wide tiles unless that leaves SMs idle
  const int64_t tiles256 = (int64_t)P.m_tiles * ((N + 255) / 256);
  const bool wide = (N % 256 == 0 || N > 1024) && tiles256 >= sm_count();
  if (int r = make_map(&m.a, A, M, K, lda, TC_BK, TC_BM)) return r;
  if (wide) {
    P.n_tiles = (int)((N + 255) / 256);
    if (int r = make_map(&m.b, B, N, K, ldb, TC_BK, 256)) return r;
    return launch<256, false, false>(m, P, st);
  }
  P.n_tiles = (int)((N + 127) / 128);
  if (int r = make_map(&m.b, B, N, K, ldb, TC_BK, 128)) return r;
  return launch<128, false, false>(m, P, st);
}

// Split count of the weight-gradient reduction: a work unit is (tile, split) and `workers` CTAs (or CTA pairs) walk the
// units in waves, so what matters is how full the LAST wave is. ceil(workers / tiles) -- the first version -- gave e.g.
// 36 tiles x 3 splits = 108 units on 74 pairs (two waves, the second 46 % full: 27 % of the kernel idle) and
// 9 tiles x 9 = 81 units (55 % efficiency). Pick the split count with the best wave efficiency, fewest splits on ties
// (each split adds one fp32 atomic pass over the tile).
static int pick_splits(int64_t tiles, int64_t workers, int k_blocks) {
  int best = 1;
  double best_eff = 0.0;
  const int max_splits = k_blocks < 32 ? k_blocks : 32;
  for (int s = 1; s <= max_splits; ++s) {
    if (s > 1 && (int64_t)(s - 1) * ((k_blocks + s - 1) / s) >= k_blocks) continue;      // would leave an empty split
    const int64_t units = tiles * s;
    const int64_t waves = (units + workers - 1) / workers;
    // every split adds one fp32-atomic epilogue pass over the tile (~60 k-blocks' worth of time): negligible against the
    // 7 000-block reductions of the fusion path, dominant for a short reduction with a huge output (the vocabulary
    // projection's weight gradient: 32 blocks, 768 MB of output -- three splits ran at 190 TFLOP/s behind their atomics)
    const double eff = (double)units / (double)(waves * workers) / (1.0 + 60.0 * s / (double)k_blocks);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
  }
  return best;
}

int gemm_tc_wgrad(const void* dY, int64_t lddy, const void* X, int64_t ldx, float* dW, int64_t M, int64_t N, int64_t K,
                  int accumulate, cudaStream_t st) {
  // dW[N,K] = sum_m dY[m,n] X[m,k]: GEMM rows = N (out features), cols = K (in features), reduction = M rows.
  TcParams P{};
  P.Mg = N; P.Ng = K; P.Kg = M;
  P.k_blocks = (int)((M + TC_BK - 1) / TC_BK);
  P.Df = dW; P.lddf = K; P.epi = FCMF_EPI_NONE;
  Maps m;
  if (int r = make_map(&m.a, dY, M, N, lddy, 64, TC_BK)) return r;       // box [64 features][64 rows]
  if (int r = make_map(&m.b, X, M, K, ldx, 64, TC_BK)) return r;
  m.d = m.a; m.x = m.a;                                                  // unused in fp32 mode
  const bool two = use_2cta() && N % 256 == 0 && K % 256 == 0;
  int bn = 256;
  if (two) {
    P.m_tiles = (int)(N / 256);
    P.n_tiles = (int)(K / 256);
    P.splits = pick_splits((int64_t)P.m_tiles * P.n_tiles, sm_count() / 2, P.k_blocks);
  } else {
    bn = (K % 256 == 0) ? 256 : 128;
    P.m_tiles = (int)((N + TC_BM - 1) / TC_BM);
    P.n_tiles = (int)((K + bn - 1) / bn);
    P.splits = pick_splits((int64_t)P.m_tiles * P.n_tiles, sm_count(), P.k_blocks);
  }
  P.f32_mode = (P.splits == 1 && !accumulate) ? 1 : 2;
  if (P.f32_mode == 2 && !accumulate) FCMF_CUDA_OK(cudaMemsetAsync(dW, 0, sizeof(float) * N * K, st));
  if (two) return launch2<true, true>(m, P, st);
  return bn == 256 ? launch<256, true, true>(m, P, st) : launch<128, true, true>(m, P, st);
}

// D(fp32)[M,N] = A[M,K] . B[N,K]^T with the reduction split over CTAs (fp32 atomics): for contractions whose output is small
// and whose reduction is long -- the input gradient of the vocabulary projection, [2048 x 768] over K = 250 112: 24 output
// tiles would occupy 24 of 74 CTA pairs.
int gemm_tc_tn_f32(const void* A, int64_t lda, const void* B, int64_t ldb, float* D, int64_t ldd, int64_t M, int64_t N,
                   int64_t K, cudaStream_t st) {
  TcParams P{};
  P.Mg = M; P.Ng = N; P.Kg = K;
  P.k_blocks = (int)((K + TC_BK - 1) / TC_BK);
  P.Df = D; P.lddf = ldd; P.epi = FCMF_EPI_NONE;
  Maps m;
  const bool two = use_2cta() && N % 256 == 0;
  int bn = 256;
  if (two) {
    P.m_tiles = (int)((M + 255) / 256);
    P.n_tiles = (int)(N / 256);
    P.splits = pick_splits((int64_t)P.m_tiles * P.n_tiles, sm_count() / 2, P.k_blocks);
    if (int r = make_map(&m.a, A, M, K, lda, TC_BK, 128)) return r;
    if (int r = make_map(&m.b, B, N, K, ldb, TC_BK, 128)) return r;
  } else {
    bn = (N % 256 == 0) ? 256 : 128;
    P.m_tiles = (int)((M + TC_BM - 1) / TC_BM);
    P.n_tiles = (int)((N + bn - 1) / bn);
    P.splits = pick_splits((int64_t)P.m_tiles * P.n_tiles, sm_count(), P.k_blocks);
    if (int r = make_map(&m.a, A, M, K, lda, TC_BK, TC_BM)) return r;
    if (int r = make_map(&m.b, B, N, K, ldb, TC_BK, bn)) return r;
  }
  m.d = m.a; m.x = m.a;                                                  // unused in fp32 mode
  P.f32_mode = P.splits == 1 ? 1 : 2;
  if (P.f32_mode == 2) FCMF_CUDA_OK(cudaMemset2DAsync(D, sizeof(float) * ldd, 0, sizeof(float) * N, M, st));
  if (two) return launch2<false, false>(m, P, st);
  return bn == 256 ? launch<256, false, false>(m, P, st) : launch<128, false, false>(m, P, st);
}

// Tiling decision of gemm_tc_wgrad for a shape, without launching anything (host only): lets callers and the CPU
// test-suite audit the wave efficiency of the split-K choice. workers = CTA pairs (2-CTA kernel) or CTAs.
void gemm_tc_wgrad_plan(int64_t M, int64_t N, int64_t K, int* pair, int* tiles, int* splits, int* workers) {
  const int k_blocks = (int)((M + TC_BK - 1) / TC_BK);
  const bool two = use_2cta() && N % 256 == 0 && K % 256 == 0;
  int t, w;
  if (two) { t = (int)(N / 256) * (int)(K / 256); w = sm_count() / 2; }
  else {
    const int bn = (K % 256 == 0) ? 256 : 128;
    t = (int)((N + TC_BM - 1) / TC_BM) * (int)((K + bn - 1) / bn);
    w = sm_count();
  }
  *pair = two ? 1 : 0; *tiles = t; *workers = w; *splits = pick_splits(t, w, k_blocks);
}

}  // namespace fcmf
