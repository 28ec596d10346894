// HBM-bound row-wise kernels: residual + TF-style LayerNorm (fwd/bwd), additive masks, row gather/sum,
// tanh backward and weight staging casts. One warp owns one row; 16-byte vector accesses; row kept in
// registers between the statistics pass and the normalisation pass (one HBM read per operand); the streamed operands
// of the next rows arrive through per-warp shared-memory row rings filled by bulk async copies.
#include "common.cuh"
#include <stdio.h>
#include <stdlib.h>

namespace fcmf {

constexpr int LN_WARPS = 4;

// raw 16-byte vectors: what the software pipeline keeps in flight (half the registers of the converted fp32 values)
template <typename T> struct Raw16;
template <> struct Raw16<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void unpack(const uint4& r, float (&v)[4]) {
    v[0] = __uint_as_float(r.x); v[1] = __uint_as_float(r.y); v[2] = __uint_as_float(r.z); v[3] = __uint_as_float(r.w);
  }
  static __device__ __forceinline__ uint4 pack(const float (&v)[4]) {
    return make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
  }
};
template <> struct Raw16<bf16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void unpack(const uint4& r, float (&v)[8]) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
  }
  static __device__ __forceinline__ uint4 pack(const float (&v)[8]) {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    return t;
  }
};
__device__ __forceinline__ uint4 ld16(const void* p) { return *reinterpret_cast<const uint4*>(p); }
// the same vectors as pairs: the arithmetic below is written on float2 (FADD2 / FMUL2 / FFMA2: one issue slot per two elements;
// both kernels were issue-limited -- 533 warp-instructions per row in the forward)
template <typename T> struct Pair16;
template <> struct Pair16<float> {
  static constexpr int NP = 2;
  static __device__ __forceinline__ void unpack(const uint4& r, float2 (&v)[2]) {
    v[0] = make_float2(__uint_as_float(r.x), __uint_as_float(r.y)); v[1] = make_float2(__uint_as_float(r.z), __uint_as_float(r.w));
  }
  static __device__ __forceinline__ uint4 pack(const float2 (&v)[2]) {
    return make_uint4(__float_as_uint(v[0].x), __float_as_uint(v[0].y), __float_as_uint(v[1].x), __float_as_uint(v[1].y));
  }
};
template <> struct Pair16<bf16> {
  static constexpr int NP = 4;
  static __device__ __forceinline__ void unpack(const uint4& r, float2 (&v)[4]) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = make_float2(__uint_as_float(w[i] << 16), __uint_as_float(w[i] & 0xffff0000u));
  }
  static __device__ __forceinline__ uint4 pack(const float2 (&v)[4]) {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __float22bfloat162_rn(v[i]);
    return t;
  }
};
__device__ __forceinline__ float2 bc2(float a) { return make_float2(a, a); }

// Row ring: every warp owns LN_SLOTS row slots in shared memory, filled by bulk async copies (cp.async.bulk global ->
// shared, completion on the slot's mbarrier) that one lane issues one to two rows AHEAD: what is in flight no longer
// lives in registers. With register prefetch alone a warp could afford one row ahead of at most three operand streams
// (168 registers, 12 warps per SM) and the backward streamed 3.1-3.6 TB/s; the ring keeps ~120 KB per SM in flight.
#ifndef FCMF_LN_SLOTS
#define FCMF_LN_SLOTS 2
#endif
#ifndef FCMF_LN_BWD_BLOCKS
#define FCMF_LN_BWD_BLOCKS 3
#endif
#ifndef FCMF_LN_FWD_BLOCKS
#define FCMF_LN_FWD_BLOCKS 5
#endif
constexpr int LN_SLOTS = FCMF_LN_SLOTS;
__device__ __forceinline__ uint32_t ln_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ln_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ln_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void ln_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ln_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ln_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(ln_smem_u32(bar)), "r"(parity), "r"(100000u) : "memory");
    if (ok) return;
    if (clock64() - t0 > 6000000000LL) { printf("fcmf layernorm: row ring wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x); __trap(); }
  }
}
__device__ __forceinline__ void ln_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(ln_smem_u32(dst)), "l"(src), "r"(bytes), "r"(ln_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint4 lds16(const void* p) { return *reinterpret_cast<const uint4*>(p); }

// ------------------------------------------------------------------------------------------- LayerNorm fwd
// Persistent warps, one warp = one row at a time (grid-stride), VPL 16-byte vectors per lane. gamma/beta are staged ONCE
// per block in shared memory (the first version re-read them per row with 2*H/32 scalar loads per lane and was LSU-issue
// bound). The streamed operand x comes through the warp's row ring (see LN_SLOTS); the gathered residual row and the residual
// index of the row after the next are prefetched in registers. Arithmetic on float2 pairs.
// Tried and dropped: a row split over three warps with one vector per lane -- 2-3x the resident warps, but the per-row work
// (reductions, exchange, addresses) is paid per warp: 1 030 instead of 530 warp-instructions per row, issue-bound at 0.58 ms.
// DROP: the dense output x goes through dropout BEFORE the residual add (BertSelfOutput / BertOutput,
// mm_modeling.py:278, 326): s = keep(row, col) * x / (1 - p) + res, mask regenerated from (seed, row, col).
template <typename T, int VPL, bool DROP>
__global__ void __launch_bounds__(LN_WARPS * 32, FCMF_LN_FWD_BLOCKS)
ln_fwd_kernel(const T* __restrict__ x, const T* __restrict__ res, const int32_t* __restrict__ res_idx,
              const float* __restrict__ gamma, const float* __restrict__ beta, T* __restrict__ y,
              float* __restrict__ mean, float* __restrict__ rstd, int64_t M, int H, float eps, fcmf_dropout drop) {
  constexpr int N = Raw16<T>::N, NP = Pair16<T>::NP;
  extern __shared__ __align__(16) uint8_t ln_sm[];              // ring [LN_WARPS][LN_SLOTS][H] of T | gamma[H] | beta[H] | mbarriers
  const uint32_t rowbytes = (uint32_t)H * (uint32_t)sizeof(T);
  const size_t ring_bytes = (size_t)LN_WARPS * LN_SLOTS * rowbytes;
  float* ln_params = reinterpret_cast<float*>(ln_sm + ring_bytes);   // 16-byte conflict-free reads per row
  uint64_t* bars = reinterpret_cast<uint64_t*>(ln_sm + ring_bytes + sizeof(float) * 2 * H);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t* my_ring = ln_sm + (size_t)warp * LN_SLOTS * rowbytes;
  uint64_t* my_bar = bars + warp * LN_SLOTS;
  DropCfg dc;
  if (DROP) dc = make_drop(drop);
  for (int c = threadIdx.x; c < H; c += blockDim.x) { ln_params[c] = gamma[c]; ln_params[H + c] = beta[c]; }
  if (lane == 0) {
#pragma unroll
    for (int sl = 0; sl < LN_SLOTS; ++sl) ln_mbar_init(&my_bar[sl], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const float inv_h = 1.0f / (float)H;
  const int64_t stride = (int64_t)gridDim.x * LN_WARPS;
  int64_t row = (int64_t)blockIdx.x * LN_WARPS + warp;
  auto fill = [&](int slot, int64_t r) {               // lane 0: row r of x -> slot
    ln_mbar_expect_tx(&my_bar[slot], rowbytes);
    ln_bulk_g2s(my_ring + (size_t)slot * rowbytes, x + r * H, rowbytes, &my_bar[slot]);
  };
  uint4 nr[VPL];                                        // the gathered residual row (L2-resident) is prefetched in registers
  int32_t idx_raw = 0;                               // residual index of the NEXT row, kept raw: widening it here would wait for the load
  if (lane == 0) {
#pragma unroll
    for (int sl = 0; sl < LN_SLOTS; ++sl) if (row + sl * stride < M) fill(sl, row + sl * stride);
  }
  if (row < M && res) {
    const int64_t ri = res_idx ? (int64_t)res_idx[row] : row;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * N;
      if (c < H) nr[i] = ld16(res + ri * H + c);
    }
  }
  if (res && res_idx && row + stride < M) idx_raw = res_idx[row + stride];
  uint32_t it = 0;
  for (; row < M; row += stride, ++it) {
    const int slot = (int)(it % LN_SLOTS);
    const uint8_t* src = my_ring + (size_t)slot * rowbytes;
    ln_mbar_wait(&my_bar[slot], (it / LN_SLOTS) & 1);
    float2 v[VPL][NP];
    float2 sum2 = make_float2(0.f, 0.f);
    uint32_t rseed = 0;
    if (DROP) rseed = drop_rowseed(dc.seed, (uint64_t)row);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * N;
      if (c < H) {
        Pair16<T>::unpack(lds16(src + c * (int)sizeof(T)), v[i]);
        if (DROP) {
#pragma unroll
          for (int j = 0; j < NP; ++j) {
            const uint32_t hsh = drop_pair(rseed, (uint32_t)(c + 2 * j));
            v[i][j] = __fmul2_rn(v[i][j], make_float2(drop_keep_lo(hsh, dc.thr16) ? dc.inv_keep : 0.f, drop_keep_hi(hsh, dc.thr16) ? dc.inv_keep : 0.f));
          }
        }
        if (res) {
          float2 b[NP];
          Pair16<T>::unpack(nr[i], b);
#pragma unroll
          for (int j = 0; j < NP; ++j) v[i][j] = __fadd2_rn(v[i][j], b[j]);
        }
#pragma unroll
        for (int j = 0; j < NP; ++j) sum2 = __fadd2_rn(sum2, v[i][j]);
      } else {
#pragma unroll
        for (int j = 0; j < NP; ++j) v[i][j] = make_float2(0.f, 0.f);
      }
    }
    __syncwarp();                                       // every lane has its vectors: the slot takes the row LN_SLOTS ahead
    if (lane == 0 && row + LN_SLOTS * stride < M) fill(slot, row + LN_SLOTS * stride);
    const int64_t nrow = row + stride;
    if (nrow < M && res) {                              // the next row's residual goes out now; it lands while this row is reduced and stored
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int c = (i * 32 + lane) * N;
        if (c < H) nr[i] = ld16(res + (res_idx ? (int64_t)idx_raw : nrow) * H + c);
      }
      if (res_idx && nrow + stride < M) idx_raw = res_idx[nrow + stride];
    }
    const float mu = warp_sum(sum2.x + sum2.y) * inv_h;
    const float2 nmu = bc2(-mu);
    float2 sq2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * N;
      if (c < H) {
#pragma unroll
        for (int j = 0; j < NP; ++j) { const float2 d = __fadd2_rn(v[i][j], nmu); sq2 = __ffma2_rn(d, d, sq2); }
      }
    }
    const float var = warp_sum(sq2.x + sq2.y) * inv_h;  // biased variance, mm_modeling.py:169
    const float rs = 1.0f / sqrtf(var + eps);           // eps inside the sqrt, mm_modeling.py:170
    if (lane == 0) { if (mean) mean[row] = mu; if (rstd) rstd[row] = rs; }
    const float2 rs2 = bc2(rs);
    T* yr = y + row * H;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * N;
      if (c < H) {
        float2 o[NP];
#pragma unroll
        for (int j = 0; j < NP; j += 2) {
          const float4 gv = *reinterpret_cast<const float4*>(ln_params + c + 2 * j);
          const float4 bv = *reinterpret_cast<const float4*>(ln_params + H + c + 2 * j);
          o[j] = __ffma2_rn(make_float2(gv.x, gv.y), __fmul2_rn(__fadd2_rn(v[i][j], nmu), rs2), make_float2(bv.x, bv.y));
          o[j + 1] = __ffma2_rn(make_float2(gv.z, gv.w), __fmul2_rn(__fadd2_rn(v[i][j + 1], nmu), rs2), make_float2(bv.z, bv.w));
        }
        *reinterpret_cast<uint4*>(yr + c) = Pair16<T>::pack(o);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------- LayerNorm bwd
// DROP: ds (gradient of the LayerNorm input s, what the residual receives) and dx = keep * ds / (1 - p) (what the
// dense output receives) are both written; the keep bits of a row are packed into `keep` while x is converted.
// The three streamed operands (x, dy, dy_add) of a row come through the warp's row ring (see LN_SLOTS); the gathered
// residual row (L2-resident: it is shared by the 7 image problems of a sample-aspect pair), mean / rstd and the residual index
// of the row after the next are prefetched in registers.
template <typename T, int VPL, bool DROP>
__global__ void __launch_bounds__(LN_WARPS * 32, FCMF_LN_BWD_BLOCKS)
ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ dy_add, const T* __restrict__ x, const T* __restrict__ res,
              const int32_t* __restrict__ res_idx, const float* __restrict__ gamma,
              const float* __restrict__ mean, const float* __restrict__ rstd, T* __restrict__ ds, T* __restrict__ dx,
              float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t M, int H, fcmf_dropout drop, int dy_every) {
  constexpr int N = Raw16<T>::N, NP = Pair16<T>::NP;
  extern __shared__ __align__(16) uint8_t ln_sm[];   // ring [LN_WARPS][LN_SLOTS][x | dy | dy_add][H] of T | gamma[H] f32 | mbarriers
  const uint32_t rowbytes = (uint32_t)H * (uint32_t)sizeof(T);
  const size_t ring_bytes = (size_t)LN_WARPS * LN_SLOTS * 3 * rowbytes;
  float* red = reinterpret_cast<float*>(ln_sm);       // [LN_WARPS][2][H] column partial sums: reuses the ring once the rows are done
  float* gsm = reinterpret_cast<float*>(ln_sm + ring_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ln_sm + ring_bytes + sizeof(float) * H);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t* my_ring = ln_sm + (size_t)warp * LN_SLOTS * 3 * rowbytes;
  uint64_t* my_bar = bars + warp * LN_SLOTS;
  DropCfg dc;
  if (DROP) dc = make_drop(drop);
  for (int c = threadIdx.x; c < H; c += blockDim.x) gsm[c] = gamma[c];
  if (lane == 0) {
#pragma unroll
    for (int sl = 0; sl < LN_SLOTS; ++sl) ln_mbar_init(&my_bar[sl], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  float2 dg[VPL][NP], db[VPL][NP];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
#pragma unroll
    for (int j = 0; j < NP; ++j) { dg[i][j] = make_float2(0.f, 0.f); db[i][j] = make_float2(0.f, 0.f); }
  }
  const float inv_h = 1.0f / (float)H;
  const float dc_inv = DROP ? dc.inv_keep : 1.0f;
  const int64_t stride = (int64_t)gridDim.x * LN_WARPS;
  int64_t row = (int64_t)blockIdx.x * LN_WARPS + warp;
  // dy_every > 0: dy is COMPACT -- only rows m with m % dy_every == 0 have a gradient (row m / dy_every of dy), the others
  // are structurally zero (BertPooler reads token 0 of every per-image branch, mm_modeling.py:428): no zero tensor is
  // materialised and nothing is read for them
  auto has_dy = [&](int64_t r) { return dy_every <= 0 || (r % dy_every) == 0; };
  auto fill = [&](int slot, int64_t r) {               // lane 0: the streamed operands of row r -> slot
    uint8_t* dst = my_ring + (size_t)slot * 3 * rowbytes;
    const bool hd = has_dy(r);
    ln_mbar_expect_tx(&my_bar[slot], rowbytes * (1u + (hd ? 1u : 0u) + (dy_add ? 1u : 0u)));
    ln_bulk_g2s(dst, x + r * H, rowbytes, &my_bar[slot]);
    if (hd) ln_bulk_g2s(dst + rowbytes, dy + (dy_every > 0 ? r / dy_every : r) * H, rowbytes, &my_bar[slot]);
    if (dy_add) ln_bulk_g2s(dst + 2 * rowbytes, dy_add + r * H, rowbytes, &my_bar[slot]);
  };
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  uint4 nr[VPL];
  float mu_n = 0.f, rs_n = 0.f;
  int32_t idx_raw = 0;                               // residual index of the NEXT row, kept raw: widening it here would wait for the load
  auto request = [&](int64_t r, int64_t ri) {          // register-side prefetch: statistics and the gathered residual row
    mu_n = mean[r]; rs_n = rstd[r];
    if (res) {
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int c = (i * 32 + lane) * N;
        if (c < H) nr[i] = ld16(res + ri * H + c);
      }
    }
  };
  if (lane == 0) {
#pragma unroll
    for (int sl = 0; sl < LN_SLOTS; ++sl) if (row + sl * stride < M) fill(sl, row + sl * stride);
  }
  if (row < M) request(row, res ? (res_idx ? (int64_t)res_idx[row] : row) : 0);
  if (res && res_idx && row + stride < M) idx_raw = res_idx[row + stride];
  uint32_t it = 0;
  for (; row < M; row += stride, ++it) {
    const int slot = (int)(it % LN_SLOTS);
    const uint8_t* src = my_ring + (size_t)slot * 3 * rowbytes;
    const bool hd = has_dy(row);
    ln_mbar_wait(&my_bar[slot], (it / LN_SLOTS) & 1);
    const float mu = mu_n, rs = rs_n;
    const float2 rs2 = bc2(rs), nmr2 = bc2(-mu * rs);
    float2 xh[VPL][NP], gy[VPL][NP];
    float2 s1v = make_float2(0.f, 0.f), s2v = make_float2(0.f, 0.f);
    uint32_t keep[VPL];                               // bit j of keep[i]: element (i, j) of this lane survived dropout
    uint32_t rseed = 0;
    if (DROP) rseed = drop_rowseed(dc.seed, (uint64_t)row);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * N;
      keep[i] = 0xffffffffu;
      if (c < H) {
        float2 a[NP], d[NP];
        const int cb = c * (int)sizeof(T);              // byte offset of this lane's vector inside the row
        Pair16<T>::unpack(lds16(src + cb), a);
        Pair16<T>::unpack(hd ? lds16(src + rowbytes + cb) : zero4, d);
        if (DROP) {
          uint32_t kb = 0;
#pragma unroll
          for (int j = 0; j < NP; ++j) {
            const uint32_t hsh = drop_pair(rseed, (uint32_t)(c + 2 * j));
            const bool k0 = drop_keep_lo(hsh, dc.thr16), k1 = drop_keep_hi(hsh, dc.thr16);
            a[j] = __fmul2_rn(a[j], make_float2(k0 ? dc.inv_keep : 0.f, k1 ? dc.inv_keep : 0.f));
            kb |= (k0 ? 1u : 0u) << (2 * j) | (k1 ? 1u : 0u) << (2 * j + 1);
          }
          keep[i] = kb;
        }
        if (dy_add) {
          float2 e[NP];
          Pair16<T>::unpack(lds16(src + 2 * rowbytes + cb), e);
#pragma unroll
          for (int j = 0; j < NP; ++j) d[j] = __fadd2_rn(d[j], e[j]);
        }
        if (res) {
          float2 b[NP];
          Pair16<T>::unpack(nr[i], b);
#pragma unroll
          for (int j = 0; j < NP; ++j) a[j] = __fadd2_rn(a[j], b[j]);
        }
#pragma unroll
        for (int j = 0; j < NP; j += 2) {
          const float4 gv = *reinterpret_cast<const float4*>(gsm + c + 2 * j);
          const float2 gq[2] = {make_float2(gv.x, gv.y), make_float2(gv.z, gv.w)};
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            xh[i][j + t] = __ffma2_rn(a[j + t], rs2, nmr2);            // (a - mu) * rs, as the forward rounds it up to one ulp
            gy[i][j + t] = __fmul2_rn(d[j + t], gq[t]);
            s1v = __fadd2_rn(s1v, gy[i][j + t]);
            s2v = __ffma2_rn(gy[i][j + t], xh[i][j + t], s2v);
            dg[i][j + t] = __ffma2_rn(d[j + t], xh[i][j + t], dg[i][j + t]);
            db[i][j + t] = __fadd2_rn(db[i][j + t], d[j + t]);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < NP; ++j) { xh[i][j] = make_float2(0.f, 0.f); gy[i][j] = make_float2(0.f, 0.f); }
      }
    }
    __syncwarp();                                       // every lane has consumed its vectors: the slot takes the row LN_SLOTS ahead
    if (lane == 0 && row + LN_SLOTS * stride < M) fill(slot, row + LN_SLOTS * stride);
    const int64_t nrow = row + stride;
    if (nrow < M) {
      request(nrow, res_idx ? (int64_t)idx_raw : nrow);
      if (res && res_idx && nrow + stride < M) idx_raw = res_idx[nrow + stride];
    }
    const float s1 = warp_sum(s1v.x + s1v.y) * inv_h;
    const float s2 = warp_sum(s2v.x + s2v.y) * inv_h;
    // ds = rs * (gy - s1 - xh * s2) = xh * (-rs s2) + (gy * rs + (-rs s1))
    const float2 c0 = bc2(-rs * s1), c1 = bc2(-rs * s2), ik2 = bc2(dc_inv);
    T* dsr = ds + row * H;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * N;
      if (c < H) {
        float2 o[NP];
#pragma unroll
        for (int j = 0; j < NP; ++j) o[j] = __ffma2_rn(xh[i][j], c1, __ffma2_rn(gy[i][j], rs2, c0));
        *reinterpret_cast<uint4*>(dsr + c) = Pair16<T>::pack(o);
        if (DROP) {
#pragma unroll
          for (int j = 0; j < NP; ++j)
            o[j] = __fmul2_rn(o[j], make_float2(((keep[i] >> (2 * j)) & 1u) ? ik2.x : 0.f, ((keep[i] >> (2 * j + 1)) & 1u) ? ik2.y : 0.f));
          *reinterpret_cast<uint4*>(dx + row * H + c) = Pair16<T>::pack(o);
        }
      }
    }
  }
  // block reduction of the column sums, then one atomic per column per block
  __syncthreads();                                    // every warp is done with its ring slots: `red` takes the space over
  float* my = red + (size_t)warp * 2 * H;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = (i * 32 + lane) * N;
    if (c < H) {
#pragma unroll
      for (int j = 0; j < NP; ++j) {
        my[c + 2 * j] = dg[i][j].x; my[c + 2 * j + 1] = dg[i][j].y;
        my[H + c + 2 * j] = db[i][j].x; my[H + c + 2 * j + 1] = db[i][j].y;
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * H; c += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < LN_WARPS; ++w) s += red[(size_t)w * 2 * H + c];
    atomicAdd((c < H ? dgamma + c : dbeta + (c - H)), s);
  }
}

template <typename T, int VPL, bool DROP>
static int ln_fwd_go(const void* x, const void* res, const int32_t* idx, const float* gamma, const float* beta, void* y, float* mean,
                     float* rstd, int64_t M, int H, float eps, const fcmf_dropout& drop, cudaStream_t st) {
  auto kern = ln_fwd_kernel<T, VPL, DROP>;
  const size_t smem = (size_t)LN_WARPS * LN_SLOTS * H * sizeof(T) + sizeof(float) * 2 * H + sizeof(uint64_t) * LN_WARPS * LN_SLOTS;
  static int bps[64] = {0};                           // resident blocks per SM of this instantiation at this H, per device
  static int bps_h[64] = {0};
  int dev = 0;
  FCMF_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) dev = 0;
  if (bps[dev] == 0 || bps_h[dev] != H) {
    FCMF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int n = 0;
    FCMF_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, LN_WARPS * 32, smem));
    if (n < 1) return fail(FCMF_ERR_UNSUPPORTED, "layernorm fwd: H=%d needs %zu bytes of shared memory per block", H, smem);
    bps[dev] = n; bps_h[dev] = H;
  }
  int64_t blocks = (M + LN_WARPS - 1) / LN_WARPS;
  const int64_t cap = (int64_t)sm_count() * bps[dev];              // one resident wave of persistent warps
  if (blocks > cap) blocks = cap;
  kern<<<(unsigned)blocks, LN_WARPS * 32, smem, st>>>((const T*)x, (const T*)res, idx, gamma, beta, (T*)y, mean, rstd, M, H, eps, drop);
  FCMF_LAUNCH_OK();
  return 0;
}

template <typename T, int VPL>
static int ln_fwd_launch(const void* x, const void* res, const int32_t* idx, const float* gamma, const float* beta,
                         void* y, float* mean, float* rstd, int64_t M, int H, float eps, const fcmf_dropout* drop,
                         cudaStream_t st) {
  if (drop_on(drop)) return ln_fwd_go<T, VPL, true>(x, res, idx, gamma, beta, y, mean, rstd, M, H, eps, *drop, st);
  return ln_fwd_go<T, VPL, false>(x, res, idx, gamma, beta, y, mean, rstd, M, H, eps, drop_or_off(nullptr), st);
}

template <typename T, int VPL, bool DROP>
static int ln_bwd_go(const void* dy, const void* dy_add, const void* x, const void* res, const int32_t* idx, const float* gamma,
                     const float* mean, const float* rstd, void* ds, void* dx, float* dgamma, float* dbeta, int64_t M, int H,
                     const fcmf_dropout& drop, int dy_every, cudaStream_t st) {
  auto kern = ln_bwd_kernel<T, VPL, DROP>;
  const size_t smem = (size_t)LN_WARPS * LN_SLOTS * 3 * H * sizeof(T) + sizeof(float) * H + sizeof(uint64_t) * LN_WARPS * LN_SLOTS;
  static int bps[64] = {0};                           // resident blocks per SM of this instantiation at this H, per device
  static int bps_h[64] = {0};
  int dev = 0;
  FCMF_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) dev = 0;
  if (bps[dev] == 0 || bps_h[dev] != H) {
    FCMF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int n = 0;
    FCMF_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, LN_WARPS * 32, smem));
    if (n < 1) return fail(FCMF_ERR_UNSUPPORTED, "layernorm bwd: H=%d needs %zu bytes of shared memory per block", H, smem);
    bps[dev] = n; bps_h[dev] = H;
  }
  int64_t blocks = (M + LN_WARPS - 1) / LN_WARPS;
  const int64_t cap = (int64_t)sm_count() * bps[dev];              // one resident wave of persistent warps
  if (blocks > cap) blocks = cap;
  kern<<<(unsigned)blocks, LN_WARPS * 32, smem, st>>>((const T*)dy, (const T*)dy_add, (const T*)x, (const T*)res, idx, gamma, mean, rstd,
                                                      (T*)ds, (T*)dx, dgamma, dbeta, M, H, drop, dy_every);
  FCMF_LAUNCH_OK();
  return 0;
}

template <typename T, int VPL>
static int ln_bwd_launch(const void* dy, const void* dy_add, const void* x, const void* res, const int32_t* idx, const float* gamma,
                         const float* mean, const float* rstd, void* ds, void* dx, float* dgamma, float* dbeta, int64_t M, int H,
                         const fcmf_dropout* drop, int dy_every, cudaStream_t st) {
  if (drop_on(drop)) return ln_bwd_go<T, VPL, true>(dy, dy_add, x, res, idx, gamma, mean, rstd, ds, dx, dgamma, dbeta, M, H, *drop, dy_every, st);
  return ln_bwd_go<T, VPL, false>(dy, dy_add, x, res, idx, gamma, mean, rstd, ds, dx, dgamma, dbeta, M, H, drop_or_off(nullptr), dy_every, st);
}

#define FCMF_LN_DISPATCH(T, fn, ...)                                          \
  do {                                                                        \
    const int per = 32 * Vec16<T>::N;                                         \
    const int vpl = (H + per - 1) / per;                                      \
    switch (vpl) {                                                            \
      case 1: return fn<T, 1>(__VA_ARGS__);                                   \
      case 2: return fn<T, 2>(__VA_ARGS__);                                   \
      case 3: return fn<T, 3>(__VA_ARGS__);                                   \
      case 4: return fn<T, 4>(__VA_ARGS__);                                   \
      case 5: case 6: return fn<T, 6>(__VA_ARGS__);                           \
      case 7: case 8: return fn<T, 8>(__VA_ARGS__);                           \
      default: return fail(FCMF_ERR_UNSUPPORTED, "layernorm: H=%d too wide", H); \
    }                                                                         \
  } while (0)

// ------------------------------------------------------------------------------------------- small kernels
__global__ void mask_additive_kernel(const int64_t* __restrict__ mask, int64_t ld, float* __restrict__ add,
                                     int64_t rows, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * n) return;
  const int64_t r = i / n, j = i - r * n;
  add[i] = (1.0f - (float)mask[r * ld + j]) * -10000.0f;     // fcmf_pretraining.py:56
}

template <typename T>
__global__ void gather_sum_rows_kernel(const T* __restrict__ src, int64_t ldsrc, const int32_t* __restrict__ idx,
                                       T* __restrict__ out, int64_t ldout, int64_t n_out, int G, int width,
                                       int accumulate) {
  constexpr int N = Vec16<T>::N;
  const int64_t o = blockIdx.x;
  const int32_t* ix = idx + o * G;
  for (int c = threadIdx.x * N; c < width; c += blockDim.x * N) {
    float acc[N];
#pragma unroll
    for (int j = 0; j < N; ++j) acc[j] = 0.f;
    if (accumulate) { Vec16<T> a; a.load(out + o * ldout + c);
#pragma unroll
      for (int j = 0; j < N; ++j) acc[j] = a.v[j]; }
    for (int g = 0; g < G; ++g) {
      const int32_t r = ix[g];
      if (r < 0) continue;
      Vec16<T> a; a.load(src + (int64_t)r * ldsrc + c);
#pragma unroll
      for (int j = 0; j < N; ++j) acc[j] += a.v[j];
    }
    Vec16<T> w;
#pragma unroll
    for (int j = 0; j < N; ++j) w.v[j] = acc[j];
    w.store(out + o * ldout + c);
  }
}

template <typename T>
__global__ void dtanh_kernel(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ out, int64_t n) {
  constexpr int N = Vec16<T>::N;
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * N;
  if (i + N <= n) {
    Vec16<T> a, b, o; a.load(dy + i); b.load(y + i);
#pragma unroll
    for (int j = 0; j < N; ++j) o.v[j] = a.v[j] * (1.0f - b.v[j] * b.v[j]);
    o.store(out + i);
  } else {
    for (int64_t k = i; k < n; ++k) { const float t = to_f(y[k]); out[k] = from_f<T>(to_f(dy[k]) * (1.0f - t * t)); }
  }
}

template <typename T>
__global__ void cast_matrix_kernel(const float* __restrict__ src, T* __restrict__ dst, int rows, int cols, int transpose, int64_t ld) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? src[(int64_t)r * cols + c] : 0.f;
  }
  __syncthreads();
  if (!transpose) {
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      const int r = r0 + i, c = c0 + threadIdx.x;
      if (r < rows && c < cols) dst[(int64_t)r * ld + c] = from_f<T>(tile[i][threadIdx.x]);
    }
  } else {                                              // dst is [cols, rows] with row stride ld
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      const int c = c0 + i, r = r0 + threadIdx.x;
      if (r < rows && c < cols) dst[(int64_t)c * ld + r] = from_f<T>(tile[threadIdx.x][i]);
    }
  }
}

template <typename T>
__global__ void cast_to_f32_kernel(const T* __restrict__ src, float* __restrict__ dst, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = to_f(src[i]);
}

}  // namespace fcmf

using namespace fcmf;

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

extern "C" int fcmf_ln_fwd(const void* x, const void* res, const int32_t* res_idx, const float* gamma,
                           const float* beta, void* y, float* mean, float* rstd, int64_t M, int64_t H64, float eps,
                           const fcmf_dropout* drop, int dtype, void* stream) {
  const int H = (int)H64;
  FCMF_CHECK_ARG(M >= 0 && H > 0, "ln_fwd: bad shape");
  FCMF_CHECK_ARG(H % (dtype == FCMF_BF16 ? 8 : 4) == 0, "ln_fwd: H=%d must be a multiple of the 16-byte vector", H);
  FCMF_CHECK_ARG(aligned16(x) && aligned16(y) && (!res || aligned16(res)) && aligned16(gamma) && aligned16(beta), "ln_fwd: pointers must be 16-byte aligned");
  FCMF_CHECK_ARG(drop_check(drop) == 0, "ln_fwd: dropout p must be in [0, 1)");
  if (M == 0) return 0;
  cudaStream_t st = as_stream(stream);
  if (dtype == FCMF_BF16) FCMF_LN_DISPATCH(bf16, ln_fwd_launch, x, res, res_idx, gamma, beta, y, mean, rstd, M, H, eps, drop, st);
  if (dtype == FCMF_F32) FCMF_LN_DISPATCH(float, ln_fwd_launch, x, res, res_idx, gamma, beta, y, mean, rstd, M, H, eps, drop, st);
  return fail(FCMF_ERR_ARG, "ln_fwd: bad dtype %d", dtype);
}

extern "C" int fcmf_ln_bwd(const void* dy, const void* dy_add, const void* x, const void* res, const int32_t* res_idx, const float* gamma,
                           const float* mean, const float* rstd, void* ds, void* dx, float* dgamma, float* dbeta, int64_t M,
                           int64_t H64, const fcmf_dropout* drop, int64_t dy_every, int dtype, void* stream) {
  const int H = (int)H64;
  FCMF_CHECK_ARG(M >= 0 && H > 0, "ln_bwd: bad shape");
  FCMF_CHECK_ARG(H % (dtype == FCMF_BF16 ? 8 : 4) == 0, "ln_bwd: H=%d must be a multiple of the 16-byte vector", H);
  FCMF_CHECK_ARG(aligned16(x) && aligned16(dy) && aligned16(ds) && (!res || aligned16(res)) && (!dy_add || aligned16(dy_add)), "ln_bwd: alignment");
  FCMF_CHECK_ARG(drop_check(drop) == 0, "ln_bwd: dropout p must be in [0, 1)");
  FCMF_CHECK_ARG(!drop_on(drop) || (dx && aligned16(dx)), "ln_bwd: dropout needs the second output dx (16-byte aligned)");
  FCMF_CHECK_ARG(dy_every >= 0 && dy_every < (1LL << 31), "ln_bwd: bad dy_every");
  if (M == 0) return 0;
  cudaStream_t st = as_stream(stream);
  if (dtype == FCMF_BF16) FCMF_LN_DISPATCH(bf16, ln_bwd_launch, dy, dy_add, x, res, res_idx, gamma, mean, rstd, ds, dx, dgamma, dbeta, M, H, drop, (int)dy_every, st);
  if (dtype == FCMF_F32) FCMF_LN_DISPATCH(float, ln_bwd_launch, dy, dy_add, x, res, res_idx, gamma, mean, rstd, ds, dx, dgamma, dbeta, M, H, drop, (int)dy_every, st);
  return fail(FCMF_ERR_ARG, "ln_bwd: bad dtype %d", dtype);
}

extern "C" int fcmf_mask_additive(const int64_t* mask, int64_t ldmask, float* add, int64_t rows, int64_t n, void* stream) {
  FCMF_CHECK_ARG(rows >= 0 && n >= 0 && ldmask >= n, "mask_additive: bad shape");
  if (rows * n == 0) return 0;
  mask_additive_kernel<<<(unsigned)((rows * n + 255) / 256), 256, 0, as_stream(stream)>>>(mask, ldmask, add, rows, n);
  FCMF_LAUNCH_OK();
  return 0;
}

extern "C" int fcmf_gather_sum_rows(const void* src, int64_t ldsrc, const int32_t* idx, void* out, int64_t ldout,
                                    int64_t n_out, int64_t G, int64_t width, int accumulate, int dtype, void* stream) {
  FCMF_CHECK_ARG(n_out >= 0 && G > 0 && width > 0, "gather_sum_rows: bad shape");
  const int vec = dtype == FCMF_BF16 ? 8 : 4;
  FCMF_CHECK_ARG(width % vec == 0 && ldsrc % vec == 0 && ldout % vec == 0 && aligned16(src) && aligned16(out),
                 "gather_sum_rows: width/ld must be multiples of %d elements and pointers 16-byte aligned", vec);
  if (n_out == 0) return 0;
  cudaStream_t st = as_stream(stream);
  int threads = (int)((width / vec + 31) / 32) * 32;
  if (threads > 256) threads = 256;
  if (dtype == FCMF_BF16)
    gather_sum_rows_kernel<bf16><<<(unsigned)n_out, threads, 0, st>>>((const bf16*)src, ldsrc, idx, (bf16*)out, ldout, n_out, (int)G, (int)width, accumulate);
  else if (dtype == FCMF_F32)
    gather_sum_rows_kernel<float><<<(unsigned)n_out, threads, 0, st>>>((const float*)src, ldsrc, idx, (float*)out, ldout, n_out, (int)G, (int)width, accumulate);
  else return fail(FCMF_ERR_ARG, "gather_sum_rows: bad dtype %d", dtype);
  FCMF_LAUNCH_OK();
  return 0;
}

extern "C" int fcmf_dtanh(const void* dy, const void* y, void* out, int64_t n, int dtype, void* stream) {
  FCMF_CHECK_ARG(n >= 0, "dtanh: bad n");
  FCMF_CHECK_ARG(aligned16(dy) && aligned16(y) && aligned16(out), "dtanh: alignment");
  if (n == 0) return 0;
  cudaStream_t st = as_stream(stream);
  const int vec = dtype == FCMF_BF16 ? 8 : 4;
  const unsigned grid = (unsigned)(((n + vec - 1) / vec + 255) / 256);
  if (dtype == FCMF_BF16) dtanh_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)dy, (const bf16*)y, (bf16*)out, n);
  else if (dtype == FCMF_F32) dtanh_kernel<float><<<grid, 256, 0, st>>>((const float*)dy, (const float*)y, (float*)out, n);
  else return fail(FCMF_ERR_ARG, "dtanh: bad dtype %d", dtype);
  FCMF_LAUNCH_OK();
  return 0;
}

extern "C" int fcmf_cast_matrix_ld(const float* src, void* dst, int64_t rows, int64_t cols, int64_t ld_dst, int transpose, int dtype,
                                   void* stream) {
  FCMF_CHECK_ARG(rows >= 0 && cols >= 0 && rows < (1LL << 31) && cols < (1LL << 31), "cast_matrix: bad shape");
  FCMF_CHECK_ARG(ld_dst >= (transpose ? rows : cols), "cast_matrix: ld_dst %lld is smaller than the destination's row length", (long long)ld_dst);
  if (rows * cols == 0) return 0;
  FCMF_CHECK_ARG((rows + 31) / 32 <= 65535, "cast_matrix: more than 65535 x 32 rows");
  dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32)), block(32, 8);
  cudaStream_t st = as_stream(stream);
  if (dtype == FCMF_BF16) cast_matrix_kernel<bf16><<<grid, block, 0, st>>>(src, (bf16*)dst, (int)rows, (int)cols, transpose, ld_dst);
  else if (dtype == FCMF_F32) cast_matrix_kernel<float><<<grid, block, 0, st>>>(src, (float*)dst, (int)rows, (int)cols, transpose, ld_dst);
  else return fail(FCMF_ERR_ARG, "cast_matrix: bad dtype %d", dtype);
  FCMF_LAUNCH_OK();
  return 0;
}

extern "C" int fcmf_cast_matrix(const float* src, void* dst, int64_t rows, int64_t cols, int transpose, int dtype, void* stream) {
  return fcmf_cast_matrix_ld(src, dst, rows, cols, transpose ? rows : cols, transpose, dtype, stream);
}

extern "C" int fcmf_cast_to_f32(const void* src, float* dst, int64_t n, int dtype, void* stream) {
  if (n <= 0) return 0;
  cudaStream_t st = as_stream(stream);
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (dtype == FCMF_BF16) cast_to_f32_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)src, dst, n);
  else if (dtype == FCMF_F32) cast_to_f32_kernel<float><<<grid, 256, 0, st>>>((const float*)src, dst, n);
  else return fail(FCMF_ERR_ARG, "cast_to_f32: bad dtype %d", dtype);
  FCMF_LAUNCH_OK();
  return 0;
}
