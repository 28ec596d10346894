// HBM-bound row-wise kernels: residual + TF-style LayerNorm (fwd/bwd), additive masks, row gather/sum,
// tanh backward and weight staging casts. One warp owns one row; 16-byte vector accesses; row kept in
// registers between the statistics pass and the normalisation pass (one HBM read per operand).
#include "common.cuh"

namespace fcmf {

constexpr int LN_WARPS = 4;

// ------------------------------------------------------------------------------------------- LayerNorm fwd
// Persistent warps (grid-stride over rows): gamma/beta are staged ONCE per block in shared memory -- the first version
// re-read them per row with 2*H/32 scalar loads per lane and was LSU-issue bound (ncu: 533 warp-instructions per row,
// 24 % of DRAM peak) -- and the next row's 16-byte loads are issued before the current row's reductions.
// DROP: the dense output x goes through dropout BEFORE the residual add (BertSelfOutput / BertOutput,
// mm_modeling.py:278, 326): s = keep(row, col) * x / (1 - p) + res, mask regenerated from (seed, row, col).
template <typename T, int VPL, bool DROP>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_fwd_kernel(const T* __restrict__ x, const T* __restrict__ res, const int32_t* __restrict__ res_idx,
              const float* __restrict__ gamma, const float* __restrict__ beta, T* __restrict__ y,
              float* __restrict__ mean, float* __restrict__ rstd, int64_t M, int H, float eps, fcmf_dropout drop) {
  constexpr int N = Vec16<T>::N;
  extern __shared__ float ln_params[];                          // gamma[H] | beta[H]: 16-byte conflict-free reads per row
  const int lane = threadIdx.x & 31;
  DropCfg dc;
  if (DROP) dc = make_drop(drop);
  for (int c = threadIdx.x; c < H; c += blockDim.x) { ln_params[c] = gamma[c]; ln_params[H + c] = beta[c]; }
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * LN_WARPS;
  for (int64_t row = (int64_t)blockIdx.x * LN_WARPS + (threadIdx.x >> 5); row < M; row += stride) {
    const T* xr = x + row * H;
    const T* rr = res ? res + (int64_t)(res_idx ? res_idx[row] : row) * H : nullptr;
    float v[VPL][N];
    float sum = 0.f;
    uint32_t rseed = 0;
    if (DROP) rseed = drop_rowseed(dc.seed, (uint64_t)row);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * N;
      if (c < H) {
        Vec16<T> a; a.load(xr + c);
        if (DROP) {
#pragma unroll
          for (int j = 0; j < N; j += 2) {
            const uint32_t hsh = drop_pair(rseed, (uint32_t)(c + j));
            a.v[j] = drop_keep_lo(hsh, dc.thr16) ? a.v[j] * dc.inv_keep : 0.f;
            a.v[j + 1] = drop_keep_hi(hsh, dc.thr16) ? a.v[j + 1] * dc.inv_keep : 0.f;
          }
        }
        if (rr) { Vec16<T> b; b.load(rr + c);
#pragma unroll
          for (int j = 0; j < N; ++j) a.v[j] += b.v[j]; }
#pragma unroll
        for (int j = 0; j < N; ++j) { v[i][j] = a.v[j]; sum += a.v[j]; }
      } else {
#pragma unroll
        for (int j = 0; j < N; ++j) v[i][j] = 0.f;
      }
    }
    const float mu = warp_sum(sum) / (float)H;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * N;
      if (c < H) {
#pragma unroll
        for (int j = 0; j < N; ++j) { const float d = v[i][j] - mu; sq += d * d; }
      }
    }
    const float var = warp_sum(sq) / (float)H;          // biased variance, mm_modeling.py:169
    const float rs = 1.0f / sqrtf(var + eps);           // eps inside the sqrt, mm_modeling.py:170
    if (lane == 0) { if (mean) mean[row] = mu; if (rstd) rstd[row] = rs; }
    T* yr = y + row * H;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * N;
      if (c < H) {
        Vec16<T> o;
#pragma unroll
        for (int j = 0; j < N; j += 4) {
          const float4 gv = *reinterpret_cast<const float4*>(ln_params + c + j);
          const float4 bv = *reinterpret_cast<const float4*>(ln_params + H + c + j);
          o.v[j] = fmaf(gv.x, (v[i][j] - mu) * rs, bv.x);
          o.v[j + 1] = fmaf(gv.y, (v[i][j + 1] - mu) * rs, bv.y);
          o.v[j + 2] = fmaf(gv.z, (v[i][j + 2] - mu) * rs, bv.z);
          o.v[j + 3] = fmaf(gv.w, (v[i][j + 3] - mu) * rs, bv.w);
        }
        o.store(yr + c);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------- LayerNorm bwd
// DROP: ds (gradient of the LayerNorm input s, what the residual receives) and dx = keep * ds / (1 - p) (what the
// dense output receives) are both written; the keep bits of a row are packed into `keep` while x is loaded.
template <typename T, int VPL, bool DROP>
__global__ void __launch_bounds__(LN_WARPS * 32, 4)
ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ dy_add, const T* __restrict__ x, const T* __restrict__ res,
              const int32_t* __restrict__ res_idx, const float* __restrict__ gamma,
              const float* __restrict__ mean, const float* __restrict__ rstd, T* __restrict__ ds, T* __restrict__ dx,
              float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t M, int H, fcmf_dropout drop, int dy_every) {
  constexpr int N = Vec16<T>::N;
  extern __shared__ float red[];                      // [LN_WARPS][2][H] column partial sums | gamma[H]
  float* gsm = red + (size_t)LN_WARPS * 2 * H;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  DropCfg dc;
  if (DROP) dc = make_drop(drop);
  for (int c = threadIdx.x; c < H; c += blockDim.x) gsm[c] = gamma[c];
  __syncthreads();
  float dg[VPL][N], db[VPL][N];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
#pragma unroll
    for (int j = 0; j < N; ++j) { dg[i][j] = 0.f; db[i][j] = 0.f; }
  }
  const int64_t stride = (int64_t)gridDim.x * LN_WARPS;
  for (int64_t row = (int64_t)blockIdx.x * LN_WARPS + warp; row < M; row += stride) {
    const T* xr = x + row * H;
    const T* rr = res ? res + (int64_t)(res_idx ? res_idx[row] : row) * H : nullptr;
    // dy_every > 0: dy is COMPACT -- only rows m with m % dy_every == 0 have a gradient (row m / dy_every of dy), the others
    // are structurally zero (BertPooler reads token 0 of every per-image branch, mm_modeling.py:428): no zero tensor is
    // materialised and nothing is read for them
    const bool has_dy = dy_every <= 0 || (row % dy_every) == 0;
    const T* dyr = dy + (dy_every > 0 ? row / dy_every : row) * H;
    const float mu = mean[row], rs = rstd[row];
    float xh[VPL][N], gy[VPL][N];
    float s1 = 0.f, s2 = 0.f;
    uint32_t keep[VPL];                               // bit j of keep[i]: element (i, j) of this lane survived dropout
    uint32_t rseed = 0;
    if (DROP) rseed = drop_rowseed(dc.seed, (uint64_t)row);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * N;
      keep[i] = 0xffffffffu;
      if (c < H) {
        Vec16<T> a, d; a.load(xr + c);
        if (has_dy) d.load(dyr + c);
        else {
#pragma unroll
          for (int j = 0; j < N; ++j) d.v[j] = 0.f;
        }
        if (DROP) {
          uint32_t kb = 0;
#pragma unroll
          for (int j = 0; j < N; j += 2) {
            const uint32_t hsh = drop_pair(rseed, (uint32_t)(c + j));
            const bool k0 = drop_keep_lo(hsh, dc.thr16), k1 = drop_keep_hi(hsh, dc.thr16);
            a.v[j] = k0 ? a.v[j] * dc.inv_keep : 0.f;
            a.v[j + 1] = k1 ? a.v[j + 1] * dc.inv_keep : 0.f;
            kb |= (k0 ? 1u : 0u) << j | (k1 ? 1u : 0u) << (j + 1);
          }
          keep[i] = kb;
        }
        if (dy_add) { Vec16<T> e; e.load(dy_add + row * H + c);
#pragma unroll
          for (int j = 0; j < N; ++j) d.v[j] += e.v[j]; }
        if (rr) { Vec16<T> b; b.load(rr + c);
#pragma unroll
          for (int j = 0; j < N; ++j) a.v[j] += b.v[j]; }
        float gq[N];
#pragma unroll
        for (int j = 0; j < N; j += 4) {
          const float4 gv = *reinterpret_cast<const float4*>(gsm + c + j);
          gq[j] = gv.x; gq[j + 1] = gv.y; gq[j + 2] = gv.z; gq[j + 3] = gv.w;
        }
#pragma unroll
        for (int j = 0; j < N; ++j) {
          xh[i][j] = (a.v[j] - mu) * rs;
          gy[i][j] = d.v[j] * gq[j];
          s1 += gy[i][j];
          s2 += gy[i][j] * xh[i][j];
          dg[i][j] += d.v[j] * xh[i][j];
          db[i][j] += d.v[j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < N; ++j) { xh[i][j] = 0.f; gy[i][j] = 0.f; }
      }
    }
    s1 = warp_sum(s1) / (float)H;
    s2 = warp_sum(s2) / (float)H;
    T* dsr = ds + row * H;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * N;
      if (c < H) {
        Vec16<T> o;
#pragma unroll
        for (int j = 0; j < N; ++j) o.v[j] = rs * (gy[i][j] - s1 - xh[i][j] * s2);
        o.store(dsr + c);
        if (DROP) {
#pragma unroll
          for (int j = 0; j < N; ++j) o.v[j] = ((keep[i] >> j) & 1u) ? o.v[j] * dc.inv_keep : 0.f;
          o.store(dx + row * H + c);
        }
      }
    }
  }
  // block reduction of the column sums, then one atomic per column per block
  float* my = red + (size_t)warp * 2 * H;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = (i * 32 + lane) * N;
    if (c < H) {
#pragma unroll
      for (int j = 0; j < N; ++j) { my[c + j] = dg[i][j]; my[H + c + j] = db[i][j]; }
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

template <typename T, int VPL>
static int ln_fwd_launch(const void* x, const void* res, const int32_t* idx, const float* gamma, const float* beta,
                         void* y, float* mean, float* rstd, int64_t M, int H, float eps, const fcmf_dropout* drop,
                         cudaStream_t st) {
  int64_t blocks = (M + LN_WARPS - 1) / LN_WARPS;
  const int64_t cap = (int64_t)sm_count() * 16;                 // up to 64 warps per SM, each looping over rows
  if (blocks > cap) blocks = cap;
  const unsigned grid = (unsigned)blocks;
  if (drop_on(drop))
    ln_fwd_kernel<T, VPL, true><<<grid, LN_WARPS * 32, sizeof(float) * 2 * H, st>>>((const T*)x, (const T*)res, idx, gamma, beta,
                                                                                 (T*)y, mean, rstd, M, H, eps, *drop);
  else
    ln_fwd_kernel<T, VPL, false><<<grid, LN_WARPS * 32, sizeof(float) * 2 * H, st>>>((const T*)x, (const T*)res, idx, gamma, beta,
                                                                                  (T*)y, mean, rstd, M, H, eps, drop_or_off(nullptr));
  FCMF_LAUNCH_OK();
  return 0;
}

template <typename T, int VPL>
static int ln_bwd_launch(const void* dy, const void* dy_add, const void* x, const void* res, const int32_t* idx, const float* gamma,
                         const float* mean, const float* rstd, void* ds, void* dx, float* dgamma, float* dbeta, int64_t M, int H,
                         const fcmf_dropout* drop, int dy_every, cudaStream_t st) {
  int64_t blocks = (M + LN_WARPS - 1) / LN_WARPS;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  const size_t smem = sizeof(float) * (LN_WARPS * 2 + 1) * H;
  if (drop_on(drop))
    ln_bwd_kernel<T, VPL, true><<<(unsigned)blocks, LN_WARPS * 32, smem, st>>>(
        (const T*)dy, (const T*)dy_add, (const T*)x, (const T*)res, idx, gamma, mean, rstd, (T*)ds, (T*)dx, dgamma, dbeta, M, H, *drop, dy_every);
  else
    ln_bwd_kernel<T, VPL, false><<<(unsigned)blocks, LN_WARPS * 32, smem, st>>>(
        (const T*)dy, (const T*)dy_add, (const T*)x, (const T*)res, idx, gamma, mean, rstd, (T*)ds, (T*)dx, dgamma, dbeta, M, H,
        drop_or_off(nullptr), dy_every);
  FCMF_LAUNCH_OK();
  return 0;
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
