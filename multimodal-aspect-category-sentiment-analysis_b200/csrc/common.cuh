// Shared device/host helpers for the fcmf_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/fcmf_b200.h"

namespace fcmf {

// ---- error plumbing (thread-local message, negative return codes) --------------------------------------
std::string& last_error_ref();
int fail(int code, const char* fmt, ...);

#define FCMF_CHECK_ARG(cond, ...) \
  do { if (!(cond)) return ::fcmf::fail(FCMF_ERR_ARG, __VA_ARGS__); } while (0)

#define FCMF_CUDA_OK(expr)                                                                          \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return ::fcmf::fail(FCMF_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),    \
                          __FILE__, __LINE__);                                                      \
  } while (0)

// every kernel launch of this library goes through this macro, so the counter is the number of OUR kernels launched
void count_launch();
#define FCMF_LAUNCH_OK()                    \
  do {                                      \
    ::fcmf::count_launch();                 \
    FCMF_CUDA_OK(cudaGetLastError());       \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
int sm_count();

// ---- scalar conversions -----------------------------------------------------------------------------
typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// 8-wide (bf16) / 4-wide (f32) 16-byte vector access ----------------------------------------------------
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 4;
  float v[4];
  __device__ __forceinline__ void load(const float* p) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec16<bf16> {
  static constexpr int N = 8;
  float v[8];
  __device__ __forceinline__ void load(const bf16* p) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  __device__ __forceinline__ void store(bf16* p) const {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = t;
  }
};

// ---- warp / block reductions ----------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- dropout: stateless counter-based masks ---------------------------------------------------------------
// keep(row, col) is a pure function of (seed, row, col): forward and backward kernels regenerate the same mask instead
// of storing it (no mask tensor ever touches HBM). Per row one 32-bit row seed (two rounds of an integer finaliser over
// the 64-bit row index and seed); per PAIR of adjacent columns one more round, 16 bits per element, so p is quantised to
// 1/65536. The effective seed is seed + *seed_dev (seed_dev may be NULL): a device-resident counter lets a replayed
// CUDA graph draw fresh masks. tests/_util.py restates these three functions in numpy (the oracle side of the mask).
struct DropCfg {
  uint32_t thr16;        // drop when the element's 16-bit hash < thr16; 0 = dropout off
  float inv_keep;        // 1 / (1 - thr16/65536)
  uint64_t seed;
};
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {          // "lowbias32" integer finaliser
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
// The effective seed goes through a 64-bit finaliser (splitmix64) first: seeds that differ in a few low bits -- a captured
// CUDA graph advances the device counter by one per replay -- would otherwise give masks that are row permutations of
// each other (rowseed(seed + n, r) == rowseed(seed, r ^ lo(seed) ^ lo(seed + n)) for the plain xor).
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint32_t drop_rowseed(uint64_t seed, uint64_t row) {
  seed = splitmix64(seed);
  const uint32_t a = mix32((uint32_t)row ^ (uint32_t)seed);
  return mix32(a ^ ((uint32_t)(row >> 32) * 0x9E3779B9U + (uint32_t)(seed >> 32)));
}
// 32 hash bits of the column pair (col & ~1, col | 1): low half-word = even column, high half-word = odd column
__host__ __device__ __forceinline__ uint32_t drop_pair(uint32_t rowseed, uint32_t col) { return mix32(rowseed + (col >> 1)); }
__host__ __device__ __forceinline__ bool drop_keep_lo(uint32_t h, uint32_t thr16) { return (h & 0xffffU) >= thr16; }
__host__ __device__ __forceinline__ bool drop_keep_hi(uint32_t h, uint32_t thr16) { return (h >> 16) >= thr16; }
__host__ __device__ __forceinline__ bool drop_keep(uint32_t rowseed, uint32_t col, uint32_t thr16) {
  const uint32_t h = drop_pair(rowseed, col);
  return (col & 1) ? drop_keep_hi(h, thr16) : drop_keep_lo(h, thr16);
}
__host__ __device__ __forceinline__ uint32_t drop_threshold(float p) {     // p <= 0 => 0 (off)
  return (uint32_t)((p > 0.f ? p : 0.f) * 65536.0f + 0.5f);
}
__device__ __forceinline__ DropCfg make_drop(const fcmf_dropout& d) {
  DropCfg c;
  c.thr16 = drop_threshold(d.p);
  c.inv_keep = 1.0f / (1.0f - (float)c.thr16 * (1.0f / 65536.0f));
  c.seed = d.seed + (d.seed_dev ? *d.seed_dev : 0ULL);
  return c;
}
inline bool drop_on(const fcmf_dropout* d) { return d != nullptr && d->p > 0.f; }
inline int drop_check(const fcmf_dropout* d) { return (d == nullptr || (d->p >= 0.f && d->p < 1.0f)) ? 0 : -1; }
inline fcmf_dropout drop_or_off(const fcmf_dropout* d) { return d ? *d : fcmf_dropout{0.f, 0ULL, nullptr}; }

// ---- math ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_erf(float x) {            // mm_modeling.py:15
  return x * 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_erf_grad(float x) {       // d/dx [x * Phi(x)] = Phi(x) + x * phi(x)
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// Epilogue-rate versions for the bf16 tensor-core GEMM. The erf-GELU epilogues are bound by the MUFU pipe, not by issue
// slots or DRAM (ncu: two MUFU per element -- RCP + EX2 of an Abramowitz-Stegun erf -- cost ~8 k cycles per 128 x 256
// tile against 6.5 k cycles of MMA, tensor pipe 50 % active). One MUFU per element instead:
//     Phi(x) = 0.5 (1 + erf(x / sqrt 2))  ~=  0.5 (1 + tanh(x (c0 + c1 x^2 + c2 x^4)))
// c fitted (minimax-weighted least squares over |x| <= 8, tools/fit_gelu.py) to the erf-GELU AND its derivative:
// max |x Phi - fit| = 5.4e-5, max |d/dx - fit'| = 1.7e-4 -- below the bf16 rounding of the stored result (2^-9 relative)
// over the range where the activation is not itself rounded to +-0 / x; MUFU.TANH adds <= 2^-11 relative on tanh.
// x^2 is clamped at 64 (tanh has saturated to +-1 in fp32 long before; the quintic turns over at |x| = 8.06).
// The derivative is the derivative OF THE FIT (all FMA, no second exponential), so forward and backward stay consistent.
// fp32 parity mode never comes here: gemm_simt.cu uses erff (gelu_erf / gelu_erf_grad above).
__device__ __forceinline__ float tanh_approx(float x) {          // MUFU.TANH
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
constexpr float kGeluC0 = 0.7972415169105289f, kGeluC1 = 0.0372098378211574f, kGeluC2 = -0.0003810452660279267f;
__device__ __forceinline__ float gelu_erf_fast(float x) {       // x * Phi(x): 6 FP32 ops + 1 MUFU
  const float x2 = fminf(x * x, 64.0f);
  const float t = tanh_approx(x * fmaf(fmaf(kGeluC2, x2, kGeluC1), x2, kGeluC0));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}
// the same two functions on PAIRS (FFMA2 / FMUL2 / FADD2: one issue slot per two elements; identical fp32 results)
__device__ __forceinline__ float2 gelu_erf_fast2(float2 x) {
  const float2 xx = __fmul2_rn(x, x);
  const float2 x2 = make_float2(fminf(xx.x, 64.0f), fminf(xx.y, 64.0f));
  const float2 p = __ffma2_rn(__ffma2_rn(make_float2(kGeluC2, kGeluC2), x2, make_float2(kGeluC1, kGeluC1)), x2, make_float2(kGeluC0, kGeluC0));
  const float2 u = __fmul2_rn(x, p);
  const float2 t = make_float2(tanh_approx(u.x), tanh_approx(u.y));
  const float2 hx = __fmul2_rn(x, make_float2(0.5f, 0.5f));
  return __ffma2_rn(hx, t, hx);
}
__device__ __forceinline__ float2 gelu_erf_grad_fast2(float2 x) {
  const float2 xx = __fmul2_rn(x, x);
  const float2 x2 = make_float2(fminf(xx.x, 64.0f), fminf(xx.y, 64.0f));
  const float2 p = __ffma2_rn(__ffma2_rn(make_float2(kGeluC2, kGeluC2), x2, make_float2(kGeluC1, kGeluC1)), x2, make_float2(kGeluC0, kGeluC0));
  const float2 u = __fmul2_rn(x, p);
  const float2 t = make_float2(tanh_approx(u.x), tanh_approx(u.y));
  const float2 du = __ffma2_rn(__ffma2_rn(make_float2(5.0f * kGeluC2, 5.0f * kGeluC2), x2, make_float2(3.0f * kGeluC1, 3.0f * kGeluC1)), x2,
                               make_float2(kGeluC0, kGeluC0));
  const float2 sech2 = __ffma2_rn(make_float2(-t.x, -t.y), t, make_float2(1.0f, 1.0f));
  const float2 a = __ffma2_rn(make_float2(0.5f, 0.5f), t, make_float2(0.5f, 0.5f));
  return __ffma2_rn(__fmul2_rn(__fmul2_rn(x, make_float2(0.5f, 0.5f)), sech2), du, a);
}
__device__ __forceinline__ float gelu_erf_grad_fast(float x) {  // Phi(x) + x * phi(x) as the derivative of the fit: 12 FP32 ops + 1 MUFU
  const float x2 = fminf(x * x, 64.0f);
  const float t = tanh_approx(x * fmaf(fmaf(kGeluC2, x2, kGeluC1), x2, kGeluC0));
  const float du = fmaf(fmaf(5.0f * kGeluC2, x2, 3.0f * kGeluC1), x2, kGeluC0);
  const float sech2 = fmaf(-t, t, 1.0f);
  return fmaf(0.5f * x * sech2, du, fmaf(0.5f, t, 0.5f));
}

}  // namespace fcmf
