// Softmax cross-entropy over a wide class axis (the 250 002-entry XLM-R vocabulary of the IAOG decoder).
// Reference: torch.nn.CrossEntropyLoss(ignore_index=-100) on logits.permute(0, 2, 1) (run_pretraining_fcmf.py:320-322).
// HBM-bound row kernels: one CTA per (batch, position) row; forward = one read of the row (online max / sum in the log2
// domain), backward = one read + one write (may be in place). Rows are V elements long with V not a multiple of the
// 16-byte vector (250 002 = 2 mod 8), so a row starts at any 2/4-byte alignment: scalar head up to the first aligned
// address, vector body, scalar tail.
#include "common.cuh"

namespace fcmf {

constexpr int VCE_THREADS = 512;
constexpr float kVceLog2e = 1.44269504088896340736f;

// (max, sum of 2^(x*log2e - max)) pairs combine associatively
__device__ __forceinline__ void ms_combine(float& m, float& s, float m2, float s2) {
  const float mn = fmaxf(m, m2);
  s = (m == -INFINITY ? 0.f : s * exp2f(m - mn)) + (m2 == -INFINITY ? 0.f : s2 * exp2f(m2 - mn));
  m = mn;
}
__device__ __forceinline__ void ms_add(float& m, float& s, float x2) {     // x2 = logit * log2e
  if (x2 > m) { s = s * exp2f(m - x2) + 1.0f; m = x2; }                     // exp2f(-inf) = 0 covers the first element
  else s += exp2f(x2 - m);
}

template <typename T, typename F>
__device__ __forceinline__ void for_each_in_row(const T* row, int64_t V, F&& f) {
  constexpr int N = Vec16<T>::N;
  const int64_t mis = (reinterpret_cast<uintptr_t>(row) & 15u) / sizeof(T);
  int64_t head = mis ? (N - mis) : 0;
  if (head > V) head = V;
  for (int64_t j = threadIdx.x; j < head; j += blockDim.x) f(j, to_f(row[j]));
  const int64_t nvec = (V - head) / N;
  for (int64_t v = threadIdx.x; v < nvec; v += blockDim.x) {
    Vec16<T> a;
    a.load(row + head + v * N);
#pragma unroll
    for (int j = 0; j < N; ++j) f(head + v * N + j, a.v[j]);
  }
  for (int64_t j = head + nvec * N + threadIdx.x; j < V; j += blockDim.x) f(j, to_f(row[j]));
}

template <typename T>
__global__ void __launch_bounds__(VCE_THREADS)
vocab_ce_fwd_kernel(const T* __restrict__ logits, int64_t ld, const int64_t* __restrict__ labels, int64_t ignore_index,
                    float* __restrict__ loss_rows, float* __restrict__ lse, int64_t V) {
  __shared__ float sm_m[VCE_THREADS / 32], sm_s[VCE_THREADS / 32];
  const int64_t r = blockIdx.x;
  const T* row = logits + r * ld;
  float m = -INFINITY, s = 0.f;
  for_each_in_row<T>(row, V, [&](int64_t, float x) { ms_add(m, s, x * kVceLog2e); });
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    ms_combine(m, s, m2, s2);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sm_m[warp] = m; sm_s[warp] = s; }
  __syncthreads();
  if (warp == 0) {
    m = lane < VCE_THREADS / 32 ? sm_m[lane] : -INFINITY;
    s = lane < VCE_THREADS / 32 ? sm_s[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
      ms_combine(m, s, m2, s2);
    }
    if (lane == 0) {
      const float l = (m + log2f(s)) * 0.69314718055994530942f;       // natural-log log-sum-exp
      lse[r] = l;
      const int64_t y = labels[r];
      loss_rows[r] = (y == ignore_index || y < 0 || y >= V) ? 0.f : l - to_f(row[y]);
    }
  }
}

// dlogits[r, j] = (softmax(logits[r])[j] - [j == label[r]]) * scale[0]   (0 for ignored rows); dlogits may alias logits
template <typename T>
__global__ void __launch_bounds__(VCE_THREADS)
vocab_ce_bwd_kernel(const T* logits, int64_t ld, const int64_t* __restrict__ labels, int64_t ignore_index,
                    const float* __restrict__ lse, const float* __restrict__ scale, T* dlogits, int64_t ldd, int64_t V) {
  constexpr int N = Vec16<T>::N;
  const int64_t r = blockIdx.x;
  const T* row = logits + r * ld;
  T* out = dlogits + r * ldd;
  const int64_t y = labels[r];
  const bool valid = !(y == ignore_index || y < 0 || y >= V);
  const float sc = valid ? scale[0] : 0.f;
  const float l2 = lse[r] * kVceLog2e;
  auto grad = [&](int64_t j, float x) { return (exp2f(fmaf(x, kVceLog2e, -l2)) - (j == y ? 1.0f : 0.0f)) * sc; };
  // same head / body / tail split as the forward; the output row must share the input row's alignment phase for the
  // vector stores (the wrapper allocates dlogits with the logits' strides, or aliases it)
  const int64_t mis = (reinterpret_cast<uintptr_t>(row) & 15u) / sizeof(T);
  int64_t head = mis ? (N - mis) : 0;
  if (head > V) head = V;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(out) & 15u) == (reinterpret_cast<uintptr_t>(row) & 15u));
  for (int64_t j = threadIdx.x; j < head; j += blockDim.x) out[j] = from_f<T>(grad(j, to_f(row[j])));
  const int64_t nvec = (V - head) / N;
  for (int64_t v = threadIdx.x; v < nvec; v += blockDim.x) {
    Vec16<T> a, o;
    a.load(row + head + v * N);
#pragma unroll
    for (int j = 0; j < N; ++j) o.v[j] = grad(head + v * N + j, a.v[j]);
    if (vec_ok) o.store(out + head + v * N);
    else {
#pragma unroll
      for (int j = 0; j < N; ++j) out[head + v * N + j] = from_f<T>(o.v[j]);
    }
  }
  for (int64_t j = head + nvec * N + threadIdx.x; j < V; j += blockDim.x) out[j] = from_f<T>(grad(j, to_f(row[j])));
}

}  // namespace fcmf

using namespace fcmf;

extern "C" int fcmf_vocab_ce_fwd(const void* logits, int64_t ld, const int64_t* labels, int64_t ignore_index,
                                 float* loss_rows, float* lse, int64_t R, int64_t V, int dtype, void* stream) {
  FCMF_CHECK_ARG(R >= 0 && V > 0 && ld >= V, "vocab_ce_fwd: bad shape R=%lld V=%lld ld=%lld", (long long)R, (long long)V, (long long)ld);
  FCMF_CHECK_ARG(logits && labels && loss_rows && lse, "vocab_ce_fwd: null buffer");
  FCMF_CHECK_ARG(R < (1LL << 31), "vocab_ce_fwd: too many rows");
  if (R == 0) return 0;
  cudaStream_t st = as_stream(stream);
  if (dtype == FCMF_BF16) vocab_ce_fwd_kernel<bf16><<<(unsigned)R, VCE_THREADS, 0, st>>>((const bf16*)logits, ld, labels, ignore_index, loss_rows, lse, V);
  else if (dtype == FCMF_F32) vocab_ce_fwd_kernel<float><<<(unsigned)R, VCE_THREADS, 0, st>>>((const float*)logits, ld, labels, ignore_index, loss_rows, lse, V);
  else return fail(FCMF_ERR_ARG, "vocab_ce_fwd: bad dtype %d", dtype);
  FCMF_LAUNCH_OK();
  return 0;
}

extern "C" int fcmf_vocab_ce_bwd(const void* logits, int64_t ld, const int64_t* labels, int64_t ignore_index,
                                 const float* lse, const float* scale, void* dlogits, int64_t ldd, int64_t R, int64_t V,
                                 int dtype, void* stream) {
  FCMF_CHECK_ARG(R >= 0 && V > 0 && ld >= V && ldd >= V, "vocab_ce_bwd: bad shape");
  FCMF_CHECK_ARG(logits && labels && lse && scale && dlogits, "vocab_ce_bwd: null buffer");
  FCMF_CHECK_ARG(R < (1LL << 31), "vocab_ce_bwd: too many rows");
  if (R == 0) return 0;
  cudaStream_t st = as_stream(stream);
  if (dtype == FCMF_BF16) vocab_ce_bwd_kernel<bf16><<<(unsigned)R, VCE_THREADS, 0, st>>>((const bf16*)logits, ld, labels, ignore_index, lse, scale, (bf16*)dlogits, ldd, V);
  else if (dtype == FCMF_F32) vocab_ce_bwd_kernel<float><<<(unsigned)R, VCE_THREADS, 0, st>>>((const float*)logits, ld, labels, ignore_index, lse, scale, (float*)dlogits, ldd, V);
  else return fail(FCMF_ERR_ARG, "vocab_ce_bwd: bad dtype %d", dtype);
  FCMF_LAUNCH_OK();
  return 0;
}
