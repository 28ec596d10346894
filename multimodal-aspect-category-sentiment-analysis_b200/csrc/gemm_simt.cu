// SIMT (CUDA-core, fp32-accumulate) GEMM: the fp32 parity engine and the checker for the tcgen05 engine.
// Generic strides so one kernel serves forward (TN), input-gradient (TN on W^T) and weight-gradient (NT^T).
#include "common.cuh"
#include "gemm.h"

namespace fcmf {

constexpr int SBM = 64, SBN = 64, SBK = 16, STHREADS = 256;

// C[m,n] = sum_k A(m,k) * B(n,k);  A(m,k) = A[m*sam + k*sak], B(n,k) = B[n*sbn + k*sbk]
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(STHREADS)
gemm_simt_kernel(const TIn* __restrict__ A, int64_t sam, int64_t sak,
                 const TIn* __restrict__ B, int64_t sbn, int64_t sbk,
                 const float* __restrict__ bias, TOut* __restrict__ D, int64_t ldd,
                 TIn* __restrict__ aux, int64_t ldaux,
                 int64_t M, int64_t N, int64_t K, int epi, int accumulate) {
  __shared__ float As[SBK][SBM + 4];
  __shared__ float Bs[SBK][SBN + 4];
  const int t = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.y * SBM, n0 = (int64_t)blockIdx.x * SBN;
  const bool a_kcontig = (sak == 1), b_kcontig = (sbk == 1);
  // loader coordinates: 4 elements per thread per operand
  const int a_r = a_kcontig ? (t >> 2) : ((t & 15) << 2);
  const int a_k = a_kcontig ? ((t & 3) << 2) : (t >> 4);
  const int b_r = b_kcontig ? (t >> 2) : ((t & 15) << 2);
  const int b_k = b_kcontig ? ((t & 3) << 2) : (t >> 4);
  const int ty = t >> 4, tx = t & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = 0; k0 < K; k0 += SBK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int r = a_kcontig ? a_r : a_r + e, kk = a_kcontig ? a_k + e : a_k;
      const int64_t m = m0 + r, k = k0 + kk;
      As[kk][r] = (m < M && k < K) ? to_f(A[m * sam + k * sak]) : 0.f;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int r = b_kcontig ? b_r : b_r + e, kk = b_kcontig ? b_k + e : b_k;
      const int64_t n = n0 + r, k = k0 + kk;
      Bs[kk][r] = (n < N && k < K) ? to_f(B[n * sbn + k * sbk]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SBK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if (epi == FCMF_EPI_GELU) {
        if (aux) aux[m * ldaux + n] = from_f<TIn>(v);
        v = gelu_erf(v);
      } else if (epi == FCMF_EPI_TANH) {
        v = tanhf(v);
      } else if (epi == FCMF_EPI_DGELU) {
        v *= gelu_erf_grad(to_f(aux[m * ldaux + n]));
      }
      TOut* o = D + m * ldd + n;
      if (accumulate) v += to_f(*o);
      *o = from_f<TOut>(v);
    }
  }
}

// db[n] (+)= sum_m dY[m, n].  HBM-bound: each warp reads whole 512-byte (bf16) / 512-byte (f32) row segments with
// 16-byte vector loads, 8 warps of a block take interleaved rows, partial sums are combined in shared memory and one
// atomic per column per block goes to global.
constexpr int CS_WARPS = 8;
template <typename T>
__global__ void __launch_bounds__(CS_WARPS * 32)
colsum_kernel(const T* __restrict__ dY, int64_t ld, float* __restrict__ db, int64_t M, int64_t N, int64_t rows_per_block) {
  constexpr int V = Vec16<T>::N;
  __shared__ float red[CS_WARPS][32 * V];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n0 = ((int64_t)blockIdx.x * 32 + lane) * V;
  const int64_t m_lo = (int64_t)blockIdx.y * rows_per_block;
  const int64_t m_hi = m_lo + rows_per_block < M ? m_lo + rows_per_block : M;
  float acc[V];
#pragma unroll
  for (int j = 0; j < V; ++j) acc[j] = 0.f;
  if (n0 < N) {                                           // N % V == 0 is checked on the host
    int64_t m = m_lo + warp;
    for (; m + 3 * CS_WARPS < m_hi; m += 4 * CS_WARPS) {  // 4 independent 16-byte loads in flight per thread
      Vec16<T> v0, v1, v2, v3;
      v0.load(dY + m * ld + n0);
      v1.load(dY + (m + CS_WARPS) * ld + n0);
      v2.load(dY + (m + 2 * CS_WARPS) * ld + n0);
      v3.load(dY + (m + 3 * CS_WARPS) * ld + n0);
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] += (v0.v[j] + v1.v[j]) + (v2.v[j] + v3.v[j]);
    }
    for (; m < m_hi; m += CS_WARPS) {
      Vec16<T> v0;
      v0.load(dY + m * ld + n0);
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] += v0.v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < V; ++j) red[warp][lane * V + j] = acc[j];
  __syncthreads();
  for (int c = threadIdx.x; c < 32 * V; c += blockDim.x) {
    const int64_t n = (int64_t)blockIdx.x * 32 * V + c;
    if (n < N) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < CS_WARPS; ++w) s += red[w][c];
      atomicAdd(db + n, s);
    }
  }
}

// scalar fallback for column counts / alignments the vector kernel cannot take
template <typename T>
__global__ void colsum_scalar_kernel(const T* __restrict__ dY, int64_t ld, float* __restrict__ db, int64_t M, int64_t N,
                                     int64_t rows_per_block) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int64_t m_lo = (int64_t)blockIdx.y * rows_per_block;
  const int64_t m_hi = m_lo + rows_per_block < M ? m_lo + rows_per_block : M;
  float s = 0.f;
  for (int64_t m = m_lo; m < m_hi; ++m) s += to_f(dY[m * ld + n]);
  atomicAdd(db + n, s);
}

template <typename T>
int colsum_launch(const T* dY, int64_t ld, float* db, int64_t M, int64_t N, int accumulate, cudaStream_t st) {
  constexpr int V = Vec16<T>::N;
  if (!accumulate) FCMF_CUDA_OK(cudaMemsetAsync(db, 0, sizeof(float) * N, st));
  if (M == 0) return 0;
  const bool vec = (N % V == 0) && (ld % V == 0) && ((reinterpret_cast<uintptr_t>(dY) & 15u) == 0);
  const int64_t col_blocks = vec ? (N + 32 * V - 1) / (32 * V) : (N + 127) / 128;
  int64_t want_y = (4LL * sm_count() + col_blocks - 1) / col_blocks;       // ~4 blocks per SM in total
  if (want_y < 1) want_y = 1;
  int64_t rows = (M + want_y - 1) / want_y;
  if (rows < 64) rows = 64;
  dim3 grid((unsigned)col_blocks, (unsigned)((M + rows - 1) / rows));
  if (vec) colsum_kernel<T><<<grid, CS_WARPS * 32, 0, st>>>(dY, ld, db, M, N, rows);
  else colsum_scalar_kernel<T><<<grid, 128, 0, st>>>(dY, ld, db, M, N, rows);
  FCMF_LAUNCH_OK();
  return 0;
}

int colsum(const void* dY, int64_t ld, float* db, int64_t M, int64_t N, int accumulate, int dtype, cudaStream_t st) {
  return dtype == FCMF_BF16 ? colsum_launch((const bf16*)dY, ld, db, M, N, accumulate, st)
                            : colsum_launch((const float*)dY, ld, db, M, N, accumulate, st);
}

template <typename T>
static int simt_tn(const void* A, int64_t lda, const void* B, int64_t ldb, const float* bias, void* D, int64_t ldd,
                   void* aux, int64_t ldaux, int64_t M, int64_t N, int64_t K, int epi, cudaStream_t st) {
  dim3 grid((unsigned)((N + SBN - 1) / SBN), (unsigned)((M + SBM - 1) / SBM));
  gemm_simt_kernel<T, T><<<grid, STHREADS, 0, st>>>((const T*)A, lda, 1, (const T*)B, ldb, 1, bias, (T*)D, ldd,
                                                    (T*)aux, ldaux, M, N, K, epi, 0);
  FCMF_LAUNCH_OK();
  return 0;
}

int gemm_simt_tn(const void* A, int64_t lda, const void* B, int64_t ldb, const float* bias, void* D, int64_t ldd,
                 void* aux, int64_t ldaux, int64_t M, int64_t N, int64_t K, int epi, int dtype, cudaStream_t st) {
  if (M == 0 || N == 0) return 0;
  return dtype == FCMF_BF16 ? simt_tn<bf16>(A, lda, B, ldb, bias, D, ldd, aux, ldaux, M, N, K, epi, st)
                            : simt_tn<float>(A, lda, B, ldb, bias, D, ldd, aux, ldaux, M, N, K, epi, st);
}

template <typename T>
static int simt_wgrad(const void* dY, int64_t lddy, const void* X, int64_t ldx, float* dW, int64_t M, int64_t N,
                      int64_t K, int accumulate, cudaStream_t st) {
  // dW[n,k] = sum_m dY[m,n] X[m,k]: "A"(n,m) = dY[m*lddy + n], "B"(k,m) = X[m*ldx + k]; reduction length M.
  dim3 grid((unsigned)((K + SBN - 1) / SBN), (unsigned)((N + SBM - 1) / SBM));
  gemm_simt_kernel<T, float><<<grid, STHREADS, 0, st>>>((const T*)dY, 1, lddy, (const T*)X, 1, ldx, nullptr, dW, K,
                                                        nullptr, 0, N, K, M, FCMF_EPI_NONE, accumulate);
  FCMF_LAUNCH_OK();
  return 0;
}

int gemm_simt_wgrad(const void* dY, int64_t lddy, const void* X, int64_t ldx, float* dW, int64_t M, int64_t N,
                    int64_t K, int accumulate, int dtype, cudaStream_t st) {
  if (N == 0 || K == 0) return 0;
  return dtype == FCMF_BF16 ? simt_wgrad<bf16>(dY, lddy, X, ldx, dW, M, N, K, accumulate, st)
                            : simt_wgrad<float>(dY, lddy, X, ldx, dW, M, N, K, accumulate, st);
}

}  // namespace fcmf

using namespace fcmf;

extern "C" int fcmf_gemm_tn(const void* A, int64_t lda, const void* B, int64_t ldb, const float* bias, void* D,
                            int64_t ldd, void* aux, int64_t ldaux, int64_t M, int64_t N, int64_t K, int epi,
                            int dtype, int engine, void* stream) {
  FCMF_CHECK_ARG(M >= 0 && N >= 0 && K > 0, "gemm_tn: bad shape M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
  FCMF_CHECK_ARG(dtype == FCMF_F32 || dtype == FCMF_BF16, "gemm_tn: bad dtype %d", dtype);
  FCMF_CHECK_ARG(epi >= FCMF_EPI_NONE && epi <= FCMF_EPI_DGELU, "gemm_tn: bad epilogue %d", epi);
  FCMF_CHECK_ARG(epi != FCMF_EPI_DGELU || aux != nullptr, "gemm_tn: EPI_DGELU needs aux");
  FCMF_CHECK_ARG(lda >= K && ldb >= K && ldd >= N, "gemm_tn: leading dimension too small");
  if (M == 0 || N == 0) return 0;
  cudaStream_t st = as_stream(stream);
  const bool tc_ok = dtype == FCMF_BF16 && gemm_tc_supported_tn(M, N, K, lda, ldb, ldd, ldaux, A, B, D, aux);
  if (engine == FCMF_ENGINE_TCGEN05 && !tc_ok)
    return fail(FCMF_ERR_UNSUPPORTED, "gemm_tn: tcgen05 engine cannot run M=%lld N=%lld K=%lld dtype=%d",
                (long long)M, (long long)N, (long long)K, dtype);
  if (engine == FCMF_ENGINE_TCGEN05 || (engine == FCMF_ENGINE_AUTO && tc_ok))
    return gemm_tc_tn(A, lda, B, ldb, bias, D, ldd, aux, ldaux, M, N, K, epi, st);
  return gemm_simt_tn(A, lda, B, ldb, bias, D, ldd, aux, ldaux, M, N, K, epi, dtype, st);
}

namespace {
// one side stream + fork/join events per (thread, device): calls from different host threads never share them
struct SideStream { cudaStream_t stream = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };
SideStream* side_stream() {
  static thread_local SideStream per_dev[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  SideStream& s = per_dev[dev];
  if (!s.stream) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;            // never create objects while a capture is running
    if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) { s.stream = nullptr; return nullptr; }
    (void)cap;
    if (cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  }
  return (s.stream && s.fork && s.join) ? &s : nullptr;
}
}  // namespace

extern "C" int fcmf_gemm_tn_f32(const void* A, int64_t lda, const void* B, int64_t ldb, float* D, int64_t ldd, int64_t M,
                                int64_t N, int64_t K, int dtype, void* stream) {
  FCMF_CHECK_ARG(M >= 0 && N >= 0 && K > 0, "gemm_tn_f32: bad shape M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
  FCMF_CHECK_ARG(lda >= K && ldb >= K && ldd >= N, "gemm_tn_f32: leading dimension too small");
  if (M == 0 || N == 0) return 0;
  if (dtype != FCMF_BF16 || !gemm_tc_supported_tn(M, N, K, lda, ldb, 8, 0, A, B, nullptr, nullptr) || (reinterpret_cast<uintptr_t>(D) & 15u))
    return fail(FCMF_ERR_UNSUPPORTED, "gemm_tn_f32: needs bf16 operands with 16-byte aligned rows (tcgen05 engine only)");
  return gemm_tc_tn_f32(A, lda, B, ldb, D, ldd, M, N, K, as_stream(stream));
}

extern "C" int fcmf_gemm_wgrad(const void* dY, int64_t lddy, const void* X, int64_t ldx, float* dW, float* db,
                               int64_t M, int64_t N, int64_t K, int accumulate, int dtype, int engine, void* stream) {
  FCMF_CHECK_ARG(M >= 0 && N > 0 && K > 0, "gemm_wgrad: bad shape");
  FCMF_CHECK_ARG(dtype == FCMF_F32 || dtype == FCMF_BF16, "gemm_wgrad: bad dtype %d", dtype);
  FCMF_CHECK_ARG(lddy >= N && ldx >= K, "gemm_wgrad: leading dimension too small");
  cudaStream_t st = as_stream(stream);
  const bool tc_ok = dtype == FCMF_BF16 && gemm_tc_supported_wgrad(M, N, K, lddy, ldx, dY, X);
  if (engine == FCMF_ENGINE_TCGEN05 && !tc_ok)
    return fail(FCMF_ERR_UNSUPPORTED, "gemm_wgrad: tcgen05 engine cannot run M=%lld N=%lld K=%lld dtype=%d",
                (long long)M, (long long)N, (long long)K, dtype);
  const bool tc = engine == FCMF_ENGINE_TCGEN05 || (engine == FCMF_ENGINE_AUTO && tc_ok);
  // The bias gradient (column sums of dY: an HBM-bound pass, 1.3 ms per config-2 step in total) runs on a side stream UNDER
  // the tensor-bound weight-gradient GEMM of the same dY (fork / join with events; capturable into a CUDA graph).
  SideStream* side = (db && tc && M >= 8192) ? side_stream() : nullptr;
  if (db && !side) {
    int r = colsum(dY, lddy, db, M, N, accumulate, dtype, st);
    if (r) return r;
  }
  if (side) FCMF_CUDA_OK(cudaEventRecord(side->fork, st));
  // the GEMM is launched FIRST: its persistent CTAs (one per SM, ~200 KB of shared memory) become resident and the column-sum
  // blocks fill what is left; launched the other way round the small blocks occupy the SMs and the GEMM CTAs wait for them
  const int rc = tc ? gemm_tc_wgrad(dY, lddy, X, ldx, dW, M, N, K, accumulate, st)
                    : gemm_simt_wgrad(dY, lddy, X, ldx, dW, M, N, K, accumulate, dtype, st);
  if (side) {
    FCMF_CUDA_OK(cudaStreamWaitEvent(side->stream, side->fork, 0));
    int r = colsum(dY, lddy, db, M, N, accumulate, dtype, side->stream);
    if (r) return r;
    FCMF_CUDA_OK(cudaEventRecord(side->join, side->stream));
    FCMF_CUDA_OK(cudaStreamWaitEvent(st, side->join, 0));
  }
  return rc;
}

// Host-only planning query: how the tcgen05 weight-gradient GEMM would tile and split this shape (no launch, no GPU needed).
extern "C" int fcmf_gemm_wgrad_plan(int64_t M, int64_t N, int64_t K, int32_t* cta_pairs, int32_t* tiles, int32_t* splits,
                                    int32_t* workers) {
  FCMF_CHECK_ARG(M > 0 && N > 0 && K > 0 && cta_pairs && tiles && splits && workers, "gemm_wgrad_plan: bad arguments");
  int p = 0, t = 0, s = 0, w = 0;
  gemm_tc_wgrad_plan(M, N, K, &p, &t, &s, &w);
  *cta_pairs = p; *tiles = t; *splits = s; *workers = w;
  return 0;
}
