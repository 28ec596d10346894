// Internal GEMM engine interface (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fcmf {

// SIMT engine (gemm_simt.cu)
int gemm_simt_tn(const void* A, int64_t lda, const void* B, int64_t ldb, const float* bias, void* D, int64_t ldd,
                 void* aux, int64_t ldaux, int64_t M, int64_t N, int64_t K, int epi, int dtype, cudaStream_t st);
int gemm_simt_wgrad(const void* dY, int64_t lddy, const void* X, int64_t ldx, float* dW, int64_t M, int64_t N,
                    int64_t K, int accumulate, int dtype, cudaStream_t st);
int colsum(const void* dY, int64_t ld, float* db, int64_t M, int64_t N, int accumulate, int dtype, cudaStream_t st);

// tcgen05 engine (gemm_tc.cu), bf16 only
bool gemm_tc_supported_tn(int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldd, int64_t ldaux,
                          const void* A, const void* B, const void* D, const void* aux);
bool gemm_tc_supported_wgrad(int64_t M, int64_t N, int64_t K, int64_t lddy, int64_t ldx, const void* dY, const void* X);
int gemm_tc_tn(const void* A, int64_t lda, const void* B, int64_t ldb, const float* bias, void* D, int64_t ldd,
               void* aux, int64_t ldaux, int64_t M, int64_t N, int64_t K, int epi, cudaStream_t st);
int gemm_tc_wgrad(const void* dY, int64_t lddy, const void* X, int64_t ldx, float* dW, int64_t M, int64_t N,
                  int64_t K, int accumulate, cudaStream_t st);

int gemm_tc_tn_f32(const void* A, int64_t lda, const void* B, int64_t ldb, float* D, int64_t ldd, int64_t M, int64_t N,
                   int64_t K, cudaStream_t st);
void gemm_tc_wgrad_plan(int64_t M, int64_t N, int64_t K, int* pair, int* tiles, int* splits, int* workers);

}  // namespace fcmf
