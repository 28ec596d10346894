// Optimizer tail kernels (GPU-validated in round 2: tests/test_gpu_optim.py).
// Optimizer tail of the training step (reference run_multimodal_fcmf.py:483-489):
//   torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0) ; optimizer.step() with torch.optim.AdamW over 4 parameter
//   groups -- ~100 small tensors, one foreach pass each plus a host synchronisation for the norm.
// Here: a table of tensors in device memory, ONE launch for the global squared norm, one tiny launch for the clip
// coefficient (stays on the device: no .item()), ONE launch for clip + AdamW over every tensor.
#include "common.cuh"

namespace fcmf {

struct OptTensor {          // mirrored by optim.py (ctypes): keep the field order
  float* p; const float* g; float* m; float* v;
  int64_t n;
  float lr, wd;
};
constexpr int OPT_CHUNK = 8192;          // elements per block
constexpr int OPT_THREADS = 256;

__global__ void __launch_bounds__(OPT_THREADS)
opt_sumsq_kernel(const OptTensor* __restrict__ tab, const int32_t* __restrict__ blk_tensor, const int32_t* __restrict__ blk_chunk,
                 float* __restrict__ sumsq) {
  const OptTensor t = tab[blk_tensor[blockIdx.x]];
  const int64_t i0 = (int64_t)blk_chunk[blockIdx.x] * OPT_CHUNK;
  const int64_t i1 = min(t.n, i0 + OPT_CHUNK);
  float s = 0.f;
  for (int64_t i = i0 + threadIdx.x; i < i1; i += OPT_THREADS) { const float g = t.g[i]; s = fmaf(g, g, s); }
  s = warp_sum(s);
  __shared__ float part[OPT_THREADS / 32];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < OPT_THREADS / 32 ? part[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0) atomicAdd(sumsq, s);
  }
}

// coef = min(1, max_norm / (sqrt(sumsq) + 1e-6))   (torch.nn.utils.clip_grad_norm_); norm_out = sqrt(sumsq)
__global__ void opt_clip_coef_kernel(const float* __restrict__ sumsq, float max_norm, float* __restrict__ coef, float* __restrict__ norm_out) {
  const float nrm = sqrtf(sumsq[0]);
  if (norm_out) norm_out[0] = nrm;
  coef[0] = max_norm > 0.f ? fminf(1.0f, max_norm / (nrm + 1e-6f)) : 1.0f;
}

// torch.optim.AdamW single-tensor update (decoupled weight decay), gradient scaled by the clip coefficient first
__global__ void __launch_bounds__(OPT_THREADS)
opt_adamw_kernel(const OptTensor* __restrict__ tab, const int32_t* __restrict__ blk_tensor, const int32_t* __restrict__ blk_chunk,
                 const float* __restrict__ coef, float beta1, float beta2, float eps, float bc1, float rsqrt_bc2, int write_back_grad) {
  const OptTensor t = tab[blk_tensor[blockIdx.x]];
  const int64_t i0 = (int64_t)blk_chunk[blockIdx.x] * OPT_CHUNK;
  const int64_t i1 = min(t.n, i0 + OPT_CHUNK);
  const float c = coef ? coef[0] : 1.0f;
  const float step_size = t.lr / bc1;
  const float decay = 1.0f - t.lr * t.wd;
  for (int64_t i = i0 + threadIdx.x; i < i1; i += OPT_THREADS) {
    const float g = t.g[i] * c;
    const float m = fmaf(1.0f - beta1, g - t.m[i], t.m[i]);          // exp_avg.lerp_(grad, 1 - beta1)
    const float v = fmaf(t.v[i], beta2, (1.0f - beta2) * g * g);
    const float denom = sqrtf(v) * rsqrt_bc2 + eps;
    t.m[i] = m;
    t.v[i] = v;
    t.p[i] = t.p[i] * decay - step_size * (m / denom);
    if (write_back_grad) const_cast<float*>(t.g)[i] = g;            // leave the clipped gradient visible, as clip_grad_norm_ does
  }
}

}  // namespace fcmf

using namespace fcmf;

extern "C" int fcmf_opt_sumsq(const void* table, const int32_t* blk_tensor, const int32_t* blk_chunk, int64_t n_blocks,
                              float* sumsq, void* stream) {
  FCMF_CHECK_ARG(table && blk_tensor && blk_chunk && sumsq && n_blocks >= 0 && n_blocks < (1LL << 31), "opt_sumsq: bad arguments");
  cudaStream_t st = as_stream(stream);
  FCMF_CUDA_OK(cudaMemsetAsync(sumsq, 0, sizeof(float), st));
  if (n_blocks == 0) return 0;
  opt_sumsq_kernel<<<(unsigned)n_blocks, OPT_THREADS, 0, st>>>((const OptTensor*)table, blk_tensor, blk_chunk, sumsq);
  FCMF_LAUNCH_OK();
  return 0;
}

extern "C" int fcmf_opt_clip_coef(const float* sumsq, float max_norm, float* coef, float* norm_out, void* stream) {
  FCMF_CHECK_ARG(sumsq && coef, "opt_clip_coef: null buffer");
  opt_clip_coef_kernel<<<1, 1, 0, as_stream(stream)>>>(sumsq, max_norm, coef, norm_out);
  FCMF_LAUNCH_OK();
  return 0;
}

extern "C" int fcmf_opt_adamw(const void* table, const int32_t* blk_tensor, const int32_t* blk_chunk, int64_t n_blocks,
                              const float* coef, float beta1, float beta2, float eps, int64_t step, int write_back_grad,
                              void* stream) {
  FCMF_CHECK_ARG(table && blk_tensor && blk_chunk && n_blocks >= 0 && n_blocks < (1LL << 31) && step >= 1, "opt_adamw: bad arguments");
  if (n_blocks == 0) return 0;
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  opt_adamw_kernel<<<(unsigned)n_blocks, OPT_THREADS, 0, as_stream(stream)>>>(
      (const OptTensor*)table, blk_tensor, blk_chunk, coef, beta1, beta2, eps, (float)bc1, (float)(1.0 / sqrt(bc2)), write_back_grad);
  FCMF_LAUNCH_OK();
  return 0;
}
