// Geometric ROI relations: the trigonometric box-pair embedding (fp64, as the reference data path) and the
// per-head geometry weight  log(max(relu(WG_h . emb + b_h), 1e-6))  that is added to the box-attention scores.
// Reference: BoxRelationalEmbedding (roi_modeling.py:79-138), WGs/relu (roi_modeling.py:160-162), log/clamp (:40).
// Tiny, latency-bound (G*NR*NR pairs): one warp per box pair, no shared memory.
#include "common.cuh"

namespace fcmf {

struct Freq8 { float f[8]; };

__global__ void box_geometry_fwd_kernel(const double* __restrict__ boxes, const float* __restrict__ wg_w,
                                        const float* __restrict__ wg_b, float* __restrict__ emb,
                                        float* __restrict__ bias, int64_t G, int NR, int heads, Freq8 fr) {
  const int lane = threadIdx.x & 31;
  const int64_t pair = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pair >= G * NR * NR) return;
  const int64_t g = pair / (NR * NR);
  const int ij = (int)(pair - g * NR * NR), i = ij / NR, j = ij - i * NR;
  const double* bi = boxes + (g * NR + i) * 4;           // (x_min, x_max, y_min, y_max)  roi_modeling.py:95
  const double* bj = boxes + (g * NR + j) * 4;
  const double cxi = (bi[0] + bi[1]) * 0.5, cyi = (bi[2] + bi[3]) * 0.5;
  const double wi = (bi[1] - bi[0]) + 1.0, hi = (bi[3] - bi[2]) + 1.0;
  const double cxj = (bj[0] + bj[1]) * 0.5, cyj = (bj[2] + bj[3]) * 0.5;
  const double wj = (bj[1] - bj[0]) + 1.0, hj = (bj[3] - bj[2]) + 1.0;
  const int comp = lane >> 3, k = lane & 7;
  double pos;
  if (comp == 0) pos = log(fmax(fabs((cxi - cxj) / wi), 1e-3));
  else if (comp == 1) pos = log(fmax(fabs((cyi - cyj) / hi), 1e-3));
  else if (comp == 2) pos = log(wi / wj);
  else pos = log(hi / hj);
  const double arg = (100.0 * pos) * (double)fr.f[k];     // f64 position x f32 frequency, roi_modeling.py:126-130
  const float es = (float)sin(arg), ec = (float)cos(arg); // cast to the activation dtype, roi_modeling.py:149
  float* e = emb + pair * 64;
  e[lane] = es;
  e[32 + lane] = ec;
  for (int h = 0; h < heads; ++h) {
    float z = wg_w[h * 64 + lane] * es + wg_w[h * 64 + 32 + lane] * ec;
    z = warp_sum(z) + wg_b[h];
    if (lane == 0) bias[((g * heads + h) * NR + i) * NR + j] = logf(fmaxf(fmaxf(z, 0.f), 1e-6f));
  }
}

// dz[g,h,i,j] = dbias / z  where z >= 1e-6 (relu and clamp both pass), else 0
__global__ void box_geometry_dz_kernel(const float* __restrict__ emb, const float* __restrict__ wg_w,
                                       const float* __restrict__ wg_b, const float* __restrict__ dbias,
                                       float* __restrict__ dz, int64_t G, int NR, int heads) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= G * heads * NR * NR) return;
  const int64_t g = t / ((int64_t)heads * NR * NR);
  const int rem = (int)(t - g * heads * NR * NR), h = rem / (NR * NR), ij = rem - h * NR * NR;
  const float* e = emb + (g * NR * NR + ij) * 64;
  float z = wg_b[h];
  for (int c = 0; c < 64; ++c) z = fmaf(wg_w[h * 64 + c], e[c], z);
  dz[t] = (z >= 1e-6f) ? dbias[t] / z : 0.f;
}

// d_w[h][c] += sum_pairs dz * emb[c] ; d_b[h] += sum_pairs dz.   grid = (heads, chunks of pairs), block = 64
__global__ void box_geometry_dw_kernel(const float* __restrict__ emb, const float* __restrict__ dz,
                                       float* __restrict__ d_w, float* __restrict__ d_b, int64_t G, int NR, int heads,
                                       int64_t pairs_per_block) {
  const int h = blockIdx.x, c = threadIdx.x;
  const int64_t npair = G * NR * NR;
  const int64_t lo = (int64_t)blockIdx.y * pairs_per_block;
  const int64_t hi = lo + pairs_per_block < npair ? lo + pairs_per_block : npair;
  float acc = 0.f, accb = 0.f;
  for (int64_t pr = lo; pr < hi; ++pr) {
    const int64_t g = pr / (NR * NR);
    const int ij = (int)(pr - g * NR * NR);
    const float d = dz[(g * heads + h) * NR * NR + ij];
    acc = fmaf(d, emb[pr * 64 + c], acc);
    accb += d;
  }
  atomicAdd(d_w + h * 64 + c, acc);
  if (c == 0) atomicAdd(d_b + h, accb);
}

}  // namespace fcmf

using namespace fcmf;

extern "C" int fcmf_box_geometry_fwd(const double* boxes, const float* wg_w, const float* wg_b, const float* freq8_host,
                                     float* emb, float* bias, int64_t G, int32_t NR, int32_t heads, void* stream) {
  FCMF_CHECK_ARG(G >= 0 && NR > 0 && heads > 0 && freq8_host, "box_geometry_fwd: bad arguments");
  if (G == 0) return 0;
  Freq8 fr;
  for (int i = 0; i < 8; ++i) fr.f[i] = freq8_host[i];
  const int64_t pairs = G * NR * NR;
  box_geometry_fwd_kernel<<<(unsigned)((pairs + 3) / 4), 128, 0, as_stream(stream)>>>(boxes, wg_w, wg_b, emb, bias, G, NR, heads, fr);
  FCMF_LAUNCH_OK();
  return 0;
}

extern "C" int fcmf_box_geometry_bwd(const float* emb, const float* wg_w, const float* wg_b, const float* dbias,
                                     float* dz_ws, float* d_wg_w, float* d_wg_b, int64_t G, int32_t NR, int32_t heads,
                                     void* stream) {
  FCMF_CHECK_ARG(G >= 0 && NR > 0 && heads > 0 && dz_ws, "box_geometry_bwd: bad arguments");
  if (G == 0) return 0;
  const int64_t n = G * heads * NR * NR;
  cudaStream_t st = as_stream(stream);
  box_geometry_dz_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(emb, wg_w, wg_b, dbias, dz_ws, G, NR, heads);
  FCMF_LAUNCH_OK();
  const int64_t npair = G * NR * NR;
  int64_t chunks = (npair + 63) / 64;
  if (chunks > 256) chunks = 256;
  const int64_t per = (npair + chunks - 1) / chunks;
  box_geometry_dw_kernel<<<dim3((unsigned)heads, (unsigned)((npair + per - 1) / per)), 64, 0, st>>>(emb, dz_ws, d_wg_w, d_wg_b, G, NR, heads, per);
  FCMF_LAUNCH_OK();
  return 0;
}
