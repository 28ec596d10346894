// How fast can a B200 WRITE to HBM, and does the pattern matter? (nvcc -arch=sm_100a -O3 -o write_bw_probe write_bw_probe.cu)
// 1) cudaMemsetAsync; 2) streaming 16-byte stores, fully coalesced; 3) the GEMM epilogue's pattern: a CTA writes a
// [128 rows x SEG bytes] box into a row-major [M, 6144-byte] matrix (SEG = 128 is what one TMA slab store does),
// boxes taken in the tile order of the GEMM; 4) a copy for reference. Sizes match the FFN activation (456 960 x 3072 bf16).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void stream_store(uint4* p, size_t n) {
  const uint4 v = make_uint4(1, 2, 3, 4);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void stream_store_cs(uint4* p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    asm volatile("st.global.cs.v4.u32 [%0], {%1, %1, %1, %1};" ::"l"(p + i), "r"(7u) : "memory");
}
__global__ void stream_copy(const uint4* __restrict__ s, uint4* __restrict__ d, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] = s[i];
}
// box pattern: work item = (m_blk of 128 rows, column segment of SEG bytes); items ordered column-fastest inside a
// 256-column-tile (512 bytes), tiles n-fastest -- the order the persistent GEMM emits its slabs.
template <int SEG>
__global__ void box_store(uint8_t* p, int64_t rows, int row_bytes) {
  const int segs = row_bytes / SEG;
  const int64_t items = (rows / 128) * segs;
  const uint4 v = make_uint4(1, 2, 3, 4);
  for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
    const int64_t mb = it / segs; const int sg = (int)(it % segs);
    uint8_t* base = p + mb * 128 * (int64_t)row_bytes + (int64_t)sg * SEG;
    for (int e = threadIdx.x; e < 128 * (SEG / 16); e += blockDim.x) {
      const int r = e / (SEG / 16), c = e % (SEG / 16);
      *reinterpret_cast<uint4*>(base + (int64_t)r * row_bytes + c * 16) = v;
    }
  }
}

template <typename F> float timed(F f, int reps = 5) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  float best = 1e9f;
  for (int i = 0; i < reps; ++i) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
  return best;
}

int main() {
  const int64_t rows = 456960; const int row_bytes = 6144;
  const size_t bytes = (size_t)rows * row_bytes;
  uint8_t *d, *s;
  CK(cudaMalloc(&d, bytes)); CK(cudaMalloc(&s, bytes));
  CK(cudaMemset(s, 1, bytes));
  int sms = 148;
  auto rep = [&](const char* name, float ms, double b) { printf("%8.3f ms %8.1f GB/s  %s\n", ms, b / ms / 1e6, name); };
  rep("cudaMemsetAsync", timed([&] { cudaMemsetAsync(d, 0, bytes); }), bytes);
  rep("16-byte stores, coalesced, grid 148x8x256", timed([&] { stream_store<<<sms * 8, 256>>>((uint4*)d, bytes / 16); }), bytes);
  rep("16-byte stores, coalesced, grid 148x32x256", timed([&] { stream_store<<<sms * 32, 256>>>((uint4*)d, bytes / 16); }), bytes);
  rep("st.global.cs 16-byte stores", timed([&] { stream_store_cs<<<sms * 8, 256>>>((uint4*)d, bytes / 16); }), bytes);
  rep("box [128 rows x 128 B] (one TMA slab)", timed([&] { box_store<128><<<sms * 4, 256>>>(d, rows, row_bytes); }), bytes);
  rep("box [128 rows x 256 B]", timed([&] { box_store<256><<<sms * 4, 256>>>(d, rows, row_bytes); }), bytes);
  rep("box [128 rows x 512 B] (a 256-column tile)", timed([&] { box_store<512><<<sms * 4, 256>>>(d, rows, row_bytes); }), bytes);
  rep("box [128 rows x 2048 B]", timed([&] { box_store<2048><<<sms * 4, 256>>>(d, rows, row_bytes); }), bytes);
  rep("copy (read + write bytes)", timed([&] { stream_copy<<<sms * 16, 256>>>((const uint4*)s, (uint4*)d, bytes / 16); }), 2.0 * bytes);
  rep("cudaMemcpyAsync d2d (read + write bytes)", timed([&] { cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToDevice); }), 2.0 * bytes);
  CK(cudaDeviceSynchronize());
  return 0;
}
