// Probe (run on B200 in round 2, result in profiles/r02a_tma_swizzle_probe.txt: ADDRESS based): does a SWIZZLE_128B TMA load whose shared-memory destination is 128-byte but
// NOT 1024-byte aligned swizzle by the ABSOLUTE shared address (chunk ^ ((addr >> 7) & 7)) or relative to the box start?
// The attention kernels want to TMA-load a second row segment (the 4 ROI rows) behind 170 text rows of the same tile, i.e.
// at tile row 170 = byte offset 21760 = 128-aligned, not 1024-aligned. If the swizzle is address based (hypothesis A) the
// rows simply continue the tile's pattern and the UMMA descriptors need no change.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_probe tools/tma_swizzle_probe.cu && /tmp/tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap map, int row_offset, int src_row, uint16_t* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 8192);
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) reinterpret_cast<uint16_t*>(sm)[i] = 0xFFFF;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(8 * 128) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(s32(sm + row_offset * 128)), "l"(reinterpret_cast<uint64_t>(&map)), "r"(s32(bar)), "r"(0), "r"(src_row) : "memory");
  }
  uint32_t ok = 0;
  long long t0 = clock64();
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(s32(bar)) : "memory");
    if (clock64() - t0 > 2000000000LL) { if (threadIdx.x == 0) printf("probe: TMA never completed\n"); break; }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(sm)[i];
}

int main() {
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q) != cudaSuccess || !fnp) { printf("no encode fn\n"); return 1; }
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fnp);
  const int R = 64, Ccols = 64;
  std::vector<uint16_t> h(R * Ccols);
  for (int r = 0; r < R; ++r) for (int c = 0; c < Ccols; ++c) h[r * Ccols + c] = (uint16_t)(r * 64 + c);
  uint16_t *g, *o;
  cudaMalloc(&g, h.size() * 2); cudaMalloc(&o, 4096 * 2);
  cudaMemcpy(g, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap map;
  cuuint64_t dims[2] = {(cuuint64_t)Ccols, (cuuint64_t)R};
  cuuint64_t strides[1] = {(cuuint64_t)Ccols * 2};
  cuuint32_t box[2] = {64, 8}, estr[2] = {1, 1};
  CUresult rc = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, g, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) { printf("encode failed %d\n", (int)rc); return 1; }
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  for (int row_offset : {0, 2, 10, 3}) {
    const int src_row = 16;
    probe<<<1, 128, 16384>>>(map, row_offset, src_row, o);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("row_offset %d: kernel error %s\n", row_offset, cudaGetErrorString(e)); return 1; }
    std::vector<uint16_t> out(4096);
    cudaMemcpy(out.data(), o, 8192, cudaMemcpyDeviceToHost);
    int okA = 1, okB = 1;
    for (int i = 0; i < 8; ++i) for (int c = 0; c < 8; ++c) {
      const int r = row_offset + i;
      const uint16_t want = (uint16_t)((src_row + i) * 64 + c * 8);           // first element of chunk c of source row
      const uint16_t atA = out[(r * 128 + ((c ^ (r & 7)) << 4)) / 2];          // swizzle by absolute tile row (address based)
      const uint16_t atB = out[(r * 128 + ((c ^ (i & 7)) << 4)) / 2];          // swizzle by row inside the box
      okA &= (atA == want); okB &= (atB == want);
    }
    printf("dst tile row %2d (byte offset %5d): address-based swizzle %s, box-relative swizzle %s\n", row_offset, row_offset * 128,
           okA ? "MATCHES" : "no", okB ? "MATCHES" : "no");
  }
  return 0;
}
