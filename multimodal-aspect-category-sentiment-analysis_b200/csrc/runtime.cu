// Error plumbing, device queries and the ABI version entry points.
#include "common.cuh"
#include <stdarg.h>
#include <stddef.h>
#include <atomic>
#include <mutex>

namespace fcmf {

std::string& last_error_ref() {
  static thread_local std::string msg;
  return msg;
}

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error_ref() = buf;
  return code;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launches() { return g_launches.load(std::memory_order_relaxed); }

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace fcmf

namespace fcmf { long long launches(); }
extern "C" int fcmf_abi_version(void) { return FCMF_ABI_VERSION; }
extern "C" long long fcmf_kernel_launches(void) { return fcmf::launches(); }

extern "C" const char* fcmf_last_error(void) { return fcmf::last_error_ref().c_str(); }

extern "C" int fcmf_device_info(int* sm, int* major, int* minor) {
  int dev = 0;
  FCMF_CUDA_OK(cudaGetDevice(&dev));
  cudaDeviceProp p;
  FCMF_CUDA_OK(cudaGetDeviceProperties(&p, dev));
  if (sm) *sm = p.multiProcessorCount;
  if (major) *major = p.major;
  if (minor) *minor = p.minor;
  return 0;
}

// Host evaluation of the dropout mask the kernels regenerate on the device (same inline functions, common.cuh).
extern "C" int fcmf_dropout_keep(float p, uint64_t seed, uint64_t row, uint32_t col) {
  const uint32_t thr = fcmf::drop_threshold(p);
  return fcmf::drop_keep(fcmf::drop_rowseed(seed, row), col, thr) ? 1 : 0;
}

// Layout audit for foreign-function bindings: sizeof / field offsets of the by-value structs of the ABI, so that a ctypes /
// cgo / JNI mirror can be checked against the compiled library without a GPU (tests/test_cpu_host.py does).
extern "C" int fcmf_abi_layout(int which) {
  switch (which) {
    case 0: return (int)sizeof(fcmf_dropout);
    case 1: return (int)offsetof(fcmf_dropout, seed);
    case 2: return (int)offsetof(fcmf_dropout, seed_dev);
    case 3: return (int)sizeof(fcmf_seg);
    case 4: return (int)offsetof(fcmf_seg, idx);
    case 5: return (int)sizeof(fcmf_attn_desc);
    case 6: return (int)offsetof(fcmf_attn_desc, mask_add);
    case 7: return (int)offsetof(fcmf_attn_desc, bias);
    case 8: return (int)offsetof(fcmf_attn_desc, scale);
    case 9: return (int)offsetof(fcmf_attn_desc, causal);
    case 10: return (int)offsetof(fcmf_attn_desc, drop);
    case 11: return (int)offsetof(fcmf_attn_desc, engine);
    case 12: return (int)offsetof(fcmf_seg, groups);
    default: return -1;
  }
}
