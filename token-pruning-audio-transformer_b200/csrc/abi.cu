// Error plumbing and the trivial entry points of the C-ABI (include/tpat.h).
#include "common.cuh"

#include <cstdlib>
#include <mutex>

namespace tpat {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  return 2;
}

thread_local int g_walk_desc = 0;

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

thread_local L2Window g_l2_window;

size_t l2_persist_bytes() {
  static long long want_mb = -1;
  static size_t granted[64] = {0};
  static bool asked[64] = {false};
  if (want_mb < 0) { const char* e = getenv("TPAT_L2_PERSIST_MB"); want_mb = e ? atoll(e) : 0; }
  if (want_mb <= 0) return 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  if (!asked[dev]) {
    int max_persist = 0, max_window = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
    size_t want = (size_t)want_mb << 20;
    if (want > (size_t)max_persist) want = (size_t)max_persist;
    size_t got = 0;
    if (want > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) cudaDeviceGetLimit(&got, cudaLimitPersistingL2CacheSize);
    cudaGetLastError();
    granted[dev] = got;
    asked[dev] = true;
    if (getenv("TPAT_DEBUG"))
      fprintf(stderr, "[tpat] L2 persisting carve-out: asked %lld MiB, device max %d B, window max %d B, granted %zu B\n", want_mb, max_persist, max_window, got);
  }
  return granted[dev];
}

bool pdl_enabled() {
  static int cached = -1;
  if (cached < 0) { const char* e = getenv("TPAT_PDL"); cached = (e != nullptr && e[0] == '1') ? 1 : 0; }
  return cached == 1;
}

}  // namespace tpat

extern "C" {

int tpat_version(void) { return TPAT_VERSION; }

size_t tpat_sizeof_forward_args(void) { return sizeof(tpat_forward_args); }

const char* tpat_last_error(void) { return tpat::g_err; }

int tpat_device_ok(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return major == 10 ? 1 : 0;
}

}  // extern "C"
