// Shared host/device helpers for libtpat.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/tpat.h"

namespace tpat {

// ---- error plumbing: thread-local message, int status across the C boundary ----
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define TPAT_CHECK(cond, ...)                       \
  do {                                              \
    if (!(cond)) {                                  \
      ::tpat::set_error(__VA_ARGS__);               \
      return 1;                                     \
    }                                               \
  } while (0)

#define TPAT_CUDA(expr)                                                        \
  do {                                                                         \
    cudaError_t _e = (expr);                                                   \
    if (_e != cudaSuccess) return ::tpat::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define TPAT_LAUNCH_CHECK() TPAT_CUDA(cudaPeekAtLastError())

static inline cudaStream_t as_stream(tpat_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline size_t dtype_size(int dt) { return dt == TPAT_BF16 ? 2 : 4; }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int sm_count();  // cached multiprocessor count of the current device

// Per-device once-flags for cudaFuncSetAttribute (function attributes belong to the device's context: a process that
// drives several GPUs, e.g. the reference's nn.DataParallel replicas, must set them on each).  `first()` only READS the
// flag; the caller sets the attribute and then calls `mark()`, so a second host thread racing on the same device sets
// the (idempotent) attribute again instead of launching before it is in place.
struct DeviceOnce {
  volatile bool done[64] = {};
  static int dev() { int d = 0; cudaGetDevice(&d); return d & 63; }
  bool first() const { return !done[dev()]; }
  void mark() { done[dev()] = true; }
};
bool pdl_enabled();  // programmatic dependent launch when TPAT_PDL=1 (measured r01: 11.18k vs 11.46k clips/s -> off by default)

// Walk direction of the next launches made by this thread (library-internal, set by tpat_forward only; the per-kernel
// entry points default to ascending).  1 = the kernel visits its rows / row tiles / clips from the END.  A kernel
// that starts on the rows its producer wrote last finds them still in the 126 MB L2, and leaves the other end of its
// own output hot for the next kernel, so tpat_forward alternates the direction along the producer -> consumer chain.
extern thread_local int g_walk_desc;

// Every kernel CAN be launched with programmatic stream serialization (TPAT_PDL=1): the next kernel's CTAs may be
// scheduled (and run their prologue: barrier init, TMEM allocation, descriptor prefetch) while the tail of the
// previous kernel drains; each kernel calls pdl_wait() before its first access to global memory and
// pdl_trigger() at its start (both are no-ops without the launch attribute).
// L2 residency of the fp32 residual stream (TPAT_L2_PERSIST_MB=<MiB>, off by default): tpat_forward describes the live
// residual rows of the current block as an access-policy window; every launch inside carries it, so that lines of that
// range are kept in the persisting carve-out of the 126 MB L2 across the 4 kernels per block that touch them
// (LayerNorm reads x twice, proj and fc2 read-modify-write it) instead of being evicted by the activations in between.
struct L2Window { const void* base = nullptr; size_t bytes = 0; float ratio = 0.f; };
extern thread_local L2Window g_l2_window;
size_t l2_persist_bytes();   // carve-out actually granted on the current device (0 = feature off)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (g_l2_window.bytes != 0) {
    at[1].id = cudaLaunchAttributeAccessPolicyWindow;
    at[1].val.accessPolicyWindow.base_ptr = const_cast<void*>(g_l2_window.base);
    at[1].val.accessPolicyWindow.num_bytes = g_l2_window.bytes;
    at[1].val.accessPolicyWindow.hitRatio = g_l2_window.ratio;
    at[1].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    at[1].val.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    cfg.numAttrs = 2;
  }
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---- device helpers ----
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);  // .x = lo (low 16 bits), .y = hi
  return *reinterpret_cast<uint32_t*>(&v);
}

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

// exact (erf) GELU, the nn.GELU() default the reference uses (models_vit.py:31)
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

}  // namespace tpat
