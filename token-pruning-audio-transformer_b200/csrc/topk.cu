// Score reduction + descending top-k selection, one CTA per clip.
//
// Replaces `.mean(...)` over the attention matrix slice and torch.topk(score, k, largest=True,
// sorted=True) (reference audiomae/models_vit.py:113-114, ast/src/models/ast_models.py:124-125).
// The partials written by the attention kernel are summed in a fixed order (row 0, 1, ... R-1), so
// the result is bit-reproducible run to run.  Selection is a full bitonic sort of 64-bit keys
// (monotone score bits << 32 | ~index) in shared memory: descending score, ties broken towards
// the LOWER index (torch leaves tie order unspecified); NaN sorts above +inf like
// torch.topk.  HBM traffic per clip: 4*R*N bytes of partials read, 4*(N-extra) + 8*k written.
#include "common.cuh"

namespace tpat {

constexpr int TK_THREADS = 512;

__device__ __forceinline__ uint32_t order_bits(float s) {
  if (s != s) return 0xFFFFFFFFu;  // NaN ranks highest
  const uint32_t u = __float_as_uint(s);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(TK_THREADS)
score_topk_kernel(const float* __restrict__ partial, int R, float divisor, float* __restrict__ score,
                  int64_t* __restrict__ topk_idx, int32_t* __restrict__ rest_idx, int N, int num_extra, int k, int npad) {
  extern __shared__ unsigned long long keys[];  // [npad]
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x;
  const int n = N - num_extra;
  const float* pb = partial + (size_t)b * R * N;
  for (int j = threadIdx.x; j < npad; j += TK_THREADS) {
    unsigned long long key = 0ull;  // padding sorts last
    if (j < n) {
      float s = 0.f;
      for (int r = 0; r < R; ++r) s += __ldg(pb + (size_t)r * N + num_extra + j);
      s = s / divisor;
      if (score != nullptr) score[(size_t)b * n + j] = s;
      key = ((unsigned long long)order_bits(s) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)j);
    }
    keys[j] = key;
  }
  if (topk_idx == nullptr || k <= 0) return;
  __syncthreads();
  // bitonic sort, descending
  for (int size = 2; size <= npad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (npad >> 1); t += TK_THREADS) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned long long a = keys[lo], c = keys[hi];
        if ((a < c) == desc) { keys[lo] = c; keys[hi] = a; }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < k; i += TK_THREADS)
    topk_idx[(size_t)b * k + i] = (int64_t)(0xFFFFFFFFu - (uint32_t)(keys[i] & 0xFFFFFFFFull));
  if (rest_idx != nullptr)   // the tokens that were NOT kept, still in descending-score order (EViT fused-token input)
    for (int i = k + threadIdx.x; i < n; i += TK_THREADS)
      rest_idx[(size_t)b * (n - k) + (i - k)] = (int32_t)(0xFFFFFFFFu - (uint32_t)(keys[i] & 0xFFFFFFFFull));
}

}  // namespace tpat

extern "C" int tpat_score_topk(const float* partial, int R, float divisor, float* score, int64_t* topk_idx,
                               int32_t* rest_idx, int B, int N, int num_extra, int k, tpat_stream_t stream) {
  using namespace tpat;
  TPAT_CHECK(partial != nullptr, "tpat_score_topk: null partial");
  const int n = N - num_extra;
  TPAT_CHECK(B >= 0 && R > 0 && n > 0 && num_extra >= 0, "tpat_score_topk: bad sizes R=%d N=%d extra=%d", R, N, num_extra);
  TPAT_CHECK(k >= 0 && k <= n, "tpat_score_topk: k=%d out of range for %d candidate tokens", k, n);
  TPAT_CHECK(n <= 8192, "tpat_score_topk: at most 8192 candidate tokens per clip (got %d)", n);
  TPAT_CHECK(divisor != 0.f, "tpat_score_topk: zero divisor");
  TPAT_CHECK(rest_idx == nullptr || (topk_idx != nullptr && k > 0), "tpat_score_topk: rest_idx needs topk_idx and k > 0");
  if (B == 0) return 0;
  int npad = 2;
  while (npad < n) npad <<= 1;
  const size_t smem = (size_t)npad * sizeof(unsigned long long);
  if (smem > 48 * 1024) TPAT_CUDA(cudaFuncSetAttribute(score_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  TPAT_CUDA(launch_kernel(score_topk_kernel, dim3(B), dim3(TK_THREADS), smem, as_stream(stream), partial, R, divisor, score, topk_idx, rest_idx, N, num_extra, k, npad));
  TPAT_LAUNCH_CHECK();
  return 0;
}
