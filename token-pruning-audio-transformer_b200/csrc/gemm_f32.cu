// General fp32 CUDA-core GEMM with either operand transposed: C (+)= op(A) op(B).
//
// The backward of nn.Linear (y = x W^T + b; reference call sites audiomae/models_vit.py:41-45,76,96,246,522 under
// autograd) needs two more operand layouts than the forward's "A [M,K] times W [N,K] transposed":
//   dgrad  dX [M, in]  = dY [M, out] . W [out, in]          (A as is, B as is)
//   wgrad  dW [out, in] += dY^T [out, M] . X [M, in]         (A transposed, B as is; reduction over the tokens)
// This kernel is the fp32 parity path for both (fixed summation order, deterministic) and also serves the tiny
// classifier-head GEMMs in every precision mode.  Same 128 x 128 x 16 register-tiled core as gemm_simt.cu.
#include "common.cuh"

namespace tpat {

constexpr int GF_BM = 128, GF_BN = 128, GF_BK = 16, GF_PAD = 4;

// element (r, c) of a row-major matrix with guards; rows x cols logical extent
__device__ __forceinline__ float gf_at(const float* p, int ld, int r, int c, int rows, int cols) {
  return (r < rows && c < cols) ? __ldg(p + (size_t)r * ld + c) : 0.f;
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
gemm_f32_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, float* __restrict__ C, int ldc,
                int M, int N, int K, int accumulate) {
  __shared__ float As[2][GF_BK][GF_BM + GF_PAD];
  __shared__ float Bs[2][GF_BK][GF_BN + GF_PAD];
  pdl_trigger();
  pdl_wait();
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * GF_BM, n0 = blockIdx.x * GF_BN;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float ra[8], rb[8];
  // loader: 128 x 16 elements per operand per k-tile = 2048 = 8 per thread.
  //   operand stored [rows = M or N][K]  (not TA / TB): thread -> (row = tid >> 1 (+0), k = (tid & 1) * 8 .. +7)
  //   operand stored [K][rows]           (TA / not TB): thread -> (k = tid >> 4, rows = (tid & 15) * 8 .. +7)
  auto gload = [&](int k0) {
    if (!TA) { const int r = tid >> 1, kk = (tid & 1) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) ra[j] = gf_at(A, lda, m0 + r, k0 + kk + j, M, K);
    } else { const int kk = tid >> 4, r = (tid & 15) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) ra[j] = gf_at(A, lda, k0 + kk, m0 + r + j, K, M);
    }
    if (TB) { const int r = tid >> 1, kk = (tid & 1) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) rb[j] = gf_at(B, ldb, n0 + r, k0 + kk + j, N, K);
    } else { const int kk = tid >> 4, r = (tid & 15) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) rb[j] = gf_at(B, ldb, k0 + kk, n0 + r + j, K, N);
    }
  };
  auto sstore = [&](int buf) {
    if (!TA) { const int r = tid >> 1, kk = (tid & 1) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) As[buf][kk + j][r] = ra[j];
    } else { const int kk = tid >> 4, r = (tid & 15) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) As[buf][kk][r + j] = ra[j];
    }
    if (TB) { const int r = tid >> 1, kk = (tid & 1) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) Bs[buf][kk + j][r] = rb[j];
    } else { const int kk = tid >> 4, r = (tid & 15) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) Bs[buf][kk][r + j] = rb[j];
    }
  };

  const int nk = (K + GF_BK - 1) / GF_BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * GF_BK);
#pragma unroll
    for (int k = 0; k < GF_BK; ++k) {
      float a[8], w[8];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4 + 64]);
      const float4 w0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 w1 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4 + 64]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w; w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 4 + (i & 3) + (i >> 2) * 64;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + tx * 4 + (j & 3) + (j >> 2) * 64;
      if (n >= N) continue;
      float* c = C + (size_t)m * ldc + n;
      *c = accumulate ? *c + acc[i][j] : acc[i][j];
    }
  }
}

}  // namespace tpat

extern "C" int tpat_gemm_f32(const float* A, int lda, int trans_a, const float* B, int ldb, int trans_b, float* C, int ldc,
                             int M, int N, int K, int accumulate, tpat_stream_t stream) {
  using namespace tpat;
  TPAT_CHECK(A && B && C, "tpat_gemm_f32: null pointer");
  TPAT_CHECK(M >= 0 && N >= 0 && K >= 0, "tpat_gemm_f32: bad sizes M=%d N=%d K=%d", M, N, K);
  TPAT_CHECK(lda >= (trans_a ? M : K) && ldb >= (trans_b ? K : N) && ldc >= N, "tpat_gemm_f32: leading dimension too small");
  if (M == 0 || N == 0) return 0;
  const dim3 grid((N + GF_BN - 1) / GF_BN, (M + GF_BM - 1) / GF_BM);
  cudaStream_t st = as_stream(stream);
  if (!trans_a && !trans_b) TPAT_CUDA(launch_kernel(gemm_f32_kernel<false, false>, dim3(grid), dim3(256), 0, st, A, lda, B, ldb, C, ldc, M, N, K, accumulate));
  else if (!trans_a && trans_b) TPAT_CUDA(launch_kernel(gemm_f32_kernel<false, true>, dim3(grid), dim3(256), 0, st, A, lda, B, ldb, C, ldc, M, N, K, accumulate));
  else if (trans_a && !trans_b) TPAT_CUDA(launch_kernel(gemm_f32_kernel<true, false>, dim3(grid), dim3(256), 0, st, A, lda, B, ldb, C, ldc, M, N, K, accumulate));
  else TPAT_CUDA(launch_kernel(gemm_f32_kernel<true, true>, dim3(grid), dim3(256), 0, st, A, lda, B, ldb, C, ldc, M, N, K, accumulate));
  TPAT_LAUNCH_CHECK();
  return 0;
}
