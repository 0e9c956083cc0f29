// CUDA-core GEMM: C = epilogue(A W^T + bias), fp32 FMA accumulation in a fixed K order.
//
// This is the "fp32-scoring" parity path: tensor cores have no fp32 mode, so the
// mode whose kept-token sets must equal the reference's fp32 CPU run uses plain FFMA.  It also
// accepts bf16 operands (rounded inputs, fp32 accumulate) so that every tcgen05 kernel has an
// on-device cross-check with identical operand rounding.  Not the throughput path.
// Reference call sites: nn.Linear at audiomae/models_vit.py:41-45,76,96,522 and the patch-embed
// conv (models_vit.py:246) restated as a GEMM over the tpat_patchify output.
#include "gemm.cuh"

namespace tpat {

constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 16, SG_PAD = 4;

template <typename T> struct Vec4Load;
template <> struct Vec4Load<float> {
  static __device__ __forceinline__ void load(const float* p, float (&o)[4]) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
};
template <> struct Vec4Load<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&o)[4]) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&v.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&v.y);
    o[0] = __low2float(a); o[1] = __high2float(a); o[2] = __low2float(b); o[3] = __high2float(b);
  }
};

template <typename OutT>
__device__ __forceinline__ void epi_store(OutT* C, int ldc, int m, int n, float acc, const EpiParams& ep) {
  float v = acc + (ep.bias ? __ldg(ep.bias + n) : 0.f);
  size_t orow = (size_t)m;
  if (ep.epilogue == TPAT_EPI_BIAS_GELU) {
    if (ep.dact_out)   // training: keep the exact erf-GELU derivative Phi(h) + h phi(h) for the backward
      reinterpret_cast<OutT*>(ep.dact_out)[(size_t)m * ep.ld_dact + n] =
          from_f32<OutT>(0.5f * (1.0f + erff(v * 0.70710678118654752440f)) + v * 0.39894228040143267794f * expf(-0.5f * v * v));
    v = gelu_erf(v);
  } else if (ep.epilogue == TPAT_EPI_DGELU) {
    v *= to_f32<OutT>(reinterpret_cast<const OutT*>(ep.aux)[(size_t)m * ep.ld_aux + n]);
  } else if (ep.epilogue == TPAT_EPI_BIAS_RESIDUAL) {
    if (ep.row_scale) v *= __ldg(ep.row_scale + m / ep.rows_per_clip);
    v = ep.residual[(size_t)m * ep.ldr + n] + v;  // plain load: C may alias the residual
  } else if (ep.epilogue == TPAT_EPI_BIAS_POS) {
    const int b = m / ep.P, p = m - b * ep.P;
    orow = (size_t)b * (ep.num_extra + ep.P) + ep.num_extra + p;
    v = v + __ldg(ep.pos + (size_t)(ep.num_extra + p) * ldc + n);
  }
  C[orow * ldc + n] = from_f32<OutT>(v);
}

template <typename InT, typename OutT>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const InT* __restrict__ A, int lda, const InT* __restrict__ W, OutT* __restrict__ C, int ldc,
                 int M, int N, int K, EpiParams ep) {
  __shared__ float As[2][SG_BK][SG_BM + SG_PAD];
  __shared__ float Ws[2][SG_BK][SG_BN + SG_PAD];
  pdl_trigger();
  pdl_wait();
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  const int tx = tid & 15, ty = tid >> 4;
  // loader mapping: 128 rows x 4 k-vectors = 512 vec4, two per thread
  const int lr = tid >> 2, lk = (tid & 3) * 4;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float ra[2][4], rw[2][4];
  auto gload = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lr + 64 * h;
      if (m0 + r < M) Vec4Load<InT>::load(A + (size_t)(m0 + r) * lda + k0 + lk, ra[h]);
      else { ra[h][0] = ra[h][1] = ra[h][2] = ra[h][3] = 0.f; }
      if (n0 + r < N) Vec4Load<InT>::load(W + (size_t)(n0 + r) * K + k0 + lk, rw[h]);
      else { rw[h][0] = rw[h][1] = rw[h][2] = rw[h][3] = 0.f; }
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lr + 64 * h;
#pragma unroll
      for (int j = 0; j < 4; ++j) { As[buf][lk + j][r] = ra[h][j]; Ws[buf][lk + j][r] = rw[h][j]; }
    }
  };

  const int nk = K / SG_BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * SG_BK);
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      float a[8], w[8];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4 + 64]);
      const float4 w0 = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
      const float4 w1 = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4 + 64]);
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
      if (n < N) epi_store<OutT>(C, ldc, m, n, acc[i][j], ep);
    }
  }
}

int gemm_simt(const void* A, int a_dtype, int lda, const void* W, void* C, int c_dtype, int ldc, int M, int N, int K,
              const EpiParams& ep, cudaStream_t st) {
  TPAT_CHECK(K % SG_BK == 0 && lda % 4 == 0, "tpat_gemm(simt): need K %% 16 == 0 and lda %% 4 == 0 (K=%d lda=%d)", K, lda);
  TPAT_CHECK(aligned16(A) && aligned16(W) || a_dtype == TPAT_BF16, "tpat_gemm(simt): fp32 operands must be 16-byte aligned");
  dim3 grid((N + SG_BN - 1) / SG_BN, (M + SG_BM - 1) / SG_BM);
  if (a_dtype == TPAT_F32 && c_dtype == TPAT_F32)
    TPAT_CUDA(launch_kernel(gemm_simt_kernel<float, float>, dim3(grid), dim3(256), 0, st, (const float*)A, lda, (const float*)W, (float*)C, ldc, M, N, K, ep));
  else if (a_dtype == TPAT_F32 && c_dtype == TPAT_BF16)
    TPAT_CUDA(launch_kernel(gemm_simt_kernel<float, __nv_bfloat16>, dim3(grid), dim3(256), 0, st, (const float*)A, lda, (const float*)W, (__nv_bfloat16*)C, ldc, M, N, K, ep));
  else if (a_dtype == TPAT_BF16 && c_dtype == TPAT_F32)
    TPAT_CUDA(launch_kernel(gemm_simt_kernel<__nv_bfloat16, float>, dim3(grid), dim3(256), 0, st, (const __nv_bfloat16*)A, lda, (const __nv_bfloat16*)W, (float*)C, ldc, M, N, K, ep));
  else
    TPAT_CUDA(launch_kernel(gemm_simt_kernel<__nv_bfloat16, __nv_bfloat16>, dim3(grid), dim3(256), 0, st, (const __nv_bfloat16*)A, lda, (const __nv_bfloat16*)W, (__nv_bfloat16*)C, ldc, M, N, K, ep));
  TPAT_LAUNCH_CHECK();
  return 0;
}

}  // namespace tpat
