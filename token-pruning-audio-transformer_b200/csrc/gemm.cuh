// Internal interface shared by the GEMM translation units and the forward engine.
#pragma once
#include "common.cuh"

namespace tpat {

struct EpiParams {
  const float* bias;
  const float* residual; int ldr;
  const float* pos; int P; int num_extra;
  int epilogue;
  // LayerNorm fold (tpat_gemm_ln): producer outputs / consumer inputs, all optional
  void* xb = nullptr; int ldxb = 0;
  float* part_out = nullptr; int part_ld = 0;
  const float* ln_part = nullptr; int ln_chunks = 0;
  const float* ln_colsum = nullptr;
  float ln_eps = 0.f;
  // training extras (tpat_gemm_train): all optional
  void* dact_out = nullptr; int ld_dact = 0;            // BIAS_GELU: also store gelu'(acc + bias), dtype of C
  const void* aux = nullptr; int ld_aux = 0;          // DGELU: the saved gelu'(h) (dtype of C): C = acc * aux
  const float* row_scale = nullptr; int rows_per_clip = 0;   // BIAS_RESIDUAL: C = R + row_scale[m / rows_per_clip] * (acc + bias)  (DropPath)
  int w_kn = 0;                                         // W stored [K, N] (MN-major B operand)
  float* cs_part = nullptr;                             // DGELU: per-32-row partial column sums of C, [ceil(M / 32) rounded to tiles][N]
};

// dst[c] += sum over the nparts rows of partials[nparts][C], fixed order (backward_rows.cu)
int finish_colsum_partials(const float* partials, int nparts, int C, float* dst, cudaStream_t st);

// CUDA-core fp32-FMA path (gemm_simt.cu)
int gemm_simt(const void* A, int a_dtype, int lda, const void* W, void* C, int c_dtype, int ldc, int M, int N, int K,
              const EpiParams& ep, cudaStream_t st);
// tcgen05 / TMEM / TMA path, bf16 operands (gemm_tc.cu)
int gemm_tc(const void* A, int lda, const void* W, void* C, int c_dtype, int ldc, int M, int N, int K,
            const EpiParams& ep, cudaStream_t st);

// CTA-pair (cta_group::2) variant, 256 x 256 tiles (gemm_tc2.cu); arguments validated by gemm_tc
int gemm_tc2(const void* A, int lda, const void* W, void* C, int c_dtype, int ldc, int M, int N, int K,
             const EpiParams& ep, cudaStream_t st);

}  // namespace tpat
