// tpat_gemm: argument validation and dispatch to the CUDA-core or tcgen05 implementation.
#include "gemm.cuh"

extern "C" int tpat_gemm(const void* A, int a_dtype, int lda, const void* W, int w_dtype, const float* bias,
                         void* C, int c_dtype, int ldc, const float* residual, int ldr, const float* pos,
                         int P, int num_extra, int M, int N, int K, int epilogue, int impl, tpat_stream_t stream) {
  return tpat_gemm_ln(A, a_dtype, lda, W, w_dtype, bias, C, c_dtype, ldc, residual, ldr, pos, P, num_extra, M, N, K, epilogue,
                      impl, nullptr, stream);
}

static int gemm_entry(const void* A, int a_dtype, int lda, const void* W, int w_dtype, const float* bias,
                      void* C, int c_dtype, int ldc, const float* residual, int ldr, const float* pos,
                      int P, int num_extra, int M, int N, int K, int epilogue, int impl, const tpat_ln_fold* fold,
                      const tpat_gemm_extra* extra, tpat_stream_t stream);

extern "C" int tpat_gemm_ln(const void* A, int a_dtype, int lda, const void* W, int w_dtype, const float* bias,
                            void* C, int c_dtype, int ldc, const float* residual, int ldr, const float* pos,
                            int P, int num_extra, int M, int N, int K, int epilogue, int impl, const tpat_ln_fold* fold,
                            tpat_stream_t stream) {
  return gemm_entry(A, a_dtype, lda, W, w_dtype, bias, C, c_dtype, ldc, residual, ldr, pos, P, num_extra, M, N, K, epilogue, impl, fold,
                    nullptr, stream);
}

/* partial rows: one per 32 output rows, rounded up to whole 256-row tiles (the 1-CTA kernel's 128-row tiles fit inside) */
extern "C" size_t tpat_gemm_colsum_ws_floats(int M, int N) { return (size_t)((M + 255) / 256 * 8) * (size_t)N; }

extern "C" int tpat_gemm_train(const void* A, int a_dtype, int lda, const void* W, int w_dtype, const float* bias,
                               void* C, int c_dtype, int ldc, const float* residual, int ldr, int M, int N, int K,
                               int epilogue, int impl, const tpat_gemm_extra* extra, tpat_stream_t stream) {
  return gemm_entry(A, a_dtype, lda, W, w_dtype, bias, C, c_dtype, ldc, residual, ldr, nullptr, 0, 0, M, N, K, epilogue, impl, nullptr,
                    extra, stream);
}

static int gemm_entry(const void* A, int a_dtype, int lda, const void* W, int w_dtype, const float* bias,
                      void* C, int c_dtype, int ldc, const float* residual, int ldr, const float* pos,
                      int P, int num_extra, int M, int N, int K, int epilogue, int impl, const tpat_ln_fold* fold,
                      const tpat_gemm_extra* extra, tpat_stream_t stream) {
  using namespace tpat;
  TPAT_CHECK(A && W && C, "tpat_gemm: null pointer");
  TPAT_CHECK(M >= 0 && N > 0 && K > 0, "tpat_gemm: bad sizes M=%d N=%d K=%d", M, N, K);
  TPAT_CHECK(a_dtype == w_dtype && (a_dtype == TPAT_F32 || a_dtype == TPAT_BF16), "tpat_gemm: A and W must share a dtype (f32 or bf16)");
  TPAT_CHECK(c_dtype == TPAT_F32 || c_dtype == TPAT_BF16, "tpat_gemm: bad C dtype %d", c_dtype);
  TPAT_CHECK(epilogue >= TPAT_EPI_BIAS && epilogue <= TPAT_EPI_DGELU, "tpat_gemm: bad epilogue %d", epilogue);
  if (epilogue == TPAT_EPI_DGELU)
    TPAT_CHECK(extra && extra->aux && extra->ld_aux >= N && bias == nullptr && aligned16(extra->aux) && (extra->ld_aux * dtype_size(c_dtype)) % 16 == 0,
               "tpat_gemm_train: the DGELU epilogue needs extra->aux (dtype of C, 16-byte aligned rows) and no bias");
  TPAT_CHECK(lda >= K && ldc >= N, "tpat_gemm: lda/ldc too small");
  if (epilogue == TPAT_EPI_BIAS_RESIDUAL) TPAT_CHECK(residual && ldr >= N && c_dtype == TPAT_F32, "tpat_gemm: residual epilogue needs residual, ldr >= N and fp32 C");
  if (epilogue == TPAT_EPI_BIAS_POS) TPAT_CHECK(pos && P > 0 && num_extra >= 0 && M % P == 0 && ldc == N && c_dtype == TPAT_F32, "tpat_gemm: pos epilogue needs pos, P | M, ldc == N and fp32 C");
  if (M == 0) return 0;
  EpiParams ep{bias, residual, ldr, pos, P, num_extra, epilogue};
  if (extra != nullptr) {
    if (extra->dact_out != nullptr) {
      TPAT_CHECK(epilogue == TPAT_EPI_BIAS_GELU && extra->ld_dact >= N && aligned16(extra->dact_out) && (extra->ld_dact * dtype_size(c_dtype)) % 16 == 0,
                 "tpat_gemm_train: dact_out needs the bias+GELU epilogue and 16-byte aligned rows");
      ep.dact_out = extra->dact_out; ep.ld_dact = extra->ld_dact;
    }
    if (epilogue == TPAT_EPI_DGELU) { ep.aux = extra->aux; ep.ld_aux = extra->ld_aux; }
    if (extra->row_scale != nullptr) {
      TPAT_CHECK(epilogue == TPAT_EPI_BIAS_RESIDUAL && extra->rows_per_clip > 0, "tpat_gemm_train: row_scale needs the residual epilogue and rows_per_clip > 0");
      ep.row_scale = extra->row_scale; ep.rows_per_clip = extra->rows_per_clip;
    }
    if (extra->colsum_out != nullptr)
      TPAT_CHECK(epilogue == TPAT_EPI_DGELU && N % 32 == 0, "tpat_gemm_train: colsum_out needs the DGELU epilogue and N %% 32 == 0");
    if (extra->w_kn) {
      TPAT_CHECK(impl == TPAT_IMPL_TC && N % 8 == 0 && aligned16(W) && fold == nullptr,
                 "tpat_gemm_train: w_kn (W stored [K, N]) exists on the tcgen05 path only and needs N %% 8 == 0");
      ep.w_kn = 1;
    }
  }
  if (fold != nullptr && (fold->xb != nullptr || fold->ln_part != nullptr)) {
    TPAT_CHECK(impl == TPAT_IMPL_TC, "tpat_gemm_ln: the LayerNorm fold exists on the tcgen05 path only");
    if (fold->xb != nullptr) {
      TPAT_CHECK(epilogue == TPAT_EPI_BIAS_RESIDUAL && fold->part_out != nullptr && fold->ldxb >= N && fold->ldxb % 8 == 0 && aligned16(fold->xb),
                 "tpat_gemm_ln: xb needs the residual epilogue, part_out and a 16-byte aligned bf16 buffer with ldxb %% 8 == 0");
      ep.xb = fold->xb; ep.ldxb = fold->ldxb; ep.part_out = fold->part_out; ep.part_ld = N / 32;
    }
    if (fold->ln_part != nullptr) {
      TPAT_CHECK((epilogue == TPAT_EPI_BIAS || epilogue == TPAT_EPI_BIAS_GELU) && fold->ln_colsum != nullptr && aligned16(fold->ln_colsum) && K % 32 == 0,
                 "tpat_gemm_ln: ln_part needs the bias / bias+GELU epilogue, ln_colsum and K %% 32 == 0");
      ep.ln_part = fold->ln_part; ep.ln_chunks = K / 32; ep.ln_colsum = fold->ln_colsum; ep.ln_eps = fold->ln_eps;
    }
  }
  float* colsum_out = extra != nullptr ? extra->colsum_out : nullptr;
  if (impl == TPAT_IMPL_SIMT) {
    if (int rc = gemm_simt(A, a_dtype, lda, W, C, c_dtype, ldc, M, N, K, ep, as_stream(stream))) return rc;
    if (colsum_out != nullptr) {
      TPAT_CHECK(extra->colsum_ws != nullptr, "tpat_gemm_train: colsum_out needs colsum_ws");
      return tpat_colsum(C, c_dtype, ldc, M, N, extra->colsum_ws, colsum_out, stream);
    }
    return 0;
  }
  if (impl == TPAT_IMPL_TC) {
    TPAT_CHECK(a_dtype == TPAT_BF16, "tpat_gemm: the tcgen05 path takes bf16 operands");
    static const bool no_fuse = getenv("TPAT_NO_FUSED_COLSUM") != nullptr;
    const bool fuse = colsum_out != nullptr && !no_fuse && extra->colsum_ws != nullptr &&
                      extra->colsum_ws_floats >= tpat_gemm_colsum_ws_floats(M, N);
    if (fuse) ep.cs_part = extra->colsum_ws;
    if (int rc = gemm_tc(A, lda, W, C, c_dtype, ldc, M, N, K, ep, as_stream(stream))) return rc;
    if (fuse) return finish_colsum_partials(extra->colsum_ws, (M + 31) / 32, N, colsum_out, as_stream(stream));   // (rows past M are zeros)
    if (colsum_out != nullptr) {
      TPAT_CHECK(extra->colsum_ws != nullptr, "tpat_gemm_train: colsum_out needs colsum_ws");
      return tpat_colsum(C, c_dtype, ldc, M, N, extra->colsum_ws, colsum_out, stream);
    }
    return 0;
  }
  set_error("tpat_gemm: bad impl %d", impl);
  return 1;
}
