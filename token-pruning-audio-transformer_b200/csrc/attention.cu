// tpat_attention: argument validation and dispatch (CUDA-core parity path / tcgen05 path).
#include "attention.cuh"

extern "C" int tpat_attention_qtiles(int N, int impl) {
  if (N <= 0) return 0;
  return impl == TPAT_IMPL_TC ? tpat::attention_tc_qtiles(N) : tpat::attention_simt_qtiles(N);
}

extern "C" int tpat_attention(const void* qkv, void* out, int dtype, float* score_partial, int score_mode,
                              int B, int N, int H, int hd, int num_extra, float scale, int impl,
                              tpat_stream_t stream) {
  return tpat_attention_split(qkv, nullptr, out, dtype, score_partial, score_mode, B, N, H, hd, num_extra, scale, impl, stream);
}

static int attention_entry(const void* qkv, const void* qk_planes, void* out, int dtype, float* score_partial,
                           int score_mode, int B, int N, int H, int hd, int num_extra, float scale, int impl, float* lse,
                           tpat_stream_t stream) {
  using namespace tpat;
  TPAT_CHECK(qk_planes == nullptr || (impl == TPAT_IMPL_TC && aligned16(qk_planes)), "tpat_attention_split: planes need the tcgen05 path and 16-byte alignment");
  TPAT_CHECK(qkv && out, "tpat_attention: null pointer");
  TPAT_CHECK(B >= 0 && N > 0 && H > 0, "tpat_attention: bad sizes B=%d N=%d H=%d", B, N, H);
  TPAT_CHECK(hd == 64, "tpat_attention: head dim must be 64 (got %d)", hd);
  TPAT_CHECK(dtype == TPAT_F32 || dtype == TPAT_BF16, "tpat_attention: bad dtype %d", dtype);
  TPAT_CHECK(score_mode >= TPAT_SCORE_NONE && score_mode <= TPAT_SCORE_COLMEAN, "tpat_attention: bad score mode %d", score_mode);
  TPAT_CHECK(score_mode == TPAT_SCORE_NONE || score_partial != nullptr, "tpat_attention: score mode %d needs score_partial", score_mode);
  TPAT_CHECK(num_extra >= 0 && num_extra < N, "tpat_attention: bad num_extra %d", num_extra);
  TPAT_CHECK(aligned16(qkv) && aligned16(out), "tpat_attention: pointers must be 16-byte aligned");
  if (B == 0) return 0;
  if (impl == TPAT_IMPL_SIMT) return attention_simt(qkv, out, dtype, score_partial, score_mode, B, N, H, num_extra, scale, as_stream(stream), lse);
  if (impl == TPAT_IMPL_TC) {
    TPAT_CHECK(dtype == TPAT_BF16, "tpat_attention: the tcgen05 path takes bf16 operands");
    return attention_tc(qkv, out, score_partial, score_mode, B, N, H, num_extra, scale, as_stream(stream), qk_planes, lse);
  }
  set_error("tpat_attention: bad impl %d", impl);
  return 1;
}

extern "C" int tpat_attention_split(const void* qkv, const void* qk_planes, void* out, int dtype, float* score_partial,
                                    int score_mode, int B, int N, int H, int hd, int num_extra, float scale, int impl,
                                    tpat_stream_t stream) {
  return attention_entry(qkv, qk_planes, out, dtype, score_partial, score_mode, B, N, H, hd, num_extra, scale, impl, nullptr, stream);
}

extern "C" int tpat_attention_train(const void* qkv, void* out, int dtype, float* score_partial, int score_mode, float* lse,
                                    int B, int N, int H, int hd, int num_extra, float scale, int impl, tpat_stream_t stream) {
  TPAT_CHECK(lse != nullptr, "tpat_attention_train: lse is required");
  return attention_entry(qkv, nullptr, out, dtype, score_partial, score_mode, B, N, H, hd, num_extra, scale, impl, lse, stream);
}

extern "C" int tpat_attention_bwd(const void* qkv, const void* out, const void* d_out, const float* lse, void* dqkv, int dtype,
                                  int B, int N, int H, int hd, float scale, int impl, float* delta_ws, float* dbias,
                                  tpat_stream_t stream) {
  using namespace tpat;
  TPAT_CHECK(qkv && out && d_out && lse && dqkv && delta_ws, "tpat_attention_bwd: null pointer");
  TPAT_CHECK(B >= 0 && N > 0 && H > 0 && hd == 64, "tpat_attention_bwd: bad sizes B=%d N=%d H=%d hd=%d", B, N, H, hd);
  TPAT_CHECK(dtype == TPAT_F32 || dtype == TPAT_BF16, "tpat_attention_bwd: bad dtype %d", dtype);
  TPAT_CHECK(aligned16(qkv) && aligned16(out) && aligned16(d_out) && aligned16(dqkv), "tpat_attention_bwd: pointers must be 16-byte aligned");
  if (B == 0) return 0;
  if (impl == TPAT_IMPL_TC && dtype == TPAT_BF16) return attention_bwd_tc(qkv, out, d_out, lse, dqkv, B, N, H, scale, delta_ws, dbias, as_stream(stream));
  TPAT_CHECK(dbias == nullptr, "tpat_attention_bwd: the fused bias gradient (dbias) exists on the tcgen05 bf16 path only; use tpat_colsum");
  return attention_bwd_simt(qkv, out, d_out, lse, dqkv, dtype, B, N, H, scale, delta_ws, as_stream(stream));
}
