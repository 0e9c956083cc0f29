// tpat_train_forward / tpat_train_backward: the fine-tune step of the token-pruned ViT in native code.
//
// Forward = the launch sequence of tpat_forward (forward.cu) with every tensor the backward needs written to its own
// slot of a caller-owned arena instead of a reused workspace buffer, the GELU pre-activation and the attention
// log-sum-exp kept, DropPath as a per-clip scale inside the residual GEMM epilogues, optional 2-D token masking.
// Backward = reverse walk over the blocks (reference: autograd over audiomae/models_vit.py:191-207):
//   g (fp32 gradient stream), gb (its DropPath-scaled copy in the GEMM operand dtype)
//   fc2:  dW2 += gb^T a;  dh = (gb W2) * gelu'(h)            fc1:  db1 = colsum(dh); dW1 += dh^T y2;  dy2 = dh W1
//   norm2 backward + residual + scatter of the gathered rows (rows pruned in this block get zero) -> g, gb, dgamma2,
//         dbeta2, db_proj                                       [one pass: tpat_row_bwd]
//   proj: dWp += gb^T ao;  d_ao = gb Wp                        attention backward -> dqkv
//   qkv:  db = colsum(dqkv); dWqkv += dqkv^T y1;  dy1 = dqkv Wqkv
//   norm1 backward + residual -> g, gb (scaled for the previous block's fc2), dgamma1, dbeta1, db_fc2(previous block)
// Data gradients dX = dY W run through the forward's GEMM kernels on [in, out] weight copies; weight gradients
// dW = dY^T X through the same kernels on transposed operand copies (tcgen05 path) or tpat_gemm_f32 (fp32 path).
#include "gemm.cuh"
#include "attention.cuh"

#include <cstdlib>

namespace tpat {

int gemm_wgrad_tc(const void* dY, int ld_dy, const void* X, int ldx, float* dW, int ldw, int K, int Mo, int No, cudaStream_t st);
int validate_forward_args(const tpat_forward_args* a, int n0);

static size_t al(size_t v) { return (v + 255) & ~size_t(255); }

struct BlockSaved {
  int N, N2;            // tokens per clip entering the block / after its gather
  float* x_in; void* y1; void* qkv; void* ao; float* lse; float* x_mid; float* x_g; void* y2; void* h; void* a;
};
struct Saved {
  void* patches; float* x0;   // x0: patch-embed output before masking (only with mask_keep_idx)
  BlockSaved blk[TPAT_MAX_DEPTH];
  float* x_final; float* pooled; float* partial; int n_final;
  size_t bytes;
};
struct BwdWs {
  float* g[2]; void* gb; void* dh; float* dy; void* d_ao; void* dqkv; void* t1; void* t2; int32_t* inv; float* parts; float* delta;
  float* dpooled; float* x0g;
  size_t parts_floats;
  size_t bytes;
};

static int num_extra_of(const tpat_forward_args* a) { return a->variant == TPAT_VARIANT_AST ? 2 : 1; }

static Saved carve_saved(const tpat_train_args* t, uint8_t* base) {
  const tpat_forward_args* a = &t->fwd;
  const int extra = num_extra_of(a);
  const size_t B = a->B, D = a->D, Dh = a->Dh, H = a->H, P = (size_t)(a->T / 16) * (a->F / 16);
  const size_t act = a->impl == TPAT_IMPL_TC ? 2 : 4;
  Saved s;
  size_t off = 0;
  auto take = [&](size_t bytes) { uint8_t* p = base ? base + off : nullptr; off += al(bytes); return p; };
  s.patches = take(B * P * 256 * act);
  s.x0 = t->mask_keep_idx ? (float*)take(B * (P + extra) * D * 4) : nullptr;
  int cur = t->mask_keep_idx ? t->n_keep : (int)P;
  for (int i = 0; i < a->depth; ++i) {
    BlockSaved& k = s.blk[i];
    k.N = extra + cur;
    const int nxt = a->prune[i] ? a->keep[i] : cur;
    k.N2 = extra + nxt;
    const size_t M = B * k.N, M2 = B * k.N2;
    k.x_in = (float*)take(M * D * 4);
    k.y1 = take(M * D * act);
    k.qkv = take(M * 3 * D * act);
    k.ao = take(M * D * act);
    k.lse = (float*)take(B * H * k.N * 4);
    k.x_mid = (float*)take(M * D * 4);
    k.x_g = a->prune[i] ? (float*)take(M2 * D * 4) : k.x_mid;
    k.y2 = take(M2 * D * act);
    k.h = take(M2 * Dh * act);
    k.a = take(M2 * Dh * act);
    cur = nxt;
  }
  s.n_final = extra + cur;
  s.x_final = (float*)take(B * s.n_final * D * 4);
  s.pooled = (float*)take(B * D * 4);
  const size_t Nmax = extra + (t->mask_keep_idx ? (size_t)t->n_keep : P);
  const size_t nqt = (size_t)tpat_attention_qtiles((int)Nmax, a->impl);
  const size_t R = a->variant == TPAT_VARIANT_AST ? H : H * nqt;
  s.partial = (float*)take(B * R * Nmax * 4);
  s.bytes = off;
  return s;
}

static BwdWs carve_bwd(const tpat_train_args* t, uint8_t* base) {
  const tpat_forward_args* a = &t->fwd;
  const int extra = num_extra_of(a);
  const size_t B = a->B, D = a->D, Dh = a->Dh, H = a->H, P = (size_t)(a->T / 16) * (a->F / 16);
  const size_t act = a->impl == TPAT_IMPL_TC ? 2 : 4;
  const size_t Nfull = extra + P;                       // the patch stage works on the unmasked grid
  const size_t Mmax = B * Nfull;
  const size_t Mpad = (Mmax + 63) / 64 * 64;
  const size_t Cmax = Dh > 3 * D ? Dh : 3 * D;
  BwdWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) { uint8_t* p = base ? base + off : nullptr; off += al(bytes); return p; };
  w.g[0] = (float*)take(Mmax * D * 4);
  w.g[1] = (float*)take(Mmax * D * 4);
  w.gb = take(Mmax * D * act);
  w.dh = take(Mmax * Dh * act);
  w.dy = (float*)take(Mmax * D * 4);
  w.d_ao = take(Mmax * D * act);
  w.dqkv = take(Mmax * 3 * D * act);
  w.t1 = a->impl == TPAT_IMPL_TC ? take(Cmax * Mpad * act) : nullptr;
  w.t2 = a->impl == TPAT_IMPL_TC ? take(Cmax * Mpad * act) : nullptr;
  w.inv = (int32_t*)take(B * P * 4);
  // per-CTA partial sums of the row kernels, or the per-32-row column sums the GELU-backward GEMM leaves for fc1's bias
  w.parts_floats = tpat_bwd_partials_floats((int)(Dh > D ? Dh : D));
  if (const size_t need = tpat_gemm_colsum_ws_floats((int)Mmax, (int)Dh); need > w.parts_floats) w.parts_floats = need;
  w.parts = (float*)take(w.parts_floats * 4);
  w.delta = (float*)take(tpat_attention_bwd_ws_floats((int)B, (int)Nfull, (int)H, 64) * 4);
  w.dpooled = (float*)take(B * D * 4);
  w.x0g = t->mask_keep_idx ? (float*)take(Mmax * D * 4) : nullptr;
  w.bytes = off;
  return w;
}

static int validate_train(const tpat_train_args* t) {
  TPAT_CHECK(t != nullptr, "tpat_train: null args");
  const tpat_forward_args* a = &t->fwd;
  if (int rc = validate_forward_args(a, t->mask_keep_idx ? t->n_keep : 0)) return rc;
  TPAT_CHECK(!a->fuse_token && !a->score32, "tpat_train: fuse_token / score32 are inference-only");
  const int P = (a->T / 16) * (a->F / 16);
  if (t->mask_keep_idx) TPAT_CHECK(t->n_keep > 0 && t->n_keep <= P, "tpat_train: n_keep=%d out of range (0, %d]", t->n_keep, P);
  return 0;
}

// dX[M, Nin] = epilogue(dY[M, Nout] . W[Nout, Nin]) through the forward GEMM kernels.  tcgen05 path: W is read as it is
// (the forward weight as an MN-major B operand, tpat_gemm_extra.w_kn); fp32 path: the [Nin, Nout] copy Wt.
static int dgrad(const void* dY, int act, int M, int Nout, int Nin, const void* W, const void* Wt, void* dX, int dx_dtype,
                 int epilogue, const void* aux, int impl, tpat_stream_t st, float* colsum_out = nullptr, float* colsum_ws = nullptr,
                 size_t colsum_ws_floats = 0) {
  tpat_gemm_extra ex{};
  ex.aux = aux; ex.ld_aux = Nin;
  ex.colsum_out = colsum_out; ex.colsum_ws = colsum_ws; ex.colsum_ws_floats = colsum_ws_floats;
  static const bool force_copy = getenv("TPAT_DGRAD_TRANSPOSE") != nullptr;
  if (impl == TPAT_IMPL_TC && !(force_copy && Wt != nullptr)) {
    ex.w_kn = 1;
    return tpat_gemm_train(dY, act, Nout, W, act, nullptr, dX, dx_dtype, Nin, nullptr, 0, M, Nin, Nout, epilogue, impl, &ex, st);
  }
  TPAT_CHECK(Wt != nullptr, "tpat_train_backward: the fp32 path needs the [in, out] weight copies (tpat_block_wt)");
  return tpat_gemm_train(dY, act, Nout, Wt, act, nullptr, dX, dx_dtype, Nin, nullptr, 0, M, Nin, Nout, epilogue, impl, &ex, st);
}

// dW[Nout, Nin] += dY[M, Nout]^T . X[M, Nin]
static int wgrad(const void* dY, const void* X, float* dW, int M, int Nout, int Nin, int act, int impl, const BwdWs& w,
                 tpat_stream_t st) {
  if (dW == nullptr) return 0;
  if (impl == TPAT_IMPL_SIMT)
    return tpat_gemm_f32((const float*)dY, Nout, 1, (const float*)X, Nin, 0, dW, Nin, Nout, Nin, M, 1, st);
  {
    // tcgen05 kernel that reads dY and X as they are (MN-major operands, split-K, TMA reduce-add); shapes it does not
    // cover (outputs that are not multiples of 256 x 256, e.g. ViT-S) take the transpose route below
    static const bool force_transpose = getenv("TPAT_WGRAD_TRANSPOSE") != nullptr;
    if (!force_transpose) {
      const int rc = gemm_wgrad_tc(dY, Nout, X, Nin, dW, Nin, M, Nout, Nin, as_stream(st));
      if (rc >= 0) return rc;
    }
  }
  const int Mpad = (M + 63) / 64 * 64;
  if (int rc = tpat_transpose(dY, act, Nout, w.t1, act, Mpad, M, Nout, st)) return rc;
  if (int rc = tpat_transpose(X, act, Nin, w.t2, act, Mpad, M, Nin, st)) return rc;
  return tpat_gemm(w.t1, act, Mpad, w.t2, act, nullptr, dW, TPAT_F32, Nin, dW, Nin, nullptr, 0, 0, Nout, Nin, Mpad,
                   TPAT_EPI_BIAS_RESIDUAL, impl, st);
}

}  // namespace tpat

extern "C" size_t tpat_sizeof_train_args(void) { return sizeof(tpat_train_args); }

extern "C" size_t tpat_train_saved_bytes(const tpat_train_args* t) {
  if (tpat::validate_train(t) != 0) return 0;
  return tpat::carve_saved(t, nullptr).bytes;
}

extern "C" size_t tpat_train_bwd_workspace_bytes(const tpat_train_args* t) {
  if (tpat::validate_train(t) != 0) return 0;
  return tpat::carve_bwd(t, nullptr).bytes;
}

extern "C" int tpat_train_forward(const tpat_train_args* t, tpat_stream_t stream) {
  using namespace tpat;
  if (int rc = validate_train(t)) return rc;
  const tpat_forward_args* a = &t->fwd;
  TPAT_CHECK(a->spec && a->logits && t->saved, "tpat_train_forward: null spec / logits / saved");
  TPAT_CHECK(aligned16(t->saved), "tpat_train_forward: saved must be 16-byte aligned");
  Saved s = carve_saved(t, reinterpret_cast<uint8_t*>(t->saved));
  TPAT_CHECK(t->saved_bytes >= s.bytes, "tpat_train_forward: saved arena too small (%zu < %zu bytes)", t->saved_bytes, s.bytes);
  const bool ast = a->variant == TPAT_VARIANT_AST;
  const int extra = ast ? 2 : 1, impl = a->impl, act = impl == TPAT_IMPL_TC ? TPAT_BF16 : TPAT_F32;
  const int B = a->B, D = a->D, H = a->H, Dh = a->Dh, P = (a->T / 16) * (a->F / 16);
  const float scale = 0.125f;

  float* x0 = t->mask_keep_idx ? s.x0 : s.blk[0].x_in;
  if (int rc = tpat_patchify(a->spec, s.patches, act, x0, a->extra_tok, a->pos, B, a->T, a->F, D, extra,
                             ast ? TPAT_TOKENS_FREQ_MAJOR : TPAT_TOKENS_TIME_MAJOR, stream)) return rc;
  if (int rc = tpat_gemm(s.patches, act, 256, a->patch_w, act, a->patch_b, x0, TPAT_F32, D, nullptr, 0, a->pos, P, extra,
                         B * P, D, 256, TPAT_EPI_BIAS_POS, impl, stream)) return rc;
  if (t->mask_keep_idx)   // forward_features_mask: keep the surviving patch tokens (models_vit.py:468-476)
    if (int rc = tpat_gather_layernorm(x0, t->mask_keep_idx, s.blk[0].x_in, nullptr, nullptr, nullptr, act, B, extra + P, t->n_keep,
                                       extra, 0, D, a->ln_eps, stream)) return rc;

  for (int i = 0; i < a->depth; ++i) {
    const tpat_block_weights& bw = a->blocks[i];
    const BlockSaved& k = s.blk[i];
    const int N = k.N, M = B * N, N2 = k.N2, M2 = B * N2;
    const bool prune = a->prune[i] != 0;
    const bool want_score = prune || a->want_all_scores;
    if (int rc = tpat_layernorm(k.x_in, bw.ln1_g, bw.ln1_b, k.y1, act, M, D, a->ln_eps, stream)) return rc;
    if (int rc = tpat_gemm(k.y1, act, D, bw.qkv_w, act, bw.qkv_b, k.qkv, act, 3 * D, nullptr, 0, nullptr, 0, 0, M, 3 * D, D,
                           TPAT_EPI_BIAS, impl, stream)) return rc;
    const int smode = !want_score ? TPAT_SCORE_NONE : (ast ? TPAT_SCORE_CLS_ROW : TPAT_SCORE_COLMEAN);
    if (int rc = tpat_attention_train(k.qkv, k.ao, act, s.partial, smode, k.lse, B, N, H, 64, extra, scale, impl, stream)) return rc;
    {
      tpat_gemm_extra ex{};
      ex.row_scale = t->drop_scale[i][0]; ex.rows_per_clip = N;
      if (int rc = tpat_gemm_train(k.ao, act, D, bw.proj_w, act, bw.proj_b, k.x_mid, TPAT_F32, D, k.x_in, D, M, D, D,
                                   TPAT_EPI_BIAS_RESIDUAL, impl, &ex, stream)) return rc;
    }
    if (want_score) {
      const int R = ast ? H : H * tpat_attention_qtiles(N, impl);
      const float divisor = ast ? (float)H : (float)H * (float)(N - extra);
      if (int rc = tpat_score_topk(s.partial, R, divisor, a->scores[i], prune ? a->topk_idx[i] : nullptr, nullptr, B, N, extra,
                                   prune ? a->keep[i] : 0, stream)) return rc;
    }
    if (prune) {
      if (int rc = tpat_gather_layernorm(k.x_mid, a->topk_idx[i], k.x_g, bw.ln2_g, bw.ln2_b, k.y2, act, B, N, a->keep[i], extra, 0,
                                         D, a->ln_eps, stream)) return rc;
    } else {
      if (int rc = tpat_layernorm(k.x_mid, bw.ln2_g, bw.ln2_b, k.y2, act, M, D, a->ln_eps, stream)) return rc;
    }
    {
      tpat_gemm_extra ex{};
      ex.dact_out = k.h; ex.ld_dact = Dh;
      if (int rc = tpat_gemm_train(k.y2, act, D, bw.fc1_w, act, bw.fc1_b, k.a, act, Dh, nullptr, 0, M2, Dh, D, TPAT_EPI_BIAS_GELU,
                                   impl, &ex, stream)) return rc;
    }
    {
      tpat_gemm_extra ex{};
      ex.row_scale = t->drop_scale[i][1]; ex.rows_per_clip = N2;
      float* x_next = i + 1 < a->depth ? s.blk[i + 1].x_in : s.x_final;
      if (int rc = tpat_gemm_train(k.a, act, Dh, bw.fc2_w, act, bw.fc2_b, x_next, TPAT_F32, D, k.x_g, D, M2, D, Dh,
                                   TPAT_EPI_BIAS_RESIDUAL, impl, &ex, stream)) return rc;
    }
  }
  float* pooled = a->pooled ? a->pooled : s.pooled;
  if (int rc = tpat_pool_norm(s.x_final, pooled, a->norm_g, a->norm_b, a->norm_eps, a->head_ln_g, a->head_ln_b, a->head_ln_eps, B,
                              s.n_final, D, a->variant, stream)) return rc;
  if (a->pooled)   // keep a private copy: the head backward reads it
    TPAT_CUDA(cudaMemcpyAsync(s.pooled, pooled, (size_t)B * D * 4, cudaMemcpyDeviceToDevice, as_stream(stream)));
  return tpat_head(pooled, a->head_w, a->head_b, a->logits, B, D, a->num_classes, stream);
}

extern "C" int tpat_train_backward(const tpat_train_args* t, int stage_hi, int stage_lo, tpat_stream_t stream) {
  using namespace tpat;
  if (int rc = validate_train(t)) return rc;
  const tpat_forward_args* a = &t->fwd;
  TPAT_CHECK(t->saved && t->bwd_workspace && t->dlogits, "tpat_train_backward: null saved / bwd_workspace / dlogits");
  TPAT_CHECK(stage_hi <= a->depth + 1 && stage_lo >= 0 && stage_lo <= stage_hi, "tpat_train_backward: bad stage range [%d, %d]", stage_lo, stage_hi);
  Saved s = carve_saved(t, reinterpret_cast<uint8_t*>(t->saved));
  BwdWs w = carve_bwd(t, reinterpret_cast<uint8_t*>(t->bwd_workspace));
  TPAT_CHECK(t->saved_bytes >= s.bytes && t->bwd_workspace_bytes >= w.bytes, "tpat_train_backward: arena too small");
  const bool ast = a->variant == TPAT_VARIANT_AST;
  const int extra = ast ? 2 : 1, impl = a->impl, act = impl == TPAT_IMPL_TC ? TPAT_BF16 : TPAT_F32;
  const int B = a->B, D = a->D, H = a->H, Dh = a->Dh, C = a->num_classes, P = (a->T / 16) * (a->F / 16);
  const float scale = 0.125f;
  // the gradient stream alternates between w.g[0] and w.g[1]; which one is live on entry to a stage is a pure function
  // of the stage number: the head stage writes g[0], every block flips twice, so blocks always start and end on g[0].
  float* g = w.g[0];
  float* g_alt = w.g[1];
  // gradients w.r.t. the LayerNorm outputs (dy) travel in the operand dtype on the tcgen05 path: half the bytes written by
  // the two data-gradient GEMMs and read by the LayerNorm backward; TPAT_TRAIN_DY_FP32=1 keeps them fp32
  static const bool dy_fp32 = getenv("TPAT_TRAIN_DY_FP32") != nullptr;
  const int dyt = (impl == TPAT_IMPL_TC && !dy_fp32) ? TPAT_BF16 : TPAT_F32;

  for (int stage = stage_hi; stage >= stage_lo; --stage) {
    if (stage == a->depth + 1) {
      // ---- head + pooling (models_vit.py:388-389,522 / ast_models.py:500-503) ----
      if (t->d_head_w) if (int rc = tpat_gemm_f32(t->dlogits, C, 1, s.pooled, D, 0, t->d_head_w, D, C, D, B, 1, stream)) return rc;
      if (t->d_head_b) if (int rc = tpat_batch_sum(t->dlogits, t->d_head_b, B, (size_t)C, C, 1, stream)) return rc;
      if (int rc = tpat_gemm_f32(t->dlogits, C, 0, a->head_w, D, 0, w.dpooled, D, B, D, C, 0, stream)) return rc;
      if (int rc = tpat_pool_norm_bwd(s.x_final, w.dpooled, g, a->norm_g, a->norm_b, a->norm_eps, a->head_ln_g, a->head_ln_eps, w.parts,
                                      t->d_norm_g, t->d_norm_b, t->d_head_ln_g, t->d_head_ln_b, B, s.n_final, D, a->variant, stream)) return rc;
      // operand copy of g for the last block's fc2 (+ its bias gradient)
      const int last = a->depth - 1;
      if (int rc = tpat_row_bwd(nullptr, TPAT_F32, nullptr, nullptr, g, nullptr, w.gb, act, t->drop_scale[last][1], nullptr, w.parts,
                                nullptr, nullptr, t->grads[last].fc2_b, B, s.n_final, s.n_final, extra, 0, D, a->ln_eps, stream)) return rc;
    } else if (stage >= 1) {
      const int i = stage - 1;
      const tpat_block_weights& bw = a->blocks[i];
      const tpat_block_grads& gr = t->grads[i];
      const tpat_block_wt& wt = t->wt[i];
      const BlockSaved& k = s.blk[i];
      const int N = k.N, M = B * N, N2 = k.N2, M2 = B * N2;
      const bool prune = a->prune[i] != 0;
      // fc2: gb = scale * dL/dx_out  [M2, D]
      if (int rc = wgrad(w.gb, k.a, gr.fc2_w, M2, D, Dh, act, impl, w, stream)) return rc;
      // (+ fc1's bias gradient = column sums of dh, taken in the same epilogue)
      if (int rc = dgrad(w.gb, act, M2, D, Dh, bw.fc2_w, wt.fc2_wt, w.dh, act, TPAT_EPI_DGELU, k.h, impl, stream, gr.fc1_b, w.parts, w.parts_floats)) return rc;
      // fc1
      if (int rc = wgrad(w.dh, k.y2, gr.fc1_w, M2, Dh, D, act, impl, w, stream)) return rc;
      if (int rc = dgrad(w.dh, act, M2, Dh, D, bw.fc1_w, wt.fc1_wt, w.dy, dyt, TPAT_EPI_BIAS, nullptr, impl, stream)) return rc;
      // norm2 backward + residual (+ scatter): g [M2] -> g_alt [M]
      if (prune) if (int rc = tpat_inverse_index(a->topk_idx[i], w.inv, B, N - extra, a->keep[i], stream)) return rc;
      if (int rc = tpat_row_bwd(w.dy, dyt, k.x_g, bw.ln2_g, g, g_alt, w.gb, act, t->drop_scale[i][0], prune ? w.inv : nullptr, w.parts,
                                gr.ln2_g, gr.ln2_b, gr.proj_b, B, N2, N, extra, 0, D, a->ln_eps, stream)) return rc;
      { float* tmp = g; g = g_alt; g_alt = tmp; }
      // proj
      if (int rc = wgrad(w.gb, k.ao, gr.proj_w, M, D, D, act, impl, w, stream)) return rc;
      if (int rc = dgrad(w.gb, act, M, D, D, bw.proj_w, wt.proj_wt, w.d_ao, act, TPAT_EPI_BIAS, nullptr, impl, stream)) return rc;
      // attention
      // (tcgen05 path: + the qkv bias gradient = column sums of dqkv, taken where dqkv is produced)
      static const bool no_fuse = getenv("TPAT_NO_FUSED_COLSUM") != nullptr;
      const bool fuse_qkv_b = impl == TPAT_IMPL_TC && act == TPAT_BF16 && !no_fuse && getenv("TPAT_ATTN_BWD_SIMT") == nullptr;
      if (int rc = tpat_attention_bwd(k.qkv, k.ao, w.d_ao, k.lse, w.dqkv, act, B, N, H, 64, scale, impl, w.delta,
                                      fuse_qkv_b ? gr.qkv_b : nullptr, stream)) return rc;
      // qkv
      if (gr.qkv_b && !fuse_qkv_b) if (int rc = tpat_colsum(w.dqkv, act, 3 * D, M, 3 * D, w.parts, gr.qkv_b, stream)) return rc;
      if (int rc = wgrad(w.dqkv, k.y1, gr.qkv_w, M, 3 * D, D, act, impl, w, stream)) return rc;
      if (int rc = dgrad(w.dqkv, act, M, 3 * D, D, bw.qkv_w, wt.qkv_wt, w.dy, dyt, TPAT_EPI_BIAS, nullptr, impl, stream)) return rc;
      // norm1 backward + residual: g [M] -> g_alt [M]; operand copy scaled for the previous block's fc2 (+ its bias gradient)
      if (int rc = tpat_row_bwd(w.dy, dyt, k.x_in, bw.ln1_g, g, g_alt, i > 0 ? w.gb : nullptr, act, i > 0 ? t->drop_scale[i - 1][1] : nullptr,
                                nullptr, w.parts, gr.ln1_g, gr.ln1_b, i > 0 ? t->grads[i - 1].fc2_b : nullptr, B, N, N, extra, 0, D, a->ln_eps, stream)) return rc;
      { float* tmp = g; g = g_alt; g_alt = tmp; }
    } else {
      // ---- patch embedding, extra tokens, position table (models_vit.py:357-362 / ast_models.py:460-466) ----
      const int N0 = s.blk[0].N;      // tokens per clip entering block 0 (after masking)
      const float* gfull = g;         // [B, extra + P, D] gradient w.r.t. the embedded tokens
      if (t->mask_keep_idx) {         // backward of the masking gather: scatter to the full patch grid
        if (int rc = tpat_inverse_index(t->mask_keep_idx, w.inv, B, P, t->n_keep, stream)) return rc;
        if (int rc = tpat_row_bwd(nullptr, TPAT_F32, nullptr, nullptr, g, w.x0g, nullptr, act, nullptr, w.inv, w.parts, nullptr, nullptr, nullptr,
                                  B, N0, extra + P, extra, 0, D, a->ln_eps, stream)) return rc;
        gfull = w.x0g;
      }
      const size_t clip = (size_t)(extra + P) * D;
      if (t->d_extra_tok) if (int rc = tpat_batch_sum(gfull, t->d_extra_tok, B, clip, extra * D, 1, stream)) return rc;
      if (t->d_pos) if (int rc = tpat_batch_sum(gfull, t->d_pos, B, clip, (int)clip, 1, stream)) return rc;
      // compact the patch rows into the operand copy [B * P, D] (+ the conv bias gradient), then the weight gradient
      if (int rc = tpat_row_bwd(nullptr, TPAT_F32, nullptr, nullptr, gfull, nullptr, w.gb, act, nullptr, nullptr, w.parts, nullptr, nullptr,
                                t->d_patch_b, B, extra + P, P, 0, extra, D, a->ln_eps, stream)) return rc;
      if (int rc = wgrad(w.gb, s.patches, t->d_patch_w, B * P, D, 256, act, impl, w, stream)) return rc;
    }
  }
  return 0;
}
