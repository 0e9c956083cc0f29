// tpat_forward: the whole token-pruned ViT forward as one native call.
//
// Native restatement of the reference's hot loop -- VisionTransformer.forward_features
// (audiomae/models_vit.py:334-396) / ASTModel.forward (ast/src/models/ast_models.py:424-508) and
// Block.forward (models_vit.py:191-207) -- as a fixed sequence of kernel launches on one stream:
//   patchify -> patch GEMM(+pos) -> 12 x [LN1 -> QKV GEMM -> attention(+score partials) ->
//   proj GEMM(+residual) -> score/top-k -> gather+LN2 | LN2 -> fc1 GEMM(+GELU) -> fc2 GEMM(+residual)]
//   -> pool/norm -> head GEMM.
// No host synchronisation, no allocation: everything lives in the caller's workspace, so the call
// can be captured in a CUDA graph.  The residual stream stays fp32 in both precisions.
#include "gemm.cuh"
#include "attention.cuh"

namespace tpat {

struct Workspace {
  float* x[2];        // residual stream ping-pong [B*Nmax*D]
  void* y;            // LayerNorm output       [B*Nmax*D]   (act dtype)
  void* qkv;          //                         [B*Nmax*3D]
  void* ao;           // attention output        [B*Nmax*D]
  void* hid;          // MLP hidden / patch matrix [B*Nmax*Dh]
  float* partial;     // score partials          [B*R*Nmax]
  float* score_tmp;   // block score when the caller does not ask for it but the fused token needs it [B*Nmax]
  int32_t* rest;      // not-kept token indices (fused token) [B*Nmax]
  float* pooled;
  float* part;      // [B*D]
  void* y3;           // score32: split-bf16 LayerNorm output [B*Nmax*3D] bf16
  float* qk32;        // score32: fp32 q | k projection       [B*Nmax*2D]
  void* qkp;          // score32: split-bf16 planes of it     [B*Nmax*4D] bf16
  size_t bytes;
};

static size_t align_up(size_t v) { return (v + 255) & ~size_t(255); }

int validate_forward_args(const tpat_forward_args* a, int n0) {
  // (shared with train.cu; n0 > 0: patch tokens entering block 0 when the fine-tune 2-D masking dropped some)
  TPAT_CHECK(a != nullptr, "tpat_forward: null args");
  TPAT_CHECK(a->variant == TPAT_VARIANT_AUDIOMAE || a->variant == TPAT_VARIANT_AST, "tpat_forward: bad variant %d", a->variant);
  TPAT_CHECK(a->impl == TPAT_IMPL_SIMT || a->impl == TPAT_IMPL_TC, "tpat_forward: bad impl %d", a->impl);
  TPAT_CHECK(a->B > 0 && a->T >= 16 && a->F >= 16 && a->T % 16 == 0 && a->F % 16 == 0, "tpat_forward: bad input shape B=%d T=%d F=%d", a->B, a->T, a->F);
  TPAT_CHECK(a->depth > 0 && a->depth <= TPAT_MAX_DEPTH, "tpat_forward: depth %d out of range", a->depth);
  TPAT_CHECK(a->D > 0 && a->D % 128 == 0 && a->H > 0 && a->D == a->H * 64, "tpat_forward: need D == 64*H and D %% 128 == 0 (D=%d H=%d)", a->D, a->H);
  TPAT_CHECK(a->Dh > 0 && a->Dh % 64 == 0 && a->num_classes > 0, "tpat_forward: bad Dh=%d or num_classes=%d", a->Dh, a->num_classes);
  int n = n0 > 0 ? n0 : (a->T / 16) * (a->F / 16);
  for (int i = 0; i < a->depth; ++i) {
    TPAT_CHECK(a->keep[i] > 0 && a->keep[i] <= n, "tpat_forward: keep[%d]=%d must be in (0, %d]", i, a->keep[i], n);
    TPAT_CHECK(a->prune[i] || a->keep[i] == n, "tpat_forward: block %d drops tokens (%d -> %d) but prune[%d] is 0", i, n, a->keep[i], i);
    if (a->prune[i]) TPAT_CHECK(a->topk_idx[i] != nullptr, "tpat_forward: block %d prunes but topk_idx[%d] is NULL", i, i);
    if (a->prune[i] && a->score32 && a->impl == TPAT_IMPL_TC)
      TPAT_CHECK(a->blocks[i].qk_w_split != nullptr, "tpat_forward: score32 needs blocks[%d].qk_w_split", i);
    n = a->keep[i] + ((a->prune[i] && a->fuse_token && a->keep[i] < n) ? 1 : 0);
  }
  return 0;
}

// LayerNorm fold (tpat_gemm_ln): norm1 of block i > 0 is folded when the block carries the gamma-scaled qkv weights (the
// fc2 of block i - 1 then emits bf16(x) + moments); norm2 is folded in blocks that do not prune (the gather comes first).
static bool fold_ln1(const tpat_forward_args* a, int i) {
  if (i < a->depth && a->prune[i] && a->score32) return false;   // the split-bf16 LayerNorm of a score32 pruning block is a real pass
  return a->impl == TPAT_IMPL_TC && i > 0 && i < a->depth && a->blocks[i].qkv_w_ln && a->blocks[i].qkv_colsum && a->blocks[i].qkv_b_ln;
}
static bool fold_ln2(const tpat_forward_args* a, int i) {
  return a->impl == TPAT_IMPL_TC && !a->prune[i] && a->blocks[i].fc1_w_ln && a->blocks[i].fc1_colsum && a->blocks[i].fc1_b_ln;
}

static Workspace carve(const tpat_forward_args* a, uint8_t* base) {
  const int extra = a->variant == TPAT_VARIANT_AST ? 2 : 1;
  const size_t P = (size_t)(a->T / 16) * (a->F / 16), Nmax = extra + P, B = a->B, D = a->D;
  const size_t act = a->impl == TPAT_IMPL_TC ? 2 : 4;
  const size_t nqt = (size_t)tpat_attention_qtiles((int)Nmax, a->impl);
  const size_t R = a->variant == TPAT_VARIANT_AST ? (size_t)a->H : (size_t)a->H * nqt;
  Workspace w;
  size_t off = 0;
  auto take = [&](size_t bytes) { uint8_t* p = base ? base + off : nullptr; off += align_up(bytes); return p; };
  w.x[0] = (float*)take(B * Nmax * D * 4);
  w.x[1] = (float*)take(B * Nmax * D * 4);
  w.y = take(B * Nmax * D * act);
  w.qkv = take(B * Nmax * 3 * D * act);
  w.ao = take(B * Nmax * D * act);
  size_t hid = B * Nmax * (size_t)a->Dh * act, pat = B * P * 256 * act;
  w.hid = take(hid > pat ? hid : pat);
  w.partial = (float*)take(B * R * Nmax * 4);
  w.score_tmp = (float*)take(B * Nmax * 4);
  w.rest = (int32_t*)take(B * Nmax * 4);
  w.pooled = (float*)take(B * D * 4);
  w.part = (float*)take(B * Nmax * (D / 32) * 2 * 4);   // LayerNorm fold: partial moments of every row of x
  w.y3 = nullptr; w.qk32 = nullptr; w.qkp = nullptr;
  if (a->score32 && a->impl == TPAT_IMPL_TC) {
    w.y3 = take(B * Nmax * 3 * D * 2);
    w.qk32 = (float*)take(B * Nmax * 2 * D * 4);
    w.qkp = take(B * Nmax * 4 * D * 2);
  }
  w.bytes = off;
  return w;
}

}  // namespace tpat

extern "C" size_t tpat_forward_workspace_bytes(const tpat_forward_args* a) {
  if (tpat::validate_forward_args(a, 0) != 0) return 0;
  return tpat::carve(a, nullptr).bytes;
}

extern "C" int tpat_forward_launch_count(const tpat_forward_args* a) {
  if (tpat::validate_forward_args(a, 0) != 0) return -1;
  using tpat::fold_ln1; using tpat::fold_ln2;
  int n = 2;  // patchify + patch GEMM
  const int extra = a->variant == TPAT_VARIANT_AST ? 2 : 1;
  int cur = (a->T / 16) * (a->F / 16);
  for (int i = 0; i < a->depth; ++i) {
    const bool prune = a->prune[i] != 0;
    const bool score = prune || a->want_all_scores;
    n += fold_ln1(a, i) ? 3 : 4;              // (LN1,) QKV, attention, proj
    if (prune && a->score32 && a->impl == TPAT_IMPL_TC) n += 2;   // split q | k GEMM + plane split
    // AST score blocks on the tensor-core path: the cls tile (two-pass) and the other tiles are separate launches
    if (score && a->variant == TPAT_VARIANT_AST && a->impl == TPAT_IMPL_TC && tpat_attention_qtiles(extra + cur, a->impl) > 1) n += 1;
    if (score) n += 1;                        // score / top-k
    n += fold_ln2(a, i) ? 2 : 3;              // ((gather+)LN2,) fc1, fc2
    if (prune && a->fuse_token && a->keep[i] < cur) { n += 1; cur = a->keep[i] + 1; }   // fused token
    else cur = a->keep[i];
  }
  return n + 2;  // pool/norm + head
}

extern "C" int tpat_forward(const tpat_forward_args* a, tpat_stream_t stream) {
  using namespace tpat;
  if (int rc = validate_forward_args(a, 0)) return rc;
  TPAT_CHECK(a->spec && a->logits && a->workspace, "tpat_forward: null spec / logits / workspace");
  TPAT_CHECK(aligned16(a->workspace), "tpat_forward: workspace must be 16-byte aligned");
  Workspace w = carve(a, reinterpret_cast<uint8_t*>(a->workspace));
  TPAT_CHECK(a->workspace_bytes >= w.bytes, "tpat_forward: workspace too small (%zu < %zu bytes)", a->workspace_bytes, w.bytes);

  const bool ast = a->variant == TPAT_VARIANT_AST;
  const int extra = ast ? 2 : 1;
  const int impl = a->impl;
  const int act = impl == TPAT_IMPL_TC ? TPAT_BF16 : TPAT_F32;
  const int B = a->B, D = a->D, H = a->H, Dh = a->Dh;
  const int P = (a->T / 16) * (a->F / 16);
  const float scale = 0.125f;  // head_dim ** -0.5 with head_dim == 64 (models_vit.py:58)

  // Alternate the walk direction of the big kernels along the producer -> consumer chain (see g_walk_desc):
  // each one starts on the rows the previous one wrote last.  TPAT_WALK=0 keeps every kernel ascending.
  static const bool alternate = [] { const char* e = getenv("TPAT_WALK"); return !(e && e[0] == '0'); }();
  struct WalkGuard { ~WalkGuard() { g_walk_desc = 0; } } walk_guard;
  g_walk_desc = 0;
  auto flip = [&]() { if (alternate) g_walk_desc ^= 1; };

  // patch embed + pos + cls/dist rows (models_vit.py:357-362 / ast_models.py:460-466)
  if (int rc = tpat_patchify(a->spec, w.hid, act, w.x[0], a->extra_tok, a->pos, B, a->T, a->F, D, extra,
                             ast ? TPAT_TOKENS_FREQ_MAJOR : TPAT_TOKENS_TIME_MAJOR, stream)) return rc;
  if (int rc = tpat_gemm(w.hid, act, 256, a->patch_w, act, a->patch_b, w.x[0], TPAT_F32, D, nullptr, 0, a->pos, P, extra,
                         B * P, D, 256, TPAT_EPI_BIAS_POS, impl, stream)) return rc;

  // optional: keep (part of) the live residual rows in the persisting part of L2 (see L2Window in common.cuh)
  const size_t l2_persist = impl == TPAT_IMPL_TC ? l2_persist_bytes() : 0;
  struct L2Guard { ~L2Guard() { g_l2_window = L2Window{}; } } l2_guard;
  auto l2_window = [&](const float* xbuf, size_t rows) {
    if (l2_persist == 0) return;
    const size_t bytes = rows * (size_t)D * sizeof(float);
    g_l2_window.base = xbuf;
    g_l2_window.bytes = bytes;
    g_l2_window.ratio = bytes <= l2_persist ? 1.0f : (float)((double)l2_persist / (double)bytes);
  };

  int cur = P;       // non-extra tokens
  int xi = 0;        // which residual buffer is live
  for (int i = 0; i < a->depth; ++i) {
    const tpat_block_weights& bw = a->blocks[i];
    const int N = extra + cur, M = B * N;
    const bool prune = a->prune[i] != 0;
    const bool want_score = prune || a->want_all_scores;
    float* x = w.x[xi];
    l2_window(x, (size_t)M);
    const bool split = prune && a->score32 && impl == TPAT_IMPL_TC;
    if (split) {
      // "bf16+score32": LayerNorm -> split-bf16 triple; qkv from its hi segment as usual; q | k again as a K = 3D
      // split GEMM in fp32, re-split into hi / lo planes for the score tiles of the attention kernel
      flip();
      if (int rc = tpat_layernorm(x, bw.ln1_g, bw.ln1_b, w.y3, TPAT_BF16_SPLIT3, M, D, a->ln_eps, stream)) return rc;
      flip();
      if (int rc = tpat_gemm(w.y3, act, 3 * D, bw.qkv_w, act, bw.qkv_b, w.qkv, act, 3 * D, nullptr, 0, nullptr, 0, 0,
                             M, 3 * D, D, TPAT_EPI_BIAS, impl, stream)) return rc;
      if (int rc = tpat_gemm(w.y3, act, 3 * D, bw.qk_w_split, act, bw.qkv_b, w.qk32, TPAT_F32, 2 * D, nullptr, 0, nullptr, 0, 0,
                             M, 2 * D, 3 * D, TPAT_EPI_BIAS, impl, stream)) return rc;
      if (int rc = tpat_split_bf16(w.qk32, w.qkp, M, 2 * D, stream)) return rc;
    } else if (fold_ln1(a, i)) {
      // norm1 folded: w.y holds bf16(x) and w.part its row moments (written by the previous block's fc2)
      const tpat_ln_fold f{nullptr, 0, nullptr, w.part, bw.qkv_colsum, a->ln_eps};
      flip();
      if (int rc = tpat_gemm_ln(w.y, act, D, bw.qkv_w_ln, act, bw.qkv_b_ln, w.qkv, act, 3 * D, nullptr, 0, nullptr, 0, 0,
                                M, 3 * D, D, TPAT_EPI_BIAS, impl, &f, stream)) return rc;
    } else {
      flip();
      if (int rc = tpat_layernorm(x, bw.ln1_g, bw.ln1_b, w.y, act, M, D, a->ln_eps, stream)) return rc;
      flip();
      if (int rc = tpat_gemm(w.y, act, D, bw.qkv_w, act, bw.qkv_b, w.qkv, act, 3 * D, nullptr, 0, nullptr, 0, 0,
                             M, 3 * D, D, TPAT_EPI_BIAS, impl, stream)) return rc;
    }
    const int smode = !want_score ? TPAT_SCORE_NONE : (ast ? TPAT_SCORE_CLS_ROW : TPAT_SCORE_COLMEAN);
    flip();
    if (int rc = tpat_attention_split(w.qkv, split ? w.qkp : nullptr, w.ao, act, w.partial, smode, B, N, H, 64, extra, scale, impl, stream)) return rc;
    flip();
    {
      // proj + residual; when norm2 is folded it also emits bf16(x) (into w.y) and the row moments
      const tpat_ln_fold f{fold_ln2(a, i) ? w.y : nullptr, D, w.part, nullptr, nullptr, 0.f};
      if (int rc = tpat_gemm_ln(w.ao, act, D, bw.proj_w, act, bw.proj_b, x, TPAT_F32, D, x, D, nullptr, 0, 0,
                                M, D, D, TPAT_EPI_BIAS_RESIDUAL, impl, &f, stream)) return rc;
    }
    if (want_score) {
      const int R = ast ? H : H * tpat_attention_qtiles(N, impl);
      const float divisor = ast ? (float)H : (float)H * (float)(N - extra);
      const bool fuse_here = prune && a->fuse_token && a->keep[i] < cur;   // something is dropped -> fused token
      float* score_out = a->scores[i] ? a->scores[i] : (fuse_here ? w.score_tmp : nullptr);
      if (int rc = tpat_score_topk(w.partial, R, divisor, score_out, prune ? a->topk_idx[i] : nullptr,
                                   fuse_here ? w.rest : nullptr, B, N, extra, prune ? a->keep[i] : 0, stream)) return rc;
    }
    int M2 = M;
    if (prune || !fold_ln2(a, i)) flip();
    if (prune) {
      float* xn = w.x[xi ^ 1];
      const bool fuse_here = a->fuse_token && a->keep[i] < cur;
      const int out_rows = extra + a->keep[i] + (fuse_here ? 1 : 0);
      if (int rc = tpat_gather_layernorm(x, a->topk_idx[i], xn, bw.ln2_g, bw.ln2_b, w.y, act, B, N, a->keep[i], extra, out_rows,
                                         D, a->ln_eps, stream)) return rc;
      if (fuse_here) {
        const float* sc = a->scores[i] ? a->scores[i] : w.score_tmp;
        if (int rc = tpat_fuse_token(x, sc, w.rest, xn, bw.ln2_g, bw.ln2_b, w.y, act, B, N, cur - a->keep[i], out_rows,
                                     out_rows - 1, extra, D, a->ln_eps, stream)) return rc;
      }
      xi ^= 1;
      x = xn;
      cur = out_rows - extra;
      M2 = B * out_rows;
      l2_window(x, (size_t)M2);
    } else if (!fold_ln2(a, i)) {
      if (int rc = tpat_layernorm(x, bw.ln2_g, bw.ln2_b, w.y, act, M, D, a->ln_eps, stream)) return rc;
    }
    flip();
    if (fold_ln2(a, i)) {
      const tpat_ln_fold f{nullptr, 0, nullptr, w.part, bw.fc1_colsum, a->ln_eps};
      if (int rc = tpat_gemm_ln(w.y, act, D, bw.fc1_w_ln, act, bw.fc1_b_ln, w.hid, act, Dh, nullptr, 0, nullptr, 0, 0,
                                M2, Dh, D, TPAT_EPI_BIAS_GELU, impl, &f, stream)) return rc;
    } else {
      if (int rc = tpat_gemm(w.y, act, D, bw.fc1_w, act, bw.fc1_b, w.hid, act, Dh, nullptr, 0, nullptr, 0, 0,
                             M2, Dh, D, TPAT_EPI_BIAS_GELU, impl, stream)) return rc;
    }
    flip();
    {
      // fc2 + residual; emits bf16(x) + moments when the next block's norm1 is folded
      const tpat_ln_fold f{fold_ln1(a, i + 1) ? w.y : nullptr, D, w.part, nullptr, nullptr, 0.f};
      if (int rc = tpat_gemm_ln(w.hid, act, Dh, bw.fc2_w, act, bw.fc2_b, x, TPAT_F32, D, x, D, nullptr, 0, 0,
                                M2, D, Dh, TPAT_EPI_BIAS_RESIDUAL, impl, &f, stream)) return rc;
    }
  }

  // pooled head (models_vit.py:387-389,522 / ast_models.py:500-503); always fp32 CUDA cores (tiny)
  float* pooled = a->pooled ? a->pooled : w.pooled;
  if (int rc = tpat_pool_norm(w.x[xi], pooled, a->norm_g, a->norm_b, a->norm_eps, a->head_ln_g, a->head_ln_b,
                              a->head_ln_eps, B, extra + cur, D, a->variant, stream)) return rc;
  return tpat_head(pooled, a->head_w, a->head_b, a->logits, B, D, a->num_classes, stream);
}
