/*
 * tpat.h -- C-ABI of libtpat.so: the B200 (sm_100a) kernels behind the token-pruned ViT-B/16
 * forward AND fine-tune step (backward, optimizer) of andylee-24/token-pruning-audio-transformer.
 *
 * The reference has no native layer and no FFI (it is 100 % PyTorch); its "plugin
 * interface" for this path is the Python model API (audiomae/models_vit.py:502-527,
 * ast/src/models/ast_models.py:424-508).  The Python mirror of that API lives in
 * token-pruning-audio-transformer_b200/tpat/ and reaches the GPU only through the entry points
 * below.  Each entry point cites the reference code whose ATen/cuBLAS/cuDNN call sequence it
 * replaces.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; tpat_last_error() then holds a
 *     thread-local message.  Nothing throws or aborts across the boundary.
 *   - all pointers are DEVICE pointers unless named host_*; the library never allocates or frees
 *     memory the caller sees.  `stream` is a cudaStream_t (torch.cuda.current_stream().cuda_stream).
 *   - tokens are rows of a dense row-major [B * N, D] matrix; all clips of a batch share N.
 *   - there is no CPU fallback: on a machine without an sm_100 GPU the compute calls fail.
 */
#ifndef TPAT_H
#define TPAT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TPAT_VERSION 11
#define TPAT_MAX_DEPTH 32

typedef void* tpat_stream_t; /* cudaStream_t */

/* element types */
enum {
  TPAT_F32 = 0,
  TPAT_BF16 = 1,
  TPAT_BF16_SPLIT3 = 2 /* tpat_layernorm output only: [rows, 3 * D] bf16 = [hi | lo | hi], hi = bf16(y), lo = bf16(y - hi) */
};
/* GEMM epilogues */
enum {
  TPAT_EPI_BIAS = 0,          /* C = A W^T + b                               */
  TPAT_EPI_BIAS_GELU = 1,     /* C = gelu_erf(A W^T + b)                     */
  TPAT_EPI_BIAS_RESIDUAL = 2, /* C = R + A W^T + b   (C may alias R)         */
  TPAT_EPI_BIAS_POS = 3,      /* patch-embed: C[b, extra+p] = A W^T + b + pos[extra+p] */
  TPAT_EPI_DGELU = 4          /* training (tpat_gemm_train): C = (A W^T) * aux, aux = saved gelu'(h) */
};
/* GEMM / attention implementation */
enum {
  TPAT_IMPL_SIMT = 0, /* fp32-FMA CUDA-core kernels: the fp32 parity ("fp32-scoring") mode     */
  TPAT_IMPL_TC = 1    /* tcgen05 / TMEM / TMA kernels, bf16 operands, fp32 accumulate          */
};
/* importance-score flavour */
enum {
  TPAT_SCORE_NONE = 0,
  TPAT_SCORE_CLS_ROW = 1, /* AST:      mean_h P[h, 0, j],      j >= extra  (ast_models.py:124) */
  TPAT_SCORE_COLMEAN = 2  /* AudioMAE: mean_{h, i>=extra} P[h, i, j], j >= extra (models_vit.py:113) */
};
/* spectrogram -> token order */
enum {
  TPAT_TOKENS_TIME_MAJOR = 0, /* AudioMAE: token = t_blk * (F/16) + f_blk, kernel index [t_off][f_off] */
  TPAT_TOKENS_FREQ_MAJOR = 1  /* AST:      token = f_blk * (T/16) + t_blk, kernel index [f_off][t_off] */
};
/* model family for tpat_forward */
enum { TPAT_VARIANT_AUDIOMAE = 0, TPAT_VARIANT_AST = 1 };

int tpat_version(void);
const char* tpat_last_error(void);
/* 1 when the current device is compute capability 10.x, 0 otherwise (no error is set) */
int tpat_device_ok(void);

/*
 * Kaldi-compatible log-mel front end: waveforms -> the normalised spectrogram tpat_forward / tpat_patchify consume.
 * Replaces (eval path) `waveform - waveform.mean()`, torchaudio.compliance.kaldi.fbank(htk_compat=True, use_energy=False,
 * window_type='hanning', num_mel_bins=n_mel, dither=0.0, frame_shift=10), pad-with-minimum / crop to T frames and
 * `(fbank - norm_mean) / (norm_std * 2)` (audiomae/dataset.py:175-230,298; ast/src/dataloader.py:98-149,204).
 * torchaudio is the reference's pinned third-party dependency (torchaudio==2.4.1, amae_pruning_miniconda.yml:90).
 *   wave     [B, L] fp32 mono; lengths [B] int32 valid samples per clip or NULL (all L)
 *   subtract_clip_mean  != 0: subtract each clip's mean first;  clip_mean_ws: 32 * B floats of scratch, 8-byte aligned
 *            (always required: partial sums and the clip's minimum log-mel value live there)
 *   window   [win] fp32 (torch.hann_window(win, periodic=False)); frames of `win` samples every `shift`, snip_edges:
 *            n_frames(b) = 1 + (len_b - win) / shift; per frame: DC removal, pre-emphasis `preemph` with replicate
 *            padding, window, zero-pad to nfft (256 | 512 | 1024), power spectrum
 *   mel      [n_mel, nfft/2 + 1] fp32 triangular filters (kaldi get_mel_banks + one zero column), mel_start / mel_len
 *            [n_mel] int32: first non-zero FFT bin and number of bins of each filter
 *   spec     [B, T, n_mel] fp32 out (n_mel % 4 == 0): log(max(energy, FLT_EPSILON)); rows >= n_frames(b) = the clip's minimum; then
 *            (v - norm_mean) / (2 norm_std)
 */
int tpat_fbank(const float* wave, const int* lengths, int B, int L, int subtract_clip_mean, float* clip_mean_ws,
               const float* window, int win, int shift, int nfft, float preemph, const float* mel,
               const int* mel_start, const int* mel_len, int n_mel, float* spec, int T, float norm_mean,
               float norm_std, tpat_stream_t stream);

/*
 * Patch extraction (im2col of the 16x16 / stride-16 conv) + extra-token rows.
 * Replaces: PatchEmbed.forward's unfold (models_vit.py:241-247, ast_models.py:36-42) and the
 * cls / dist row assembly (models_vit.py:359-362, ast_models.py:463-466).
 *   spec      [B, T, F] fp32 row-major (AudioMAE [B,1,T,F] and AST [B,T,F] share this layout)
 *   patches   [B * P, 256] (out_dtype), P = (T/16)*(F/16), row order / column order per `order`
 *   tokens    [B, extra + P, D] fp32: rows 0..extra-1 of every clip are written with
 *             extra_tok[e] + pos[e]  (extra_tok = cls (and dist), [extra, D] fp32; pos [extra+P, D])
 */
int tpat_patchify(const float* spec, void* patches, int out_dtype, float* tokens,
                  const float* extra_tok, const float* pos, int B, int T, int F, int D,
                  int num_extra, int order, tpat_stream_t stream);

/*
 * Per-patch statistics of the spectrogram, in token order: the ablation ranking vectors.
 * Replaces `rearrange(x, 'b c (h p) (w q) -> b (c p q) (h w)', p=16, q=16).mean(dim=1)` / `.std(dim=1)`
 * (unbiased, n-1) of models_vit.py:345-349,354-355 and ast_models.py:447-451,456-457.
 *   spec [B, T, F] fp32;  mean / std [B, P] fp32, either may be NULL;  order as for tpat_patchify.
 */
int tpat_patch_stats(const float* spec, float* mean, float* std, int B, int T, int F, int order,
                     tpat_stream_t stream);

/*
 * out[b, i] = rank[b, idx[b, i]]: carries a custom ranking vector through a pruning block
 * (`custom_rank = torch.gather(custom_rank, dim=1, index=topk_idx)`, models_vit.py:373-374, ast_models.py:482-483).
 *   rank [B, n] fp32, idx [B, k] int64 with values in [0, n), out [B, k] fp32.
 */
int tpat_gather_rank(const float* rank, const int64_t* idx, float* out, int B, int n, int k,
                     tpat_stream_t stream);

/*
 * LayerNorm over the last dim.  Replaces nn.LayerNorm (models_vit.py:197,205; ast_models.py:209,217,500).
 *   x [rows, D] fp32 -> y [rows, D] (y_dtype).  D % 128 == 0, D <= 2048.
 *   y_dtype TPAT_BF16_SPLIT3 (D in {384, 768, 1024}): y is [rows, 3 * D] bf16 = [hi | lo | hi]; against weights laid out
 *   [w_hi | w_hi | w_lo] a K = 3 * D bf16 GEMM yields y w to ~2^-16 relative (split-bf16, "bf16x3"): the q / k projection of
 *   the pruning blocks in the "bf16+score32" precision mode.  Segment 0 is the plain bf16 output (lda 3 D).
 */
int tpat_layernorm(const float* x, const float* gamma, const float* beta, void* y, int y_dtype,
                   int rows, int D, float eps, tpat_stream_t stream);

/*
 * C[M, N] = epilogue(A[M, K] * W[N, K]^T + bias[N]).  Replaces nn.Linear + the elementwise op
 * that follows it: qkv (models_vit.py:76), proj + residual (:96,198), fc1 + GELU (:41-42),
 * fc2 + residual (:44,205), patch-embed conv as GEMM + pos add (:246,358), head (:522).
 *   A [M, K] row-major, lda elements between rows, a_dtype; W [N, K] row-major (nn.Linear layout),
 *   w_dtype (must equal a_dtype); bias fp32 [N] or NULL; C [*, ldc] c_dtype.
 *   TPAT_EPI_BIAS_RESIDUAL: residual fp32 [M, ldr], C fp32.
 *   TPAT_EPI_BIAS_POS: row m = b*P + p is written to row b*(extra+P) + extra + p and pos[(extra+p), :]
 *     (fp32 [extra+P, N]) is added; pass P and extra, C fp32.
 *   impl TPAT_IMPL_TC requires bf16 operands, K % 64 == 0, N % 16 == 0, 16-byte aligned pointers.
 */
int tpat_gemm(const void* A, int a_dtype, int lda, const void* W, int w_dtype, const float* bias,
              void* C, int c_dtype, int ldc, const float* residual, int ldr, const float* pos,
              int P, int num_extra, int M, int N, int K, int epilogue, int impl, tpat_stream_t stream);

/*
 * tpat_gemm with the LayerNorm that FOLLOWS a residual GEMM folded into its neighbours (tcgen05 path only), so that
 * `x = x + f(...)` followed by `norm(x)` and the next nn.Linear (models_vit.py:198,205 then :197/:205 of the next
 * sub-block) need no separate pass over x:
 *   producer  (TPAT_EPI_BIAS_RESIDUAL): besides C it writes xb = bf16(C) and, for every row and 32-column chunk, the
 *             partial moments part_out[m][n/32] = (sum, sum of squared deviations from the chunk mean);
 *   consumer  (TPAT_EPI_BIAS, TPAT_EPI_BIAS_GELU): A = xb, W = bf16(W * gamma) (gamma along K), bias = W beta + b,
 *             ln_colsum[n] = sum_k W'[n, k]; the epilogue computes
 *             rstd[m] * (acc[m, n] - mean[m] * ln_colsum[n]) + bias[n]   with mean / rstd of row m combined from
 *             ln_part[m][0 .. K/32) (Chan's update in a fixed order: deterministic), variance biased, eps = ln_eps.
 * fold == NULL or all pointers NULL: identical to tpat_gemm.
 */
typedef struct tpat_ln_fold {
  void* xb;                /* producer: bf16 [M, ldxb] or NULL */
  int ldxb;
  float* part_out;         /* producer: fp32 [M, N/32, 2] */
  const float* ln_part;    /* consumer: fp32 [M, K/32, 2] or NULL */
  const float* ln_colsum;  /* consumer: fp32 [N] */
  float ln_eps;
} tpat_ln_fold;
int tpat_gemm_ln(const void* A, int a_dtype, int lda, const void* W, int w_dtype, const float* bias,
                 void* C, int c_dtype, int ldc, const float* residual, int ldr, const float* pos,
                 int P, int num_extra, int M, int N, int K, int epilogue, int impl, const tpat_ln_fold* fold,
                 tpat_stream_t stream);

/*
 * Fused multi-head attention that also emits the importance-score partials.
 * Replaces: q k^T * scale, softmax, attn @ v, transpose/reshape (models_vit.py:79-95) and the
 * score slice/mean over the materialised [B,H,N,N] matrix (:113; ast_models.py:124).
 *   qkv   [B * N, 3 * H * hd] (dtype): the qkv Linear output, column = which*H*hd + h*hd + d
 *   out   [B * N, H * hd] (dtype)
 *   score_partial  fp32, NULL when score_mode == NONE.
 *       CLS_ROW : [B, H, N]            row 0 of P per head
 *       COLMEAN : [B, H * n_qt, N]     per (head, query tile) column sums of P over rows >= extra,
 *                 n_qt = tpat_attention_qtiles(N, impl)
 *   softmax is exact (max-subtracted, normalised over all N keys), fp32.
 */
int tpat_attention_qtiles(int N, int impl);
/* fp32 [rows, cols] -> split-bf16 planes [rows, 2 * cols] = [hi | lo]  (cols % 4 == 0) */
int tpat_split_bf16(const float* x, void* out, int rows, int cols, tpat_stream_t stream);
/*
 * tpat_attention whose SCORE tiles (all tiles for COLMEAN, the cls tile for CLS_ROW) take q and k as split-bf16 planes
 *   qk_planes [B * N, 4 * H * hd] bf16 = [q_hi k_hi | q_lo k_lo]  (tpat_split_bf16 of the fp32 q | k projection)
 * and form S = Q_hi K_hi^T + Q_hi K_lo^T + Q_lo K_hi^T on the tensor cores (scores to ~2^-16 instead of 2^-8 relative:
 * models_vit.py:113-114 / ast_models.py:124-125 decide the kept tokens from them).  v, and q / k of the non-score tiles,
 * still come from `qkv`.  qk_planes == NULL: identical to tpat_attention.  tcgen05 path only.
 */
int tpat_attention_split(const void* qkv, const void* qk_planes, void* out, int dtype, float* score_partial,
                         int score_mode, int B, int N, int H, int hd, int num_extra, float scale, int impl,
                         tpat_stream_t stream);
int tpat_attention(const void* qkv, void* out, int dtype, float* score_partial, int score_mode,
                   int B, int N, int H, int hd, int num_extra, float scale, int impl,
                   tpat_stream_t stream);

/*
 * Reduce the score partials in a fixed order (deterministic), write the score, and select the
 * top-k tokens.  Replaces `.mean(...)` + torch.topk(score, k, largest=True, sorted=True)
 * (models_vit.py:113-114, ast_models.py:124-125).
 *   partial  [B, R, N] fp32 (R rows to sum per clip), score[b, j] = sum_r partial[b, r, extra+j] / divisor
 *   score    [B, N - extra] fp32 (may be NULL)
 *   topk_idx [B, k] int64, descending score; ties: lower index first (torch leaves ties
 *            unspecified).  NaN ranks above every number, like torch.topk.
 *            k == 0 or topk_idx == NULL: only the score is written.
 *   rest_idx [B, (N - extra) - k] int32 or NULL: the tokens that were NOT kept, in descending-score order
 *            (input of tpat_fuse_token).
 */
int tpat_score_topk(const float* partial, int R, float divisor, float* score, int64_t* topk_idx,
                    int32_t* rest_idx, int B, int N, int num_extra, int k, tpat_stream_t stream);

/*
 * Token gather / compaction fused with the LayerNorm that follows it.
 * Replaces torch.gather + torch.cat (models_vit.py:200-203) and norm2 (:205).
 *   x      [B, N_in, D] fp32   residual stream after the attention residual add
 *   x_out  [B, extra + k, D] fp32: rows < extra copied, row extra+j = x[b, extra + idx[b, j]]
 *   y_out  [B, extra + k, D] (y_dtype) = LayerNorm(x_out) ; NULL to skip
 *   out_rows  rows per clip of x_out / y_out (0 = extra + k); extra + k + 1 leaves room for the fused token
 */
int tpat_gather_layernorm(const float* x, const int64_t* topk_idx, float* x_out, const float* gamma,
                          const float* beta, void* y_out, int y_dtype, int B, int N_in, int k,
                          int num_extra, int out_rows, int D, float eps, tpat_stream_t stream);

/*
 * EViT "fused inattentive token": x_out[b, out_row] = sum_r score[b, rest_idx[b, r]] * x[b, extra + rest_idx[b, r]]
 * (un-normalised weighted sum over the tokens that were not kept) and y_out[b, out_row] = LayerNorm of it.
 * NOT in the reference forward (only util/token_reduction_utils.py:8-19 knows a fuse_token flag);
 * semantics follow upstream EViT (Block.forward: extra_token = sum(non_topk * non_topk_attn)).  Parity unpinned.
 *   score [B, N_in - extra] fp32 (tpat_score_topk output), rest_idx [B, n_rest] int32.
 */
int tpat_fuse_token(const float* x, const float* score, const int32_t* rest_idx, float* x_out,
                    const float* gamma, const float* beta, void* y_out, int y_dtype, int B, int N_in,
                    int n_rest, int out_rows, int out_row, int num_extra, int D, float eps, tpat_stream_t stream);

/*
 * Pooled classifier input.  Replaces
 *   AudioMAE: x[:, 1:].mean(1) -> fc_norm            (models_vit.py:388-389)
 *   AST     : v.norm -> (x[:,0]+x[:,1])/2 -> mlp_head[0] LayerNorm   (ast_models.py:500-503)
 *   x [B, N, D] fp32 -> pooled [B, D] fp32.  AST uses (g1,b1,eps1) = v.norm, (g2,b2,eps2) = mlp_head.0;
 *   AudioMAE uses (g1,b1,eps1) = fc_norm and ignores the second set.
 */
int tpat_pool_norm(const float* x, float* pooled, const float* g1, const float* b1, float eps1,
                   const float* g2, const float* b2, float eps2, int B, int N, int D, int variant,
                   tpat_stream_t stream);

/*
 * Classifier head: logits[B, C] = pooled[B, D] W[C, D]^T + bias, fp32 CUDA cores (tiny, latency-bound).
 * Replaces self.head (models_vit.py:522) / mlp_head[1] (ast_models.py:503).  D % 128 == 0, D <= 1024.
 */
int tpat_head(const float* pooled, const float* W, const float* bias, float* logits, int B, int D, int C,
              tpat_stream_t stream);

/* ---- whole forward (the hot loop of models_vit.py:365-385 / ast_models.py:470-497 in native code) ---- */

typedef struct {
  const float* ln1_g; const float* ln1_b;
  const void* qkv_w;  const float* qkv_b;   /* [3D, D] */
  const void* proj_w; const float* proj_b;  /* [D, D]  */
  const float* ln2_g; const float* ln2_b;
  const void* fc1_w;  const float* fc1_b;   /* [Dh, D] */
  const void* fc2_w;  const float* fc2_b;   /* [D, Dh] */
  /* Optional (TPAT_IMPL_TC only; NULL = keep the separate LayerNorm kernels): the LayerNorm fold of tpat_gemm_ln.
   *   qkv_w_ln = bf16(qkv_w * ln1_g[k]), qkv_colsum[n] = sum_k qkv_w_ln[n, k], qkv_b_ln = qkv_w ln1_b + qkv_b;
   *   fc1_* likewise with ln2_g / ln2_b.  With them, norm1 of block i > 0 and norm2 of non-pruning blocks cost no
   *   pass over x: fc2 / proj emit bf16(x) + partial moments, qkv / fc1 normalise in their epilogue. */
  const void* qkv_w_ln;  const float* qkv_colsum;  const float* qkv_b_ln;
  const void* fc1_w_ln;  const float* fc1_colsum;  const float* fc1_b_ln;
  /* Optional (TPAT_IMPL_TC, args.score32): [2D, 3D] bf16 = [w_hi | w_hi | w_lo] of the q and k rows of qkv_w (split-bf16) */
  const void* qk_w_split;
} tpat_block_weights;

typedef struct {
  int variant;        /* TPAT_VARIANT_*                                                       */
  int impl;           /* TPAT_IMPL_SIMT: fp32 weights & activations; TPAT_IMPL_TC: bf16        */
  int B, T, F;        /* clips, time frames, mel bins (F == 128 in the reference)              */
  int depth, D, H, Dh, num_classes;
  int prune[TPAT_MAX_DEPTH]; /* 1: block i runs top-k + gather (reference: keep_rate < 1.0)        */
  int keep[TPAT_MAX_DEPTH];  /* k of the top-k when prune[i], else the (unchanged) non-extra token count;
                                with fuse_token a pruning block hands keep[i] + 1 tokens to the next one   */
  int fuse_token;            /* EViT fused inattentive token appended after the kept tokens (unpinned)    */
  int want_all_scores;       /* extract mode: emit the score of every block                    */
  int score32;               /* TPAT_IMPL_TC only: pruning blocks compute q / k (split-bf16 GEMM) and Q K^T of the score tiles
                                (tpat_attention_split) to ~fp32 accuracy; needs blocks[i].qk_w_split where prune[i]  */
  float ln_eps;              /* 1e-6 for every block norm                                      */
  /* weights: matrices in the impl's operand dtype, vectors fp32 */
  const void* patch_w; const float* patch_b;       /* [D, 256], columns in the order of `variant` */
  const float* extra_tok;                          /* [extra, D]: cls (, dist)                  */
  const float* pos;                                /* [extra + P, D]                            */
  tpat_block_weights blocks[TPAT_MAX_DEPTH];
  const float* norm_g; const float* norm_b; float norm_eps;   /* fc_norm (AudioMAE) / v.norm (AST) */
  const float* head_ln_g; const float* head_ln_b; float head_ln_eps;  /* AST mlp_head.0; NULL for AudioMAE */
  const float* head_w; const float* head_b;        /* [C, D] fp32                               */
  /* inputs / outputs */
  const float* spec;          /* [B, T, F] fp32                                                 */
  float* logits;              /* [B, C] fp32                                                    */
  float* scores[TPAT_MAX_DEPTH];     /* [B, n_i] fp32 or NULL (n_i = non-extra tokens entering block i) */
  int64_t* topk_idx[TPAT_MAX_DEPTH]; /* [B, keep[i]] int64; required where prune[i]                  */
  void* workspace; size_t workspace_bytes;
  float* pooled;              /* optional [B, D] fp32: the classifier input (AudioMAE: fc_norm(mean of tokens 1:),
                                 i.e. what forward_features returns, models_vit.py:387-389; AST: mlp_head.0 of the
                                 (cls + dist) / 2 of v.norm) -- NULL: kept in the workspace only */
} tpat_forward_args;

/* sizeof(tpat_forward_args) as compiled into the library (lets a foreign-language binding verify its struct layout) */
size_t tpat_sizeof_forward_args(void);
size_t tpat_forward_workspace_bytes(const tpat_forward_args* args);
int tpat_forward(const tpat_forward_args* args, tpat_stream_t stream);
/* number of kernels tpat_forward launches for `args` (for bench.py's gpu_launches) */
int tpat_forward_launch_count(const tpat_forward_args* args);

/* ======================================================================================================
 * Fine-tune step (BASELINE.json configs[3]): forward that keeps what the backward
 * needs, and the backward of every op of the path.  Replaces PyTorch autograd over Block.forward
 * (audiomae/models_vit.py:191-207), Attention.forward (:68-135), Mlp (:40-46), PatchEmbed (:241-247),
 * forward_features(_mask) (:334-396,466-497) and the head (:522), as driven by train_one_epoch
 * (audiomae/engine_finetune.py:102-105 forward + loss, util/misc.py:259-273 backward + optimizer step).
 * ====================================================================================================== */

/* extras of tpat_gemm_train: all optional */
typedef struct tpat_gemm_extra {
  void* dact_out; int ld_dact;                   /* TPAT_EPI_BIAS_GELU: also store gelu'(A W^T + b) (dtype of C) for the backward    */
  const void* aux; int ld_aux;                   /* TPAT_EPI_DGELU: that saved derivative (dtype of C); C = (A W^T) * aux           */
  const float* row_scale; int rows_per_clip;     /* TPAT_EPI_BIAS_RESIDUAL: C = R + row_scale[m / rows_per_clip] * (A W^T + b):
                                                    timm DropPath's per-sample scale 0 | 1 / keep_prob (models_vit.py:149,198,205)   */
  int w_kn;                                      /* 1: W is stored [K, N] row-major (C = A W): the data gradient dX = dY W reads the
                                                    forward weight [out, in] as it is, as an MN-major tcgen05 B operand -- no
                                                    transposed copy.  tcgen05 path only (N % 8 == 0)                                   */
  float* colsum_out;                             /* TPAT_EPI_DGELU: colsum_out[n] += sum_m C[m, n] -- the bias gradient of the Linear
                                                    whose dY this GEMM produces (fc1).  tcgen05 path: summed in the epilogue from
                                                    the fp32 values (per 32-row partials in colsum_ws, fixed-order finish: no extra
                                                    pass over C); otherwise, or when colsum_ws is too small, one tpat_colsum pass     */
  float* colsum_ws; size_t colsum_ws_floats;     /* >= tpat_gemm_colsum_ws_floats(M, N) for the fused form                           */
} tpat_gemm_extra;
size_t tpat_gemm_colsum_ws_floats(int M, int N);
int tpat_gemm_train(const void* A, int a_dtype, int lda, const void* W, int w_dtype, const float* bias, void* C, int c_dtype,
                    int ldc, const float* residual, int ldr, int M, int N, int K, int epilogue, int impl,
                    const tpat_gemm_extra* extra, tpat_stream_t stream);

/* C[M, N] (+)= op(A) op(B), fp32 CUDA cores, fixed summation order.  trans_a: A is stored [K, M]; trans_b: B is stored
 * [N, K].  The fp32-parity weight gradient dW += dY^T X (trans_a = 1) and the classifier-head GEMMs of every mode. */
int tpat_gemm_f32(const float* A, int lda, int trans_a, const float* B, int ldb, int trans_b, float* C, int ldc,
                  int M, int N, int K, int accumulate, tpat_stream_t stream);

/* Weight gradient on the tensor cores: dW[Mo, No] (fp32) += dY[K, Mo]^T X[K, No], bf16 operands read in their natural
 * row-major layout (MN-major tcgen05 operands), reduction over the K tokens split across SM pairs, partial tiles added
 * with TMA reduce operations.  Mo % 256 == 0, No % 256 == 0. */
int tpat_gemm_wgrad(const void* dY, int ld_dy, const void* X, int ldx, float* dW, int ldw, int K, int Mo, int No,
                    tpat_stream_t stream);

/* dst[c][r] = cast(src[r][c]); destination rows padded with zeros up to ld_dst (>= rows).  dtype pairs f32->f32,
 * f32->bf16, bf16->bf16. */
int tpat_transpose(const void* src, int src_dtype, int ld_src, void* dst, int dst_dtype, int ld_dst, int rows, int cols,
                   tpat_stream_t stream);

/* tpat_attention that also writes lse[B, H, N] = log sum_j exp(scale * q_i . k_j) (fp32) for the backward */
int tpat_attention_train(const void* qkv, void* out, int dtype, float* score_partial, int score_mode, float* lse,
                         int B, int N, int H, int hd, int num_extra, float scale, int impl, tpat_stream_t stream);
/* Attention backward: dqkv [B * N, 3 * H * hd] (dtype) from qkv, out (= O), d_out and lse; P is recomputed, nothing of
 * size N x N touches HBM; no gradient through the importance score / top-k (indices).
 * delta_ws: tpat_attention_bwd_ws_floats(B, N, H, hd) floats (row sums of dO o O, and the fp32 dQ accumulator of the
 * tcgen05 kernel, which adds each key tile's contribution with TMA reduce operations).
 * dbias (optional, tcgen05 bf16 path only): dbias[3 * H * hd] += column sums of dqkv, i.e. the bias gradient of the qkv
 * projection, taken where dqkv is produced (dK / dV tiles in shared memory, dQ in the fp32 -> bf16 pass) instead of a
 * separate pass over dqkv. */
size_t tpat_attention_bwd_ws_floats(int B, int N, int H, int hd);
int tpat_attention_bwd(const void* qkv, const void* out, const void* d_out, const float* lse, void* dqkv, int dtype,
                       int B, int N, int H, int hd, float scale, int impl, float* delta_ws, float* dbias, tpat_stream_t stream);

/* floats a `partials_ws` scratch buffer must hold for the three functions below */
size_t tpat_bwd_partials_floats(int D_max);
/* inv[b, idx[b, j]] = j, -1 elsewhere: inverse of a top-k list (idx [B, k] int64 with values in [0, n)) */
int tpat_inverse_index(const int64_t* idx, int32_t* inv, int B, int n, int k, tpat_stream_t stream);
/*
 * One pass over the gradient stream (one warp per OUTPUT row):
 *   g = g_up[src row] (+ LayerNorm backward of dy at the saved LayerNorm input x, when dy != NULL)
 *   src row of output row (b, j): j + src_offset, or -- with inv -- j < extra ? j : extra + inv[b, j - extra]
 *   (inv < 0: the token was pruned after this point: zeros; this is the backward of gather + cat, models_vit.py:200-203)
 *   g_out = g (fp32), gb_out = row_scale[b] * g (gb_dtype; the dY operand of the Linear that produced the branch)
 *   dgamma += sum dy * xhat, dbeta += sum dy, dbias += column sums of gb_out   (any may be NULL; fixed-order reduction)
 */
int tpat_row_bwd(const void* dy, int dy_dtype, const float* x, const float* gamma, const float* g_up, float* g_out,
                 void* gb_out, int gb_dtype, const float* row_scale, const int32_t* inv, float* partials_ws,
                 float* dgamma, float* dbeta, float* dbias, int B, int rows_src, int rows_out, int num_extra,
                 int src_offset, int D, float eps, tpat_stream_t stream);
/* dst[c] += sum_m x[m, c]  (bias gradient of a Linear from its dY); C % 4 == 0, C <= 4096 */
int tpat_colsum(const void* x, int dtype, int ld, int M, int C, float* partials_ws, float* dst, tpat_stream_t stream);
/* out[i] (+)= sum_b x[b * stride + i], i < n  (parameters broadcast over the clips; small column sums) */
int tpat_batch_sum(const float* x, float* out, int B, size_t stride, int n, int accumulate, tpat_stream_t stream);
/* backward of tpat_pool_norm: dx [B, N, D] from dpooled [B, D]; dg1 / db1 (and dg2 / db2 for AST) accumulated */
int tpat_pool_norm_bwd(const float* x, const float* dpooled, float* dx, const float* g1, const float* b1, float eps1,
                       const float* g2, float eps2, float* partials_ws, float* dg1, float* db1, float* dg2, float* db2,
                       int B, int N, int D, int variant, tpat_stream_t stream);

/*
 * torch.optim.AdamW's update over flat buffers (reference optimizer: main_finetune.py:478 on the layer-wise lr-decay
 * groups of util/lr_decay.py:15-75): p *= 1 - lr_g wd_g; m, v moments; p -= lr_g / bc1 * m / (sqrt(v) / sqrt(bc2) + eps),
 * g = grad * grad_scale (1 / world size after a SUM all-reduce).  chunks [n_chunks][4] int32 = (offset, length, group, 0),
 * groups [n_groups][2] fp32 = (lr scale, weight decay).  p_bf16 (optional): bf16 copy of p refreshed in the same pass.
 * step_dev (optional): the step count t of the bias corrections 1 - beta^t lives in device memory (tpat_counter_inc
 * advances it), so that a CUDA graph of the whole step can be replayed; NULL: `step` (>= 1) is used.
 */
int tpat_adamw(float* p, const float* g, float* m, float* v, void* p_bf16, const int32_t* chunks, int n_chunks,
               const float* groups, float lr, float beta1, float beta2, float eps, int step, const int32_t* step_dev,
               float grad_scale, tpat_stream_t stream);
int tpat_counter_inc(int32_t* counter, tpat_stream_t stream);

typedef struct {
  float* ln1_g; float* ln1_b; float* qkv_w; float* qkv_b; float* proj_w; float* proj_b;
  float* ln2_g; float* ln2_b; float* fc1_w; float* fc1_b; float* fc2_w; float* fc2_b;
} tpat_block_grads;
typedef struct { const void* qkv_wt; const void* proj_wt; const void* fc1_wt; const void* fc2_wt; } tpat_block_wt;

typedef struct {
  tpat_forward_args fwd;     /* model, weights, input, logits / scores / topk_idx, workspace (tpat_forward's fields; fuse_token,
                                score32 and the LayerNorm fold are not available in training) */
  void* saved; size_t saved_bytes;                 /* activations kept for the backward (tpat_train_saved_bytes) */
  const float* drop_scale[TPAT_MAX_DEPTH][2];      /* DropPath: per block, [0] attention branch, [1] MLP branch: [B] fp32 of
                                                      0 | 1 / keep_prob, or NULL (no drop)                          */
  const int64_t* mask_keep_idx; int n_keep;        /* fine-tune 2-D masking: patch tokens kept right after the patch embedding
                                                      ([B, n_keep], models_vit.py:425-497), or NULL                 */
  /* ---- backward ---- */
  const float* dlogits;                            /* [B, C] fp32                                                  */
  tpat_block_wt wt[TPAT_MAX_DEPTH];                /* [in, out] copies of the four matrices, operand dtype (dX = dY W) */
  tpat_block_grads grads[TPAT_MAX_DEPTH];          /* fp32, ACCUMULATED into (the caller zeroes them); NULL = skip  */
  float* d_patch_w; float* d_patch_b; float* d_extra_tok; float* d_pos;   /* d_pos NULL for a frozen pos_embed      */
  float* d_norm_g; float* d_norm_b; float* d_head_ln_g; float* d_head_ln_b; float* d_head_w; float* d_head_b;
  void* bwd_workspace; size_t bwd_workspace_bytes;
} tpat_train_args;

size_t tpat_sizeof_train_args(void);
size_t tpat_train_saved_bytes(const tpat_train_args* args);
size_t tpat_train_bwd_workspace_bytes(const tpat_train_args* args);
/* forward in training mode: same math as tpat_forward (+ DropPath scales, + masking), activations kept in `saved` */
int tpat_train_forward(const tpat_train_args* args, tpat_stream_t stream);
/* backward stages hi .. lo (inclusive, descending): depth + 1 = head + pooling, i + 1 = block i, 0 = patch embedding.
 * A full backward is (depth + 1, 0); a caller overlapping the gradient all-reduce with compute (one bucket per block,
 * main_finetune.py:459-461's DDP) calls it stage by stage.  The gradient stream lives in bwd_workspace between calls. */
int tpat_train_backward(const tpat_train_args* args, int stage_hi, int stage_lo, tpat_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TPAT_H */
