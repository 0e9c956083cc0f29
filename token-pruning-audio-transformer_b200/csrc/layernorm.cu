// Vectorised LayerNorm, token gather + LayerNorm, and the pooled classifier-input kernel.
//
// All three are HBM-bound (no reuse): one warp owns one token row, reads it once with 128-bit
// loads (lane l reads float4 #l, #l+32, ... so every warp instruction covers 512 contiguous
// bytes), keeps it in registers for the two-pass mean / variance, and writes the result once.
// Algorithmic bytes per row: 4*D read + (2|4)*D written (+ 4*D for the compacted copy in the
// gather variant).  Reference: nn.LayerNorm at audiomae/models_vit.py:197,205,389 and
// ast/src/models/ast_models.py:209,217,500; torch.gather + torch.cat at models_vit.py:200-203.
#include "common.cuh"

namespace tpat {

constexpr int LN_WARPS = 8;

template <int NV>
__device__ __forceinline__ void ln_row_load(const float* __restrict__ row, int lane, float4 (&v)[NV]) {
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = __ldg(reinterpret_cast<const float4*>(row) + lane + 32 * i);
}

template <int NV>
__device__ __forceinline__ void ln_row_stats(const float4 (&v)[NV], float inv_d, float eps, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  mean = warp_sum(s) * inv_d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float var = warp_sum(q) * inv_d;
  rstd = 1.0f / sqrtf(var + eps);
}

template <int NV, typename OutT>
__device__ __forceinline__ void ln_row_store(const float4 (&v)[NV], float mean, float rstd,
                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                             OutT* __restrict__ out, int lane) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c4 = lane + 32 * i;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c4);
    const float y0 = (v[i].x - mean) * rstd * g.x + b.x;
    const float y1 = (v[i].y - mean) * rstd * g.y + b.y;
    const float y2 = (v[i].z - mean) * rstd * g.z + b.z;
    const float y3 = (v[i].w - mean) * rstd * g.w + b.w;
    if constexpr (sizeof(OutT) == 4) {
      reinterpret_cast<float4*>(out)[c4] = make_float4(y0, y1, y2, y3);
    } else {
      reinterpret_cast<uint2*>(out)[c4] = make_uint2(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3));
    }
  }
}

// rows -> LayerNorm(rows).  grid = ceil(rows / LN_WARPS), block = 32 * LN_WARPS.
template <int NV, typename OutT>
__global__ void __launch_bounds__(32 * LN_WARPS)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 OutT* __restrict__ y, int rows, float eps, int desc) {
  constexpr int D = NV * 128;
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  int row = blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  if (desc) row = rows - 1 - row;        // walk from the end (see g_walk_desc)
  float4 v[NV];
  ln_row_load<NV>(x + (size_t)row * D, lane, v);
  float mean, rstd;
  ln_row_stats<NV>(v, 1.0f / D, eps, mean, rstd);
  ln_row_store<NV, OutT>(v, mean, rstd, gamma, beta, y + (size_t)row * D, lane);
}

// LayerNorm whose output is written as a split-bf16 triple [hi | lo | hi] (row pitch 3 * D): hi = bf16(y),
// lo = bf16(y - hi).  Against weights laid out [w_hi | w_hi | w_lo] a plain K = 3 * D bf16 GEMM then computes
// y_hi w_hi + y_lo w_hi + y_hi w_lo ~ y w to ~2^-16 relative -- the "precision where it matters" path of the
// pruning blocks' q / k projection.  Segment 0 doubles as the ordinary bf16 LayerNorm output
// (lda = 3 * D) for the v projection.
template <int NV>
__global__ void __launch_bounds__(32 * LN_WARPS)
layernorm_split3_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                        __nv_bfloat16* __restrict__ y, int rows, float eps, int desc) {
  constexpr int D = NV * 128;
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  int row = blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  if (desc) row = rows - 1 - row;
  float4 v[NV];
  ln_row_load<NV>(x + (size_t)row * D, lane, v);
  float mean, rstd;
  ln_row_stats<NV>(v, 1.0f / D, eps, mean, rstd);
  uint2* out = reinterpret_cast<uint2*>(y + (size_t)row * 3 * D);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c4 = lane + 32 * i;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c4);
    const float y0 = (v[i].x - mean) * rstd * g.x + b.x, y1 = (v[i].y - mean) * rstd * g.y + b.y;
    const float y2 = (v[i].z - mean) * rstd * g.z + b.z, y3 = (v[i].w - mean) * rstd * g.w + b.w;
    const float h0 = __bfloat162float(__float2bfloat16_rn(y0)), h1 = __bfloat162float(__float2bfloat16_rn(y1));
    const float h2 = __bfloat162float(__float2bfloat16_rn(y2)), h3 = __bfloat162float(__float2bfloat16_rn(y3));
    const uint2 hi = make_uint2(pack_bf16x2(h0, h1), pack_bf16x2(h2, h3));
    const uint2 lo = make_uint2(pack_bf16x2(y0 - h0, y1 - h1), pack_bf16x2(y2 - h2, y3 - h3));
    out[c4] = hi;
    out[D / 4 + c4] = lo;
    out[2 * (D / 4) + c4] = hi;
  }
}

// fp32 [rows, cols] -> split-bf16 planes [rows, 2 * cols] = [hi | lo] (the q / k operands of the split QK^T)
__global__ void __launch_bounds__(256)
split_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int rows, int cols) {
  pdl_trigger();
  pdl_wait();
  const int c4n = cols / 4;
  const size_t total = (size_t)rows * c4n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / c4n; const int c4 = (int)(i - r * c4n);
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    const float h0 = __bfloat162float(__float2bfloat16_rn(v.x)), h1 = __bfloat162float(__float2bfloat16_rn(v.y));
    const float h2 = __bfloat162float(__float2bfloat16_rn(v.z)), h3 = __bfloat162float(__float2bfloat16_rn(v.w));
    uint2* o = reinterpret_cast<uint2*>(out + r * 2 * cols);
    o[c4] = make_uint2(pack_bf16x2(h0, h1), pack_bf16x2(h2, h3));
    o[c4n + c4] = make_uint2(pack_bf16x2(v.x - h0, v.y - h1), pack_bf16x2(v.z - h2, v.w - h3));
  }
}

// Token compaction fused with norm2: output row (b, j) <- input row (b, j < extra ? j : extra + idx[b, j-extra]).
template <int NV, typename OutT>
__global__ void __launch_bounds__(32 * LN_WARPS)
gather_layernorm_kernel(const float* __restrict__ x, const int64_t* __restrict__ idx, float* __restrict__ x_out,
                        const float* __restrict__ gamma, const float* __restrict__ beta, OutT* __restrict__ y_out,
                        int B, int N_in, int k, int num_extra, int out_rows, float eps, int desc) {
  constexpr int D = NV * 128;
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int N_out = num_extra + k;                 // rows this kernel writes per clip; the clip stride is out_rows
  int grow = blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  if (grow >= B * N_out) return;
  if (desc) grow = B * N_out - 1 - grow;
  const int b = grow / N_out, j = grow - b * N_out;
  const size_t orow = (size_t)b * out_rows + j;
  int src = j;
  if (j >= num_extra) src = num_extra + (int)__ldg(idx + (size_t)b * k + (j - num_extra));
  float4 v[NV];
  ln_row_load<NV>(x + ((size_t)b * N_in + src) * D, lane, v);
  float4* xo = reinterpret_cast<float4*>(x_out + orow * D);
#pragma unroll
  for (int i = 0; i < NV; ++i) xo[lane + 32 * i] = v[i];
  if (y_out != nullptr) {
    float mean, rstd;
    ln_row_stats<NV>(v, 1.0f / D, eps, mean, rstd);
    ln_row_store<NV, OutT>(v, mean, rstd, gamma, beta, y_out + orow * D, lane);
  }
}

// ---- pooled classifier input: one CTA per clip ----
__device__ __forceinline__ float block_sum_256(float v, float* red) {   // any block size up to 1024 threads
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (l < (int)(blockDim.x >> 5)) ? red[l] : 0.f;
  t = warp_sum(t);
  return t;
}

// LayerNorm of a D-vector held in shared memory (in place), block-cooperative.
__device__ void block_layernorm_inplace(float* vec, int D, const float* __restrict__ g, const float* __restrict__ b,
                                        float eps, float* red) {
  float s = 0.f;
  for (int c = threadIdx.x; c < D; c += blockDim.x) s += vec[c];
  const float mean = block_sum_256(s, red) / D;
  float q = 0.f;
  for (int c = threadIdx.x; c < D; c += blockDim.x) { const float d = vec[c] - mean; q += d * d; }
  const float var = block_sum_256(q, red) / D;
  const float rstd = 1.0f / sqrtf(var + eps);
  for (int c = threadIdx.x; c < D; c += blockDim.x) vec[c] = (vec[c] - mean) * rstd * g[c] + b[c];
  __syncthreads();
}

// grid = B clips, block = 1024 threads.  AudioMAE: thread (tg, c4) sums the float4 column c4 over
// tokens t = 1 + tg, 1 + tg + TG, ... (coalesced 3 KB rows, 4 independent accumulators per thread for
// memory-level parallelism), the TG partial sums are combined through shared memory in a fixed order.
__global__ void __launch_bounds__(1024)
pool_norm_kernel(const float* __restrict__ x, float* __restrict__ pooled, const float* __restrict__ g1,
                 const float* __restrict__ b1, float eps1, const float* __restrict__ g2,
                 const float* __restrict__ b2, float eps2, int N, int D, int variant) {
  extern __shared__ float sm[];  // [TG*D (pool partials) | 2*D | 8]
  pdl_trigger();
  pdl_wait();
  const float* xb = x + (size_t)blockIdx.x * N * D;
  const int nv = D / 4;                       // float4 columns
  const int TG = blockDim.x / nv;             // token groups (>= 1)
  float* part = sm;                           // [TG][D]
  float* v0 = sm + (size_t)TG * D;
  float* v1 = v0 + D;
  float* red = v1 + D;
  if (variant == TPAT_VARIANT_AUDIOMAE) {
    // x[:, 1:, :].mean(dim=1) -> fc_norm   (models_vit.py:388-389)
    const int tg = threadIdx.x / nv, c4 = threadIdx.x - tg * nv;
    if (tg < TG) {
      float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
      int t = 1 + tg;
      for (; t + 3 * TG < N; t += 4 * TG) {
        const float4 u0 = __ldg(reinterpret_cast<const float4*>(xb + (size_t)t * D) + c4);
        const float4 u1 = __ldg(reinterpret_cast<const float4*>(xb + (size_t)(t + TG) * D) + c4);
        const float4 u2 = __ldg(reinterpret_cast<const float4*>(xb + (size_t)(t + 2 * TG) * D) + c4);
        const float4 u3 = __ldg(reinterpret_cast<const float4*>(xb + (size_t)(t + 3 * TG) * D) + c4);
        a0.x += u0.x; a0.y += u0.y; a0.z += u0.z; a0.w += u0.w;
        a1.x += u1.x; a1.y += u1.y; a1.z += u1.z; a1.w += u1.w;
        a2.x += u2.x; a2.y += u2.y; a2.z += u2.z; a2.w += u2.w;
        a3.x += u3.x; a3.y += u3.y; a3.z += u3.z; a3.w += u3.w;
      }
      for (; t < N; t += TG) {
        const float4 u0 = __ldg(reinterpret_cast<const float4*>(xb + (size_t)t * D) + c4);
        a0.x += u0.x; a0.y += u0.y; a0.z += u0.z; a0.w += u0.w;
      }
      reinterpret_cast<float4*>(part + (size_t)tg * D)[c4] =
          make_float4((a0.x + a1.x) + (a2.x + a3.x), (a0.y + a1.y) + (a2.y + a3.y), (a0.z + a1.z) + (a2.z + a3.z),
                      (a0.w + a1.w) + (a2.w + a3.w));
    }
    __syncthreads();
    const float inv = 1.0f / (float)(N - 1);
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      float s = 0.f;
      for (int g = 0; g < TG; ++g) s += part[(size_t)g * D + c];
      v0[c] = s * inv;
    }
    __syncthreads();
    block_layernorm_inplace(v0, D, g1, b1, eps1, red);
  } else {
    // v.norm on rows 0 (cls) and 1 (dist) -- LayerNorm is per row, the other rows are never
    // read by the head -- average, then mlp_head[0] LayerNorm   (ast_models.py:500-503)
    for (int c = threadIdx.x; c < D; c += blockDim.x) { v0[c] = xb[c]; v1[c] = xb[D + c]; }
    __syncthreads();
    block_layernorm_inplace(v0, D, g1, b1, eps1, red);
    block_layernorm_inplace(v1, D, g1, b1, eps1, red);
    for (int c = threadIdx.x; c < D; c += blockDim.x) v0[c] = (v0[c] + v1[c]) / 2.0f;
    __syncthreads();
    block_layernorm_inplace(v0, D, g2, b2, eps2, red);
  }
  for (int c = threadIdx.x; c < D; c += blockDim.x) pooled[(size_t)blockIdx.x * D + c] = v0[c];
}

// Classifier head: logits[b, c] = pooled[b, :] . W[c, :] + bias[c].  One warp per class keeps W[c, :] in
// registers (D <= 1024 -> 32 floats per lane) and loops over the clips; pooled rows come from L1/L2.
// Replaces self.head / mlp_head[1] (models_vit.py:522; ast_models.py:503).  fp32, fixed summation order.
template <int NV>
__global__ void __launch_bounds__(256)
head_kernel(const float* __restrict__ pooled, const float* __restrict__ W, const float* __restrict__ bias,
            float* __restrict__ logits, int B, int C) {
  constexpr int D = NV * 128;
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= C) return;
  float4 w[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) w[i] = __ldg(reinterpret_cast<const float4*>(W + (size_t)c * D) + lane + 32 * i);
  const float bc = bias ? __ldg(bias + c) : 0.f;
  // the clips are split over blockIdx.y (more, shorter dependency chains: this kernel is pure latency);
  // four clips per step: four independent load / FMA / shuffle chains in flight
  const int per = ((B + (int)gridDim.y - 1) / (int)gridDim.y + 3) & ~3;
  const int b_end = min(B, ((int)blockIdx.y + 1) * per);
  for (int b0 = (int)blockIdx.y * per; b0 < b_end; b0 += 4) {
    float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int b = b0 + u < B ? b0 + u : B - 1;
      const float4* pr = reinterpret_cast<const float4*>(pooled + (size_t)b * D);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float4 pv = __ldg(pr + lane + 32 * i);
        s[u] = fmaf(pv.x, w[i].x, s[u]); s[u] = fmaf(pv.y, w[i].y, s[u]); s[u] = fmaf(pv.z, w[i].z, s[u]); s[u] = fmaf(pv.w, w[i].w, s[u]);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int u = 0; u < 4; ++u) s[u] += __shfl_xor_sync(0xffffffffu, s[u], o);
    }
    if (lane < 4 && b0 + lane < B) logits[(size_t)(b0 + lane) * C + c] = (lane == 0 ? s[0] : lane == 1 ? s[1] : lane == 2 ? s[2] : s[3]) + bc;
  }
}


// ---- EViT fused inattentive token: one CTA per clip ----
// fused[b, :] = sum_r score[b, rest[b, r]] * x[b, extra + rest[b, r], :]   (un-normalised weighted sum, as in EViT's
// Block.forward: extra_token = sum(non_topk * non_topk_attn)), written as output row `out_row` of the clip together
// with its LayerNorm.  The reference repository does NOT implement this (parity unpinned); the
// semantics follow upstream EViT and oracle/vit_oracle.py restates them.  Fixed summation order: thread (tg, c4)
// takes rest rows tg, tg + TG, ...; the TG partials are added in order.
template <typename OutT>
__global__ void __launch_bounds__(1024)
fuse_token_kernel(const float* __restrict__ x, const float* __restrict__ score, const int32_t* __restrict__ rest_idx,
                  float* __restrict__ x_out, const float* __restrict__ gamma, const float* __restrict__ beta,
                  OutT* __restrict__ y_out, int N_in, int n_rest, int out_rows, int out_row, int num_extra, int D, float eps) {
  extern __shared__ float sm[];  // [TG*D | D | 32]
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x;
  const int nv = D / 4, TG = blockDim.x / nv;
  float* part = sm;
  float* v0 = sm + (size_t)TG * D;
  float* red = v0 + D;
  const float* xb = x + (size_t)b * N_in * D;
  const float* sb = score + (size_t)b * (N_in - num_extra);
  const int32_t* rb = rest_idx + (size_t)b * n_rest;
  const int tg = threadIdx.x / nv, c4 = threadIdx.x - tg * nv;
  if (tg < TG) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = tg; r < n_rest; r += TG) {
      const int t = __ldg(rb + r);
      const float w = __ldg(sb + t);
      const float4 u = __ldg(reinterpret_cast<const float4*>(xb + (size_t)(num_extra + t) * D) + c4);
      a.x = fmaf(w, u.x, a.x); a.y = fmaf(w, u.y, a.y); a.z = fmaf(w, u.z, a.z); a.w = fmaf(w, u.w, a.w);
    }
    reinterpret_cast<float4*>(part + (size_t)tg * D)[c4] = a;
  }
  __syncthreads();
  float* xo = x_out + ((size_t)b * out_rows + out_row) * D;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float s = 0.f;
    for (int g = 0; g < TG; ++g) s += part[(size_t)g * D + c];
    v0[c] = s;
    xo[c] = s;
  }
  __syncthreads();
  if (y_out != nullptr) {
    block_layernorm_inplace(v0, D, gamma, beta, eps, red);
    OutT* yo = y_out + ((size_t)b * out_rows + out_row) * D;
    for (int c = threadIdx.x; c < D; c += blockDim.x) yo[c] = from_f32<OutT>(v0[c]);
  }
}

}  // namespace tpat

extern "C" int tpat_fuse_token(const float* x, const float* score, const int32_t* rest_idx, float* x_out,
                               const float* gamma, const float* beta, void* y_out, int y_dtype, int B, int N_in,
                               int n_rest, int out_rows, int out_row, int num_extra, int D, float eps,
                               tpat_stream_t stream) {
  using namespace tpat;
  TPAT_CHECK(x && score && rest_idx && x_out, "tpat_fuse_token: null pointer");
  TPAT_CHECK(y_out == nullptr || (gamma && beta), "tpat_fuse_token: y_out needs gamma and beta");
  TPAT_CHECK(B >= 0 && n_rest >= 0 && n_rest <= N_in - num_extra && out_row >= 0 && out_row < out_rows,
             "tpat_fuse_token: bad sizes N_in=%d n_rest=%d out_row=%d out_rows=%d", N_in, n_rest, out_row, out_rows);
  TPAT_CHECK(D > 0 && D % 4 == 0 && D <= 2048, "tpat_fuse_token: unsupported D=%d", D);
  TPAT_CHECK(y_dtype == TPAT_F32 || y_dtype == TPAT_BF16, "tpat_fuse_token: bad dtype %d", y_dtype);
  if (B == 0) return 0;
  const int threads = 1024, TG = threads / (D / 4);
  const size_t smem = ((size_t)TG * D + D + 32) * sizeof(float);
  if (y_dtype == TPAT_F32)
    TPAT_CUDA(launch_kernel(fuse_token_kernel<float>, dim3(B), dim3(threads), smem, as_stream(stream), x, score, rest_idx, x_out, gamma, beta,
                            (float*)y_out, N_in, n_rest, out_rows, out_row, num_extra, D, eps));
  else
    TPAT_CUDA(launch_kernel(fuse_token_kernel<__nv_bfloat16>, dim3(B), dim3(threads), smem, as_stream(stream), x, score, rest_idx, x_out, gamma, beta,
                            (__nv_bfloat16*)y_out, N_in, n_rest, out_rows, out_row, num_extra, D, eps));
  TPAT_LAUNCH_CHECK();
  return 0;
}

namespace tpat {
template <typename OutT>
static int launch_layernorm(const float* x, const float* g, const float* b, OutT* y, int rows, int D, float eps,
                            cudaStream_t st) {
  const int grid = (rows + LN_WARPS - 1) / LN_WARPS;
#define TPAT_LN_CASE(nv) \
  case nv: TPAT_CUDA(launch_kernel(layernorm_kernel<nv, OutT>, dim3(grid), dim3(32 * LN_WARPS), 0, st, x, g, b, y, rows, eps, g_walk_desc)); break;
  switch (D / 128) {
    TPAT_LN_CASE(1) TPAT_LN_CASE(2) TPAT_LN_CASE(3) TPAT_LN_CASE(4) TPAT_LN_CASE(5) TPAT_LN_CASE(6)
    TPAT_LN_CASE(8) TPAT_LN_CASE(10) TPAT_LN_CASE(12) TPAT_LN_CASE(16)
    default: set_error("tpat_layernorm: unsupported D=%d", D); return 1;
  }
#undef TPAT_LN_CASE
  TPAT_LAUNCH_CHECK();
  return 0;
}

static int launch_layernorm_split3(const float* x, const float* g, const float* b, __nv_bfloat16* y, int rows, int D, float eps,
                                   cudaStream_t st) {
  const int grid = (rows + LN_WARPS - 1) / LN_WARPS;
#define TPAT_LN_CASE(nv) \
  case nv: TPAT_CUDA(launch_kernel(layernorm_split3_kernel<nv>, dim3(grid), dim3(32 * LN_WARPS), 0, st, x, g, b, y, rows, eps, g_walk_desc)); break;
  switch (D / 128) {
    TPAT_LN_CASE(3) TPAT_LN_CASE(6) TPAT_LN_CASE(8)
    default: set_error("tpat_layernorm(split3): unsupported D=%d", D); return 1;
  }
#undef TPAT_LN_CASE
  TPAT_LAUNCH_CHECK();
  return 0;
}

template <typename OutT>
static int launch_gather_ln(const float* x, const int64_t* idx, float* xo, const float* g, const float* b, OutT* yo,
                            int B, int N_in, int k, int extra, int out_rows, int D, float eps, cudaStream_t st) {
  const int rows = B * (extra + k);
  const int grid = (rows + LN_WARPS - 1) / LN_WARPS;
#define TPAT_GLN_CASE(nv) \
  case nv: TPAT_CUDA(launch_kernel(gather_layernorm_kernel<nv, OutT>, dim3(grid), dim3(32 * LN_WARPS), 0, st, x, idx, xo, g, b, yo, B, N_in, k, extra, out_rows, eps, g_walk_desc)); break;
  switch (D / 128) {
    TPAT_GLN_CASE(1) TPAT_GLN_CASE(2) TPAT_GLN_CASE(3) TPAT_GLN_CASE(4) TPAT_GLN_CASE(5) TPAT_GLN_CASE(6)
    TPAT_GLN_CASE(8) TPAT_GLN_CASE(10) TPAT_GLN_CASE(12) TPAT_GLN_CASE(16)
    default: set_error("tpat_gather_layernorm: unsupported D=%d", D); return 1;
  }
#undef TPAT_GLN_CASE
  TPAT_LAUNCH_CHECK();
  return 0;
}

}  // namespace tpat

extern "C" int tpat_layernorm(const float* x, const float* gamma, const float* beta, void* y, int y_dtype,
                              int rows, int D, float eps, tpat_stream_t stream) {
  using namespace tpat;
  TPAT_CHECK(x && gamma && beta && y, "tpat_layernorm: null pointer");
  TPAT_CHECK(rows >= 0 && D > 0 && D % 128 == 0 && D <= 2048, "tpat_layernorm: need D %% 128 == 0, D <= 2048 (D=%d)", D);
  TPAT_CHECK(aligned16(x) && aligned16(y) && aligned16(gamma) && aligned16(beta), "tpat_layernorm: pointers must be 16-byte aligned");
  if (rows == 0) return 0;
  if (y_dtype == TPAT_F32) return launch_layernorm<float>(x, gamma, beta, (float*)y, rows, D, eps, as_stream(stream));
  if (y_dtype == TPAT_BF16) return launch_layernorm<__nv_bfloat16>(x, gamma, beta, (__nv_bfloat16*)y, rows, D, eps, as_stream(stream));
  if (y_dtype == TPAT_BF16_SPLIT3) return launch_layernorm_split3(x, gamma, beta, (__nv_bfloat16*)y, rows, D, eps, as_stream(stream));
  set_error("tpat_layernorm: bad dtype %d", y_dtype);
  return 1;
}

extern "C" int tpat_split_bf16(const float* x, void* out, int rows, int cols, tpat_stream_t stream) {
  using namespace tpat;
  TPAT_CHECK(x && out, "tpat_split_bf16: null pointer");
  TPAT_CHECK(rows >= 0 && cols > 0 && cols % 4 == 0, "tpat_split_bf16: cols must be a positive multiple of 4 (cols=%d)", cols);
  TPAT_CHECK(aligned16(x) && aligned16(out), "tpat_split_bf16: pointers must be 16-byte aligned");
  if (rows == 0) return 0;
  const size_t total = (size_t)rows * (cols / 4);
  const int grid = (int)((total + 255) / 256 < (size_t)sm_count() * 16 ? (total + 255) / 256 : (size_t)sm_count() * 16);
  TPAT_CUDA(launch_kernel(split_bf16_kernel, dim3(grid), dim3(256), 0, as_stream(stream), x, (__nv_bfloat16*)out, rows, cols));
  TPAT_LAUNCH_CHECK();
  return 0;
}

extern "C" int tpat_gather_layernorm(const float* x, const int64_t* topk_idx, float* x_out, const float* gamma,
                                     const float* beta, void* y_out, int y_dtype, int B, int N_in, int k,
                                     int num_extra, int out_rows, int D, float eps, tpat_stream_t stream) {
  using namespace tpat;
  TPAT_CHECK(x && topk_idx && x_out, "tpat_gather_layernorm: null pointer");
  TPAT_CHECK(y_out == nullptr || (gamma && beta), "tpat_gather_layernorm: y_out needs gamma and beta");
  TPAT_CHECK(B >= 0 && k > 0 && num_extra >= 0 && num_extra + k <= N_in, "tpat_gather_layernorm: bad sizes N_in=%d k=%d extra=%d", N_in, k, num_extra);
  TPAT_CHECK(D > 0 && D % 128 == 0 && D <= 2048, "tpat_gather_layernorm: unsupported D=%d", D);
  TPAT_CHECK(x != x_out, "tpat_gather_layernorm: in-place compaction is not supported");
  TPAT_CHECK(aligned16(x) && aligned16(x_out) && (y_out == nullptr || aligned16(y_out)), "tpat_gather_layernorm: pointers must be 16-byte aligned");
  if (out_rows == 0) out_rows = num_extra + k;
  TPAT_CHECK(out_rows >= num_extra + k, "tpat_gather_layernorm: out_rows=%d smaller than extra + k = %d", out_rows, num_extra + k);
  if (B == 0) return 0;
  if (y_dtype == TPAT_F32) return launch_gather_ln<float>(x, topk_idx, x_out, gamma, beta, (float*)y_out, B, N_in, k, num_extra, out_rows, D, eps, as_stream(stream));
  if (y_dtype == TPAT_BF16) return launch_gather_ln<__nv_bfloat16>(x, topk_idx, x_out, gamma, beta, (__nv_bfloat16*)y_out, B, N_in, k, num_extra, out_rows, D, eps, as_stream(stream));
  set_error("tpat_gather_layernorm: bad dtype %d", y_dtype);
  return 1;
}

extern "C" int tpat_pool_norm(const float* x, float* pooled, const float* g1, const float* b1, float eps1,
                              const float* g2, const float* b2, float eps2, int B, int N, int D, int variant,
                              tpat_stream_t stream) {
  using namespace tpat;
  TPAT_CHECK(x && pooled && g1 && b1, "tpat_pool_norm: null pointer");
  TPAT_CHECK(variant == TPAT_VARIANT_AUDIOMAE || variant == TPAT_VARIANT_AST, "tpat_pool_norm: bad variant %d", variant);
  TPAT_CHECK(variant == TPAT_VARIANT_AUDIOMAE || (g2 && b2), "tpat_pool_norm: AST needs the mlp_head LayerNorm parameters");
  TPAT_CHECK(B >= 0 && N >= 2 && D > 0 && D <= 2048, "tpat_pool_norm: bad sizes N=%d D=%d", N, D);
  if (B == 0) return 0;
  TPAT_CHECK(D % 4 == 0 && D <= 2048, "tpat_pool_norm: D must be a multiple of 4 and <= 2048");
  const int threads = 1024;
  const int TG = threads / (D / 4);
  const size_t smem = ((size_t)TG * D + 2 * D + 32) * sizeof(float);
  TPAT_CUDA(launch_kernel(pool_norm_kernel, dim3(B), dim3(threads), smem, as_stream(stream), x, pooled, g1, b1, eps1, g2, b2, eps2, N, D, variant));
  TPAT_LAUNCH_CHECK();
  return 0;
}

extern "C" int tpat_head(const float* pooled, const float* W, const float* bias, float* logits, int B, int D, int C,
                         tpat_stream_t stream) {
  using namespace tpat;
  TPAT_CHECK(pooled && W && logits, "tpat_head: null pointer");
  TPAT_CHECK(B >= 0 && C > 0 && D > 0 && D % 128 == 0 && D <= 1024, "tpat_head: need D %% 128 == 0, D <= 1024 (D=%d)", D);
  TPAT_CHECK(aligned16(pooled) && aligned16(W), "tpat_head: pooled and W must be 16-byte aligned");
  if (B == 0) return 0;
  const dim3 grid((C + 7) / 8, B >= 32 ? 8 : (B >= 8 ? 2 : 1));
  cudaStream_t st = as_stream(stream);
  switch (D / 128) {
    case 1: TPAT_CUDA(launch_kernel(head_kernel<1>, dim3(grid), dim3(256), 0, st, pooled, W, bias, logits, B, C)); break;
    case 2: TPAT_CUDA(launch_kernel(head_kernel<2>, dim3(grid), dim3(256), 0, st, pooled, W, bias, logits, B, C)); break;
    case 3: TPAT_CUDA(launch_kernel(head_kernel<3>, dim3(grid), dim3(256), 0, st, pooled, W, bias, logits, B, C)); break;
    case 4: TPAT_CUDA(launch_kernel(head_kernel<4>, dim3(grid), dim3(256), 0, st, pooled, W, bias, logits, B, C)); break;
    case 6: TPAT_CUDA(launch_kernel(head_kernel<6>, dim3(grid), dim3(256), 0, st, pooled, W, bias, logits, B, C)); break;
    case 8: TPAT_CUDA(launch_kernel(head_kernel<8>, dim3(grid), dim3(256), 0, st, pooled, W, bias, logits, B, C)); break;
    default: set_error("tpat_head: unsupported D=%d", D); return 1;
  }
  TPAT_LAUNCH_CHECK();
  return 0;
}
