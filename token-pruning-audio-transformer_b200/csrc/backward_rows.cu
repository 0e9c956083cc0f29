// Row-wise (HBM-bound) kernels of the fine-tune backward pass.
//
// The reference's backward is PyTorch autograd over audiomae/models_vit.py:191-207 (Block.forward): LayerNorm backward
// (:197,205), the scatter that is the backward of `torch.gather` + `torch.cat` (:200-203; no gradient flows through
// the score / top-k, they are indices), DropPath's per-sample scale (:149,198,205), bias gradients (column sums) and
// the pooled head (:388-389,522 / ast_models.py:500-503).  Here those are three kernels:
//
//   row_bwd_kernel      one warp per OUTPUT row of the gradient stream: optional LayerNorm backward of the row it maps
//                       to (statistics recomputed from the saved LayerNorm input: the row is read anyway), + the
//                       residual-path gradient, scattered back to the pre-gather row order (rows that were pruned get
//                       zeros), written twice: fp32 (the stream) and `scale[b] * g` in the GEMM operand dtype (the
//                       dY of the Linear that produced the branch).  Per-CTA partial column sums of dgamma, dbeta and
//                       of the operand copy (= that Linear's bias gradient) go to a partials buffer.
//   colsum_kernel       partial column sums of a [M, C] matrix (bias gradients of qkv / fc1).
//   partials_finish     fixed-order sum of the per-CTA partials into the gradient buffers (+=): deterministic.
//   pool_norm_bwd       backward of tpat_pool_norm, one CTA per clip.
// Bytes per row (D = 768, bf16 operands): 3 KB (dy fp32) + 3 KB (x) + 3 KB (g_up) read, 3 KB + 1.5 KB written.
#include "common.cuh"
#include "gemm.cuh"

namespace tpat {

constexpr int BR_WARPS = 8;

__global__ void __launch_bounds__(256)
inverse_index_kernel(const int64_t* __restrict__ idx, int* __restrict__ inv, int n, int k) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < n; i += blockDim.x) inv[(size_t)b * n + i] = -1;
  __syncthreads();
  for (int j = threadIdx.x; j < k; j += blockDim.x) inv[(size_t)b * n + (int)idx[(size_t)b * k + j]] = j;
}

struct RowBwdParams {
  const void* dy;          // [B * rows_src, D]  gradient w.r.t. the LayerNorm output (HAS_LN)
  const float* x;          // [B * rows_src, D]  the LayerNorm input (HAS_LN)
  const float* gamma;      // [D] (HAS_LN)
  const float* g_up;       // [B * rows_src, D]  gradient on the residual path, or NULL
  float* g_out;            // [B * rows_out, D]  fp32, or NULL
  void* gb_out;            // [B * rows_out, D]  operand dtype: row_scale[b] * g, or NULL
  const float* row_scale;  // [B] or NULL (DropPath)
  const int* inv;          // [B, rows_out - extra]: position among the kept rows or -1; NULL = no scatter
  float* partials;         // [gridDim.x][3][D]: dgamma, dbeta, colsum(gb)
  int B, rows_src, rows_out, extra, src_offset;
  float eps;
};

template <typename T> __device__ __forceinline__ float4 ld_row4(const T* row, int c4);
template <> __device__ __forceinline__ float4 ld_row4<float>(const float* row, int c4) {
  return __ldg(reinterpret_cast<const float4*>(row) + c4);
}
template <> __device__ __forceinline__ float4 ld_row4<__nv_bfloat16>(const __nv_bfloat16* row, int c4) {
  const uint2 v = __ldg(reinterpret_cast<const uint2*>(row) + c4);
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&v.x), b = *reinterpret_cast<const __nv_bfloat162*>(&v.y);
  return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
}
template <typename T> __device__ __forceinline__ void st_row4(T* row, int c4, float4 v);
template <> __device__ __forceinline__ void st_row4<float>(float* row, int c4, float4 v) { reinterpret_cast<float4*>(row)[c4] = v; }
template <> __device__ __forceinline__ void st_row4<__nv_bfloat16>(__nv_bfloat16* row, int c4, float4 v) {
  reinterpret_cast<uint2*>(row)[c4] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
}

template <int NV, typename DyT, typename OutT, bool HAS_LN>
__global__ void __launch_bounds__(32 * BR_WARPS)
row_bwd_kernel(const RowBwdParams p) {
  constexpr int D = NV * 128;
  extern __shared__ float br_sm[];     // [BR_WARPS][D]
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 a_dg[NV], a_db[NV], a_cs[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) a_dg[i] = a_db[i] = a_cs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int total = p.B * p.rows_out;
  for (int o = blockIdx.x * BR_WARPS + warp; o < total; o += gridDim.x * BR_WARPS) {
    const int b = o / p.rows_out, j = o - b * p.rows_out;
    int src = j + p.src_offset;
    if (p.inv != nullptr && j >= p.extra) {
      const int t = __ldg(p.inv + (size_t)b * (p.rows_out - p.extra) + (j - p.extra));
      src = t < 0 ? -1 : p.extra + t;
    }
    float* go = p.g_out ? p.g_out + (size_t)o * D : nullptr;
    OutT* gbo = p.gb_out ? reinterpret_cast<OutT*>(p.gb_out) + (size_t)o * D : nullptr;
    if (src < 0) {                       // this token was pruned after the branch point: no gradient reaches it
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if (go) st_row4<float>(go, lane + 32 * i, make_float4(0.f, 0.f, 0.f, 0.f));
        if (gbo) st_row4<OutT>(gbo, lane + 32 * i, make_float4(0.f, 0.f, 0.f, 0.f));
      }
      continue;
    }
    const size_t srow = (size_t)b * p.rows_src + src;
    float4 g[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) g[i] = p.g_up ? __ldg(reinterpret_cast<const float4*>(p.g_up + srow * D) + lane + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
    if constexpr (HAS_LN) {
      float4 xv[NV], dyv[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        xv[i] = __ldg(reinterpret_cast<const float4*>(p.x + srow * D) + lane + 32 * i);
        dyv[i] = ld_row4<DyT>(reinterpret_cast<const DyT*>(p.dy) + srow * D, lane + 32 * i);
      }
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) s += (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w);
      const float mean = warp_sum(s) * (1.0f / D);
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        xv[i].x -= mean; xv[i].y -= mean; xv[i].z -= mean; xv[i].w -= mean;
        q += (xv[i].x * xv[i].x + xv[i].y * xv[i].y) + (xv[i].z * xv[i].z + xv[i].w * xv[i].w);
      }
      const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / D) + p.eps);
      float c1 = 0.f, c2 = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float4 gm = __ldg(reinterpret_cast<const float4*>(p.gamma) + lane + 32 * i);
        xv[i].x *= rstd; xv[i].y *= rstd; xv[i].z *= rstd; xv[i].w *= rstd;          // xhat
        a_dg[i].x = fmaf(dyv[i].x, xv[i].x, a_dg[i].x); a_dg[i].y = fmaf(dyv[i].y, xv[i].y, a_dg[i].y);
        a_dg[i].z = fmaf(dyv[i].z, xv[i].z, a_dg[i].z); a_dg[i].w = fmaf(dyv[i].w, xv[i].w, a_dg[i].w);
        a_db[i].x += dyv[i].x; a_db[i].y += dyv[i].y; a_db[i].z += dyv[i].z; a_db[i].w += dyv[i].w;
        dyv[i].x *= gm.x; dyv[i].y *= gm.y; dyv[i].z *= gm.z; dyv[i].w *= gm.w;      // dy * gamma
        c1 += (dyv[i].x + dyv[i].y) + (dyv[i].z + dyv[i].w);
        c2 += (dyv[i].x * xv[i].x + dyv[i].y * xv[i].y) + (dyv[i].z * xv[i].z + dyv[i].w * xv[i].w);
      }
      c1 = warp_sum(c1) * (1.0f / D);
      c2 = warp_sum(c2) * (1.0f / D);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        g[i].x += rstd * (dyv[i].x - c1 - xv[i].x * c2); g[i].y += rstd * (dyv[i].y - c1 - xv[i].y * c2);
        g[i].z += rstd * (dyv[i].z - c1 - xv[i].z * c2); g[i].w += rstd * (dyv[i].w - c1 - xv[i].w * c2);
      }
    }
    const float sc = p.row_scale ? __ldg(p.row_scale + b) : 1.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (go) st_row4<float>(go, lane + 32 * i, g[i]);
      const float4 gs = make_float4(g[i].x * sc, g[i].y * sc, g[i].z * sc, g[i].w * sc);
      if (gbo) st_row4<OutT>(gbo, lane + 32 * i, gs);
      a_cs[i].x += gs.x; a_cs[i].y += gs.y; a_cs[i].z += gs.z; a_cs[i].w += gs.w;
    }
  }
  if (p.partials == nullptr) return;
  // per-CTA partials: the 8 warps' accumulators are summed in a fixed order through shared memory
  float* dst = p.partials + (size_t)blockIdx.x * 3 * D;
#pragma unroll 1
  for (int qn = 0; qn < 3; ++qn) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i)
      reinterpret_cast<float4*>(br_sm + warp * D)[lane + 32 * i] = qn == 0 ? a_dg[i] : (qn == 1 ? a_db[i] : a_cs[i]);
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < BR_WARPS; ++w) s += br_sm[w * D + c];
      dst[qn * D + c] = s;
    }
  }
}

// partial column sums of x [M, C] (ld elements between rows): grid = (column blocks of 512, row slabs), partials
// [slab][C].  128 column groups of 4 x 2 row phases per CTA, eight rows in flight per thread.
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ x, int ld, int M, int C, float* __restrict__ partials) {
  __shared__ float4 red[128];
  pdl_trigger();
  pdl_wait();
  const int cgp = threadIdx.x & 127, ph = threadIdx.x >> 7;
  const int c4 = blockIdx.x * 128 + cgp;
  const int per = (M + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * per, r1 = min(M, r0 + per);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c4 < C / 4) {
    int r = r0 + ph;
    for (; r + 14 < r1; r += 16) {
      float4 u[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) u[k] = ld_row4<T>(x + (size_t)(r + 2 * k) * ld, c4);
#pragma unroll
      for (int k = 0; k < 8; ++k) { acc.x += u[k].x; acc.y += u[k].y; acc.z += u[k].z; acc.w += u[k].w; }
    }
    for (; r < r1; r += 2) { const float4 u = ld_row4<T>(x + (size_t)r * ld, c4); acc.x += u.x; acc.y += u.y; acc.z += u.z; acc.w += u.w; }
  }
  if (ph == 1) red[cgp] = acc;
  __syncthreads();
  if (ph == 0 && c4 < C / 4) {
    const float4 o = red[cgp];
    reinterpret_cast<float4*>(partials + (size_t)blockIdx.y * C)[c4] = make_float4(acc.x + o.x, acc.y + o.y, acc.z + o.z, acc.w + o.w);
  }
}

// the same for bf16 rows with C % 8 == 0: 16-byte loads (a warp instruction covers 512 contiguous bytes, as in the
// LayerNorm kernels), blockDim.x / 2 column groups of 8 x 2 row phases, eight rows in flight per thread
__global__ void __launch_bounds__(256)
colsum8_kernel(const __nv_bfloat16* __restrict__ x, int ld, int M, int C, float* __restrict__ partials) {
  __shared__ float red[128][9];
  pdl_trigger();
  pdl_wait();
  const int gpb = blockDim.x >> 1;
  const int cgp = threadIdx.x % gpb, ph = threadIdx.x / gpb;
  const int c8 = blockIdx.x * gpb + cgp;
  const int per = (M + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * per, r1 = min(M, r0 + per);
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  auto add = [&](const uint4& v) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {               // bf16 -> fp32 is a 16-bit shift
      acc[2 * e] += __uint_as_float(w[e] << 16);
      acc[2 * e + 1] += __uint_as_float(w[e] & 0xffff0000u);
    }
  };
  if (c8 < C / 8) {
    const uint4* base = reinterpret_cast<const uint4*>(x) + c8;
    const size_t ld16 = (size_t)ld / 8;
    int r = r0 + ph;
    for (; r + 14 < r1; r += 16) {
      uint4 u[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) u[k] = __ldg(base + (size_t)(r + 2 * k) * ld16);
#pragma unroll
      for (int k = 0; k < 8; ++k) add(u[k]);
    }
    for (; r < r1; r += 2) add(__ldg(base + (size_t)r * ld16));
  }
  if (ph == 1) {
#pragma unroll
    for (int e = 0; e < 8; ++e) red[cgp][e] = acc[e];
  }
  __syncthreads();
  if (ph == 0 && c8 < C / 8) {
    float4* dst = reinterpret_cast<float4*>(partials + (size_t)blockIdx.y * C) + 2 * c8;
    dst[0] = make_float4(acc[0] + red[cgp][0], acc[1] + red[cgp][1], acc[2] + red[cgp][2], acc[3] + red[cgp][3]);
    dst[1] = make_float4(acc[4] + red[cgp][4], acc[5] + red[cgp][5], acc[6] + red[cgp][6], acc[7] + red[cgp][7]);
  }
}

// dst_q[c] += sum_p partials[p][q][c], q < nq (<= 4), in a fixed order.  CTA = 32 columns x 16 part-slices: slice s sums
// parts s, s + 16, ... (four independent accumulators), the 16 slice sums are combined through shared memory in order --
// the loop over up to 296 parts is 5 dependent rounds instead of 74.
struct FinishDst { float* d[4]; };
__global__ void __launch_bounds__(512)
partials_finish_kernel(const float* __restrict__ partials, int nparts, int nq, int C, FinishDst dst) {
  __shared__ float red[16][33];
  pdl_trigger();
  pdl_wait();
  const int cx = threadIdx.x & 31, sy = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const int q = blockIdx.y;
  if (dst.d[q] == nullptr) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (c < C) {
    int pi = sy;
    for (; pi + 48 < nparts; pi += 64) {
      s0 += partials[((size_t)pi * nq + q) * C + c];
      s1 += partials[((size_t)(pi + 16) * nq + q) * C + c];
      s2 += partials[((size_t)(pi + 32) * nq + q) * C + c];
      s3 += partials[((size_t)(pi + 48) * nq + q) * C + c];
    }
    for (; pi < nparts; pi += 16) s0 += partials[((size_t)pi * nq + q) * C + c];
  }
  red[sy][cx] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (sy == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) t += red[k][cx];
    dst.d[q][c] += t;
  }
}

// ---- pooled head backward: one CTA per clip, 256 threads ----
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = (threadIdx.x & 31) < (int)(blockDim.x >> 5) ? red[threadIdx.x & 31] : 0.f;
  return warp_sum(t);
}

// LayerNorm backward of one D-vector in shared memory: given v (input, overwritten by xhat), dy (overwritten by dx)
__device__ void block_ln_bwd(float* v, float* dy, int D, const float* __restrict__ gamma, float eps, float* red,
                             float* dgamma_acc, float* dbeta_acc) {
  float s = 0.f;
  for (int c = threadIdx.x; c < D; c += blockDim.x) s += v[c];
  const float mean = block_sum(s, red) / D;
  float q = 0.f;
  for (int c = threadIdx.x; c < D; c += blockDim.x) { const float d = v[c] - mean; q += d * d; }
  const float rstd = 1.0f / sqrtf(block_sum(q, red) / D + eps);
  float c1 = 0.f, c2 = 0.f;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    const float xh = (v[c] - mean) * rstd;
    v[c] = xh;
    dgamma_acc[c] += dy[c] * xh;
    dbeta_acc[c] += dy[c];
    const float dg = dy[c] * gamma[c];
    dy[c] = dg;
    c1 += dg; c2 += dg * xh;
  }
  c1 = block_sum(c1, red) / D;
  c2 = block_sum(c2, red) / D;
  for (int c = threadIdx.x; c < D; c += blockDim.x) dy[c] = rstd * (dy[c] - c1 - v[c] * c2);
  __syncthreads();
}

// out[c] = LayerNorm(v)[c] * gamma + beta for a D-vector in shared memory (v is left untouched)
__device__ void block_ln_fwd(const float* v, float* out, int D, const float* __restrict__ gamma, const float* __restrict__ beta,
                             float eps, float* red) {
  float s = 0.f;
  for (int c = threadIdx.x; c < D; c += blockDim.x) s += v[c];
  const float mean = block_sum(s, red) / D;
  float q = 0.f;
  for (int c = threadIdx.x; c < D; c += blockDim.x) { const float d = v[c] - mean; q += d * d; }
  const float rstd = 1.0f / sqrtf(block_sum(q, red) / D + eps);
  for (int c = threadIdx.x; c < D; c += blockDim.x) out[c] = (v[c] - mean) * rstd * gamma[c] + beta[c];
  __syncthreads();
}

// x [B, N, D] (input of tpat_pool_norm), dpooled [B, D] -> dx [B, N, D]; partials [B][4][D] = (dg1, db1, dg2, db2)
//   AudioMAE: pooled = LN1(mean_{t >= 1} x_t)                          (models_vit.py:388-389)
//   AST:      pooled = LN2((LN1(x_0) + LN1(x_1)) / 2), LN1 = v.norm, LN2 = mlp_head.0   (ast_models.py:500-503)
__global__ void __launch_bounds__(256)
pool_norm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dpooled, float* __restrict__ dx,
                     const float* __restrict__ g1, const float* __restrict__ b1, float eps1,
                     const float* __restrict__ g2, float eps2, float* __restrict__ partials, int N, int D, int variant) {
  extern __shared__ float pb_sm[];     // [v0 | v1 | d0 | d1 | acc 4*D | red 32]
  pdl_trigger();
  pdl_wait();
  float* v0 = pb_sm; float* v1 = v0 + D; float* d0 = v1 + D; float* d1 = d0 + D; float* acc = d1 + D; float* red = acc + 4 * D;
  const float* xb = x + (size_t)blockIdx.x * N * D;
  float* dxb = dx + (size_t)blockIdx.x * N * D;
  for (int c = threadIdx.x; c < 4 * D; c += blockDim.x) acc[c] = 0.f;
  for (int c = threadIdx.x; c < D; c += blockDim.x) d0[c] = dpooled[(size_t)blockIdx.x * D + c];
  if (variant == TPAT_VARIANT_AUDIOMAE) {
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      float s0 = 0.f, s1 = 0.f;
      int t = 1;
      for (; t + 1 < N; t += 2) { s0 += xb[(size_t)t * D + c]; s1 += xb[(size_t)(t + 1) * D + c]; }
      if (t < N) s0 += xb[(size_t)t * D + c];
      v0[c] = (s0 + s1) / (float)(N - 1);
    }
    __syncthreads();
    block_ln_bwd(v0, d0, D, g1, eps1, red, acc, acc + D);
    const float inv = 1.0f / (float)(N - 1);
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      const float g = d0[c] * inv;
      dxb[c] = 0.f;
      for (int t = 1; t < N; ++t) dxb[(size_t)t * D + c] = g;
    }
  } else {
    for (int c = threadIdx.x; c < D; c += blockDim.x) { v0[c] = xb[c]; v1[c] = xb[D + c]; }
    __syncthreads();
    float* p = d1;                       // p = (LN1(x0) + LN1(x1)) / 2, built in two steps
    block_ln_fwd(v0, p, D, g1, b1, eps1, red);
    float* t1 = acc + 2 * D;             // scratch: the dg2 slot is still zero and unused until the LN2 backward
    block_ln_fwd(v1, t1, D, g1, b1, eps1, red);
    for (int c = threadIdx.x; c < D; c += blockDim.x) { p[c] = 0.5f * (p[c] + t1[c]); t1[c] = 0.f; }
    __syncthreads();
    block_ln_bwd(p, d0, D, g2, eps2, red, acc + 2 * D, acc + 3 * D);          // d0 = dL/dp
    for (int c = threadIdx.x; c < D; c += blockDim.x) { const float h = 0.5f * d0[c]; d0[c] = h; d1[c] = h; }
    __syncthreads();
    block_ln_bwd(v0, d0, D, g1, eps1, red, acc, acc + D);                     // d0 = dL/dx0
    block_ln_bwd(v1, d1, D, g1, eps1, red, acc, acc + D);                     // d1 = dL/dx1
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      dxb[c] = d0[c];
      dxb[D + c] = d1[c];
      for (int t = 2; t < N; ++t) dxb[(size_t)t * D + c] = 0.f;
    }
  }
  __syncthreads();
  float* dst = partials + (size_t)blockIdx.x * 4 * D;
  for (int c = threadIdx.x; c < 4 * D; c += blockDim.x) dst[c] = acc[c];
}

}  // namespace tpat

namespace tpat {

static int finish_partials(const float* partials, int nparts, int nq, int C, float* d0, float* d1, float* d2, float* d3, cudaStream_t st) {
  FinishDst dst{{d0, d1, d2, d3}};
  TPAT_CUDA(launch_kernel(partials_finish_kernel, dim3((C + 31) / 32, nq), dim3(512), 0, st, partials, nparts, nq, C, dst));
  TPAT_LAUNCH_CHECK();
  return 0;
}

int finish_colsum_partials(const float* partials, int nparts, int C, float* dst, cudaStream_t st) {
  return finish_partials(partials, nparts, 1, C, dst, nullptr, nullptr, nullptr, st);
}

template <typename DyT, typename OutT, bool HAS_LN>
static int launch_row_bwd(const RowBwdParams& p, int D, int grid, cudaStream_t st) {
  const size_t smem = (size_t)BR_WARPS * D * sizeof(float);
#define TPAT_RB_CASE(nv) \
  case nv: TPAT_CUDA(launch_kernel(row_bwd_kernel<nv, DyT, OutT, HAS_LN>, dim3(grid), dim3(32 * BR_WARPS), smem, st, p)); break;
  switch (D / 128) {
    TPAT_RB_CASE(3) TPAT_RB_CASE(6) TPAT_RB_CASE(8)
    default: set_error("tpat_row_bwd: unsupported D=%d (384, 768, 1024)", D); return 1;
  }
#undef TPAT_RB_CASE
  TPAT_LAUNCH_CHECK();
  return 0;
}

}  // namespace tpat

extern "C" size_t tpat_bwd_partials_floats(int D_max) { return (size_t)2 * 148 * 4 * (size_t)(D_max > 4096 ? D_max : 4096); }

extern "C" int tpat_inverse_index(const int64_t* idx, int32_t* inv, int B, int n, int k, tpat_stream_t stream) {
  using namespace tpat;
  TPAT_CHECK(idx && inv && B >= 0 && n > 0 && k > 0 && k <= n, "tpat_inverse_index: bad arguments (n=%d k=%d)", n, k);
  if (B == 0) return 0;
  TPAT_CUDA(launch_kernel(inverse_index_kernel, dim3(B), dim3(256), 0, as_stream(stream), idx, inv, n, k));
  TPAT_LAUNCH_CHECK();
  return 0;
}

extern "C" int tpat_row_bwd(const void* dy, int dy_dtype, const float* x, const float* gamma, const float* g_up, float* g_out,
                            void* gb_out, int gb_dtype, const float* row_scale, const int32_t* inv, float* partials_ws,
                            float* dgamma, float* dbeta, float* dbias, int B, int rows_src, int rows_out, int num_extra,
                            int src_offset, int D, float eps, tpat_stream_t stream) {
  using namespace tpat;
  const bool has_ln = dy != nullptr;
  TPAT_CHECK(!has_ln || (x && gamma), "tpat_row_bwd: the LayerNorm backward needs x and gamma");
  TPAT_CHECK(has_ln || g_up, "tpat_row_bwd: nothing to do (no dy, no g_up)");
  TPAT_CHECK(B >= 0 && rows_src > 0 && rows_out > 0 && num_extra >= 0 && src_offset >= 0, "tpat_row_bwd: bad sizes");
  TPAT_CHECK(inv == nullptr || src_offset == 0, "tpat_row_bwd: scatter and row offset are exclusive");
  TPAT_CHECK(inv != nullptr || rows_out + src_offset <= rows_src, "tpat_row_bwd: rows_out + src_offset exceeds rows_src");
  TPAT_CHECK(dy_dtype == TPAT_F32 || dy_dtype == TPAT_BF16, "tpat_row_bwd: bad dy dtype");
  TPAT_CHECK(gb_dtype == TPAT_F32 || gb_dtype == TPAT_BF16, "tpat_row_bwd: bad operand dtype");
  TPAT_CHECK((dgamma == nullptr && dbeta == nullptr && dbias == nullptr) || partials_ws, "tpat_row_bwd: column sums need partials_ws");
  if (B == 0) return 0;
  RowBwdParams p{dy, x, gamma, g_up, g_out, gb_out, row_scale, inv, nullptr, B, rows_src, rows_out, num_extra, src_offset, eps};
  const int total = B * rows_out;
  int grid = (total + BR_WARPS - 1) / BR_WARPS;
  if (grid > 2 * sm_count()) grid = 2 * sm_count();
  const bool sums = dgamma || dbeta || dbias;
  p.partials = sums ? partials_ws : nullptr;
  cudaStream_t st = as_stream(stream);
  int rc;
  if (has_ln) {
    if (dy_dtype == TPAT_F32) rc = gb_dtype == TPAT_F32 ? launch_row_bwd<float, float, true>(p, D, grid, st) : launch_row_bwd<float, __nv_bfloat16, true>(p, D, grid, st);
    else rc = gb_dtype == TPAT_F32 ? launch_row_bwd<__nv_bfloat16, float, true>(p, D, grid, st) : launch_row_bwd<__nv_bfloat16, __nv_bfloat16, true>(p, D, grid, st);
  } else {
    rc = gb_dtype == TPAT_F32 ? launch_row_bwd<float, float, false>(p, D, grid, st) : launch_row_bwd<float, __nv_bfloat16, false>(p, D, grid, st);
  }
  if (rc) return rc;
  if (sums) return finish_partials(partials_ws, grid, 3, D, has_ln ? dgamma : nullptr, has_ln ? dbeta : nullptr, dbias, nullptr, st);
  return 0;
}

extern "C" int tpat_colsum(const void* x, int dtype, int ld, int M, int C, float* partials_ws, float* dst, tpat_stream_t stream) {
  using namespace tpat;
  TPAT_CHECK(x && partials_ws && dst, "tpat_colsum: null pointer");
  TPAT_CHECK(M >= 0 && C > 0 && C % 4 == 0 && C <= 4096 && ld >= C && aligned16(x) && (ld * dtype_size(dtype)) % 8 == 0, "tpat_colsum: need C %% 4 == 0, C <= 4096, aligned rows");
  if (M == 0) return 0;
  const bool wide = dtype != TPAT_F32 && C % 8 == 0 && ld % 8 == 0;
  const int groups = wide ? C / 8 : C / 4;
  const int cblocks = (groups + 127) / 128;
  const int gpb = wide ? ((groups + cblocks - 1) / cblocks + 31) / 32 * 32 : 128;   // balanced column blocks (C = 2304: 3 x 96)
  int grid = (M + 63) / 64;                                 // row slabs
  // up to ~8 CTAs per SM in flight (this kernel is pure streaming: occupancy = bytes in flight); the partials buffer holds
  // tpat_bwd_partials_floats / C >= 1184 * 4096 / C slabs
  int max_slabs = (8 * sm_count()) / cblocks > 0 ? (8 * sm_count()) / cblocks : 1;
  if ((size_t)max_slabs * C > (size_t)2 * 148 * 4 * 4096) max_slabs = (int)((size_t)2 * 148 * 4 * 4096 / C);
  if (grid > max_slabs) grid = max_slabs;
  cudaStream_t st = as_stream(stream);
  if (dtype == TPAT_F32) TPAT_CUDA(launch_kernel(colsum_kernel<float>, dim3(cblocks, grid), dim3(256), 0, st, (const float*)x, ld, M, C, partials_ws));
  else if (wide) TPAT_CUDA(launch_kernel(colsum8_kernel, dim3(cblocks, grid), dim3(2 * gpb), 0, st, (const __nv_bfloat16*)x, ld, M, C, partials_ws));
  else TPAT_CUDA(launch_kernel(colsum_kernel<__nv_bfloat16>, dim3(cblocks, grid), dim3(256), 0, st, (const __nv_bfloat16*)x, ld, M, C, partials_ws));
  TPAT_LAUNCH_CHECK();
  return finish_partials(partials_ws, grid, 1, C, dst, nullptr, nullptr, nullptr, st);
}

extern "C" int tpat_pool_norm_bwd(const float* x, const float* dpooled, float* dx, const float* g1, const float* b1, float eps1,
                                  const float* g2, float eps2, float* partials_ws, float* dg1, float* db1, float* dg2,
                                  float* db2, int B, int N, int D, int variant, tpat_stream_t stream) {
  using namespace tpat;
  TPAT_CHECK(x && dpooled && dx && g1 && b1 && partials_ws && dg1 && db1, "tpat_pool_norm_bwd: null pointer");
  TPAT_CHECK(variant == TPAT_VARIANT_AUDIOMAE || (variant == TPAT_VARIANT_AST && g2 && dg2 && db2), "tpat_pool_norm_bwd: AST needs the mlp_head LayerNorm");
  TPAT_CHECK(B >= 0 && N >= 2 && D > 0 && D <= 2048 && B <= 2 * 148, "tpat_pool_norm_bwd: bad sizes (B <= 296)");
  if (B == 0) return 0;
  const size_t smem = ((size_t)8 * D + 32) * sizeof(float);
  cudaStream_t st = as_stream(stream);
  static DeviceOnce once;
  if (once.first()) { TPAT_CUDA(cudaFuncSetAttribute(pool_norm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024)); once.mark(); }
  TPAT_CUDA(launch_kernel(pool_norm_bwd_kernel, dim3(B), dim3(256), smem, st, x, dpooled, dx, g1, b1, eps1, g2, eps2, partials_ws, N, D, variant));
  TPAT_LAUNCH_CHECK();
  return finish_partials(partials_ws, B, 4, D, dg1, db1, variant == TPAT_VARIANT_AST ? dg2 : nullptr, variant == TPAT_VARIANT_AST ? db2 : nullptr, st);
}
