// Patch extraction for the 16x16 / stride-16 patch-embed conv + the cls/dist rows.
//
// Replaces the unfold half of PatchEmbed.forward (reference audiomae/models_vit.py:241-247,
// ast/src/models/ast_models.py:36-42) and the extra-token assembly (models_vit.py:359-362,
// ast_models.py:463-466).  HBM-bound: reads the fp32 spectrogram once (coalesced 512 B rows),
// transposes 16 x F tiles through shared memory and writes the [B*P, 256] patch matrix with
// fully coalesced 1 KB / 512 B rows.  Algorithmic bytes per clip (T=1024, F=128):
// 512 KB read + 256 KB (bf16) or 512 KB (fp32) written.
#include "common.cuh"

namespace tpat {

// One CTA per (clip, 16-frame time block).  smem tile [16][F + 1] fp32.
template <typename OutT>
__global__ void __launch_bounds__(256)
patchify_kernel(const float* __restrict__ spec, OutT* __restrict__ patches, float* __restrict__ tokens,
                const float* __restrict__ extra_tok, const float* __restrict__ pos,
                int T, int F, int D, int num_extra, int order) {
  extern __shared__ float tile[];  // [16][F+1]
  pdl_trigger();
  pdl_wait();
  const int tb = blockIdx.x, b = blockIdx.y;
  const int TB = T / 16, FB = F / 16, P = TB * FB;
  const int ld = F + 1;
  const float* src = spec + ((size_t)b * T + (size_t)tb * 16) * F;
  // coalesced load: 16 rows x F floats, float4 per thread
  const int nvec = 16 * F / 4;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    const int r = (v * 4) / F, c = (v * 4) % F;
    const float4 val = __ldg(reinterpret_cast<const float4*>(src + (size_t)r * F + c));
    float* d = tile + r * ld + c;
    d[0] = val.x; d[1] = val.y; d[2] = val.z; d[3] = val.w;
  }
  __syncthreads();
  // FB patches x 256 columns; consecutive threads -> consecutive columns (coalesced stores)
  for (int o = threadIdx.x; o < FB * 256; o += blockDim.x) {
    const int fb = o >> 8, col = o & 255;
    int to, fo;
    size_t row;
    if (order == TPAT_TOKENS_TIME_MAJOR) {  // AudioMAE: conv over [T, F], kernel [t_off][f_off]
      to = col >> 4; fo = col & 15;
      row = (size_t)b * P + (size_t)tb * FB + fb;
    } else {                                // AST: conv over [F, T], kernel [f_off][t_off]
      fo = col >> 4; to = col & 15;
      row = (size_t)b * P + (size_t)fb * TB + tb;
    }
    patches[row * 256 + col] = from_f32<OutT>(tile[to * ld + fb * 16 + fo]);
  }
  // extra-token rows of this clip: cls (+dist) + pos
  if (tb == 0 && tokens != nullptr) {
    float* dst = tokens + (size_t)b * (num_extra + P) * D;
    for (int i = threadIdx.x; i < num_extra * D; i += blockDim.x) dst[i] = extra_tok[i] + pos[i];
  }
}

// ---- ablation ranking vectors ----
// One CTA per (16-frame block, clip): the 16 x F tile goes through shared memory, each warp then reduces whole
// patches (lane = 8 of the 256 samples).  Sums in double: the result is the correctly rounded fp32 statistic up to
// the final rounding (torch reduces in fp32 / Welford; both sit within 1-2 ulp of the exact value).
__global__ void __launch_bounds__(256)
patch_stats_kernel(const float* __restrict__ spec, float* __restrict__ mean_out, float* __restrict__ std_out,
                   int T, int F, int order) {
  extern __shared__ float tile[];  // [16][F+1]
  pdl_trigger();
  pdl_wait();
  const int tb = blockIdx.x, b = blockIdx.y;
  const int TB = T / 16, FB = F / 16, P = TB * FB;
  const int ld = F + 1;
  const float* src = spec + ((size_t)b * T + (size_t)tb * 16) * F;
  for (int v = threadIdx.x; v < 16 * F / 4; v += blockDim.x) {
    const int r = (v * 4) / F, c = (v * 4) % F;
    const float4 val = __ldg(reinterpret_cast<const float4*>(src + (size_t)r * F + c));
    float* d = tile + r * ld + c;
    d[0] = val.x; d[1] = val.y; d[2] = val.z; d[3] = val.w;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int fb = warp; fb < FB; fb += blockDim.x >> 5) {
    float v[8];
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int e = lane * 8 + i;                       // sample e of the patch: frame e / 16, bin e % 16
      v[i] = tile[(e >> 4) * ld + fb * 16 + (e & 15)];
      s += (double)v[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const double mu = s / 256.0;
    double q = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { const double d = (double)v[i] - mu; q += d * d; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (lane == 0) {
      const size_t tok = (size_t)b * P + (order == TPAT_TOKENS_TIME_MAJOR ? (size_t)tb * FB + fb : (size_t)fb * TB + tb);
      if (mean_out) mean_out[tok] = (float)mu;
      if (std_out) std_out[tok] = (float)sqrt(q / 255.0);
    }
  }
}

__global__ void gather_rank_kernel(const float* __restrict__ rank, const int64_t* __restrict__ idx, float* __restrict__ out,
                                   int n, int k, int total) {
  pdl_trigger();
  pdl_wait();   // idx comes from the score / top-k kernel launched just before
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int b = i / k;
  out[i] = __ldg(rank + (size_t)b * n + (int)__ldg(idx + i));
}

}  // namespace tpat

extern "C" int tpat_patch_stats(const float* spec, float* mean, float* std, int B, int T, int F, int order,
                                tpat_stream_t stream) {
  using namespace tpat;
  TPAT_CHECK(spec && (mean || std), "tpat_patch_stats: null pointer");
  TPAT_CHECK(B > 0 && T > 0 && F > 0 && T % 16 == 0 && F % 16 == 0 && F <= 512,
             "tpat_patch_stats: need T %% 16 == 0, F %% 16 == 0, F <= 512 (got B=%d T=%d F=%d)", B, T, F);
  TPAT_CHECK(order == TPAT_TOKENS_TIME_MAJOR || order == TPAT_TOKENS_FREQ_MAJOR, "tpat_patch_stats: bad order %d", order);
  TPAT_CHECK(aligned16(spec), "tpat_patch_stats: spec must be 16-byte aligned");
  const size_t smem = (size_t)16 * (F + 1) * sizeof(float);
  TPAT_CUDA(launch_kernel(patch_stats_kernel, dim3(T / 16, B), dim3(256), smem, as_stream(stream), spec, mean, std, T, F, order));
  TPAT_LAUNCH_CHECK();
  return 0;
}

extern "C" int tpat_gather_rank(const float* rank, const int64_t* idx, float* out, int B, int n, int k,
                                tpat_stream_t stream) {
  using namespace tpat;
  TPAT_CHECK(rank && idx && out, "tpat_gather_rank: null pointer");
  TPAT_CHECK(B >= 0 && n > 0 && k > 0 && k <= n, "tpat_gather_rank: bad sizes n=%d k=%d", n, k);
  if (B == 0) return 0;
  const int total = B * k;
  TPAT_CUDA(launch_kernel(gather_rank_kernel, dim3((total + 255) / 256), dim3(256), 0, as_stream(stream), rank, idx, out, n, k, total));
  TPAT_LAUNCH_CHECK();
  return 0;
}

extern "C" int tpat_patchify(const float* spec, void* patches, int out_dtype, float* tokens,
                             const float* extra_tok, const float* pos, int B, int T, int F, int D,
                             int num_extra, int order, tpat_stream_t stream) {
  using namespace tpat;
  TPAT_CHECK(spec && patches, "tpat_patchify: null pointer");
  TPAT_CHECK(B > 0 && T > 0 && F > 0 && T % 16 == 0 && F % 16 == 0 && F <= 512,
             "tpat_patchify: need T %% 16 == 0, F %% 16 == 0, F <= 512 (got B=%d T=%d F=%d)", B, T, F);
  TPAT_CHECK(order == TPAT_TOKENS_TIME_MAJOR || order == TPAT_TOKENS_FREQ_MAJOR, "tpat_patchify: bad order %d", order);
  TPAT_CHECK(out_dtype == TPAT_F32 || out_dtype == TPAT_BF16, "tpat_patchify: bad dtype %d", out_dtype);
  TPAT_CHECK(tokens == nullptr || (extra_tok && pos && num_extra >= 0), "tpat_patchify: tokens needs extra_tok and pos");
  TPAT_CHECK(aligned16(spec), "tpat_patchify: spec must be 16-byte aligned");
  dim3 grid(T / 16, B);
  const size_t smem = (size_t)16 * (F + 1) * sizeof(float);
  if (out_dtype == TPAT_F32)
    TPAT_CUDA(launch_kernel(patchify_kernel<float>, dim3(grid), dim3(256), smem, as_stream(stream), spec, (float*)patches, tokens, extra_tok, pos, T, F, D, num_extra, order));
  else
    TPAT_CUDA(launch_kernel(patchify_kernel<__nv_bfloat16>, dim3(grid), dim3(256), smem, as_stream(stream), spec, (__nv_bfloat16*)patches, tokens, extra_tok, pos, T, F, D, num_extra, order));
  TPAT_LAUNCH_CHECK();
  return 0;
}
