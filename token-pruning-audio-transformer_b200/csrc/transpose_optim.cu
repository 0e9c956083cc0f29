// Transposes and the fused AdamW step of the fine-tune path.
//
//   tpat_transpose   dst[c][r] = cast(src[r][c]) with the destination row padded (zero-filled) to ld_dst: the
//                    [tokens, channels] -> [channels, tokens] operand copies that turn the weight gradient
//                    dW = dY^T X (reduction over the tokens) into the forward's "A W^T" GEMM shape, and the
//                    W -> W^T operand copies of the data gradient dX = dY W.
//   tpat_adamw       torch.optim.AdamW's update (the reference's optimizer, audiomae/main_finetune.py:478 over the
//                    layer-wise-lr-decay groups of util/lr_decay.py:15-75) over ONE flat fp32 parameter / gradient /
//                    moment buffer: a chunk table maps 16 Ki-element chunks to their parameter group (lr scale, weight
//                    decay), so the 151 tensors are updated by one launch; it also refreshes the bf16 operand copy.
#include "common.cuh"

namespace tpat {

template <typename SrcT, typename DstT>
__global__ void __launch_bounds__(256)
transpose_kernel(const SrcT* __restrict__ src, int ld_src, DstT* __restrict__ dst, int ld_dst, int rows, int cols) {
  __shared__ float tile[64][65];
  pdl_trigger();
  pdl_wait();
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;      // 64 x 4
#pragma unroll
  for (int i = 0; i < 64; i += 4) {
    const int r = r0 + ty + i, c = c0 + tx;
    tile[ty + i][tx] = (r < rows && c < cols) ? to_f32<SrcT>(src[(size_t)r * ld_src + c]) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 64; i += 4) {
    const int c = c0 + ty + i, r = r0 + tx;                    // dst row = c, dst col = r
    if (c < cols && r < ld_dst) dst[(size_t)c * ld_dst + r] = from_f32<DstT>(tile[tx][ty + i]);   // r >= rows: zero padding
  }
}

struct AdamWParams {
  float* p; const float* g; float* m; float* v; __nv_bfloat16* p_bf16;
  const int4* chunks;        // (offset, length, group, unused)
  const float2* groups;      // (lr scale, weight decay)
  float lr, beta1, beta2, eps, bc1, bc2, grad_scale;
  const int* step_dev;       // optional: step counter in device memory (CUDA-graph replays: the host cannot pass a new step)
};

__global__ void __launch_bounds__(256)
adamw_kernel(const AdamWParams a) {
  const int4 ch = __ldg(a.chunks + blockIdx.x);
  const float2 gr = __ldg(a.groups + ch.z);
  const float lr = a.lr * gr.x, wd = gr.y;
  float bc1 = a.bc1, bc2 = a.bc2;
  if (a.step_dev != nullptr) {
    const float t = (float)__ldg(a.step_dev);
    bc1 = 1.0f - powf(a.beta1, t);
    bc2 = 1.0f - powf(a.beta2, t);
  }
  const float step = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  const size_t base = (size_t)(unsigned)ch.x;
  for (int i = threadIdx.x; i < ch.y; i += blockDim.x) {
    const size_t k = base + i;
    const float g = a.g[k] * a.grad_scale;
    float p = a.p[k];
    p *= 1.0f - lr * wd;                                       // decoupled weight decay
    const float m = a.beta1 * a.m[k] + (1.0f - a.beta1) * g;
    const float v = a.beta2 * a.v[k] + (1.0f - a.beta2) * g * g;
    p -= step * m / (sqrtf(v) * inv_sqrt_bc2 + a.eps);
    a.p[k] = p; a.m[k] = m; a.v[k] = v;
    if (a.p_bf16) a.p_bf16[k] = __float2bfloat16_rn(p);
  }
}

__global__ void counter_inc_kernel(int* ctr) { *ctr += 1; }

// out[i] (+)= sum_b x[b * n + i]  (gradient of a parameter broadcast over the clips: cls / dist token, pos_embed)
__global__ void __launch_bounds__(256)
batch_sum_kernel(const float* __restrict__ x, float* __restrict__ out, int B, size_t stride, int n, int accumulate) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s0 = 0.f, s1 = 0.f;
  int b = 0;
  for (; b + 1 < B; b += 2) { s0 += x[(size_t)b * stride + i]; s1 += x[(size_t)(b + 1) * stride + i]; }
  if (b < B) s0 += x[(size_t)b * stride + i];
  out[i] = (accumulate ? out[i] : 0.f) + (s0 + s1);
}

}  // namespace tpat

extern "C" int tpat_transpose(const void* src, int src_dtype, int ld_src, void* dst, int dst_dtype, int ld_dst, int rows,
                              int cols, tpat_stream_t stream) {
  using namespace tpat;
  TPAT_CHECK(src && dst, "tpat_transpose: null pointer");
  TPAT_CHECK(rows > 0 && cols > 0 && ld_src >= cols && ld_dst >= rows, "tpat_transpose: bad sizes rows=%d cols=%d ld_src=%d ld_dst=%d", rows, cols, ld_src, ld_dst);
  const dim3 grid((cols + 63) / 64, (ld_dst + 63) / 64);
  cudaStream_t st = as_stream(stream);
  if (src_dtype == TPAT_F32 && dst_dtype == TPAT_F32) TPAT_CUDA(launch_kernel(transpose_kernel<float, float>, dim3(grid), dim3(256), 0, st, (const float*)src, ld_src, (float*)dst, ld_dst, rows, cols));
  else if (src_dtype == TPAT_F32 && dst_dtype == TPAT_BF16) TPAT_CUDA(launch_kernel(transpose_kernel<float, __nv_bfloat16>, dim3(grid), dim3(256), 0, st, (const float*)src, ld_src, (__nv_bfloat16*)dst, ld_dst, rows, cols));
  else if (src_dtype == TPAT_BF16 && dst_dtype == TPAT_BF16) TPAT_CUDA(launch_kernel(transpose_kernel<__nv_bfloat16, __nv_bfloat16>, dim3(grid), dim3(256), 0, st, (const __nv_bfloat16*)src, ld_src, (__nv_bfloat16*)dst, ld_dst, rows, cols));
  else { set_error("tpat_transpose: unsupported dtype pair %d -> %d", src_dtype, dst_dtype); return 1; }
  TPAT_LAUNCH_CHECK();
  return 0;
}

extern "C" int tpat_counter_inc(int32_t* counter, tpat_stream_t stream) {
  using namespace tpat;
  TPAT_CHECK(counter != nullptr, "tpat_counter_inc: null pointer");
  TPAT_CUDA(launch_kernel(counter_inc_kernel, dim3(1), dim3(1), 0, as_stream(stream), counter));
  TPAT_LAUNCH_CHECK();
  return 0;
}

extern "C" int tpat_adamw(float* p, const float* g, float* m, float* v, void* p_bf16, const int32_t* chunks, int n_chunks,
                          const float* groups, float lr, float beta1, float beta2, float eps, int step, const int32_t* step_dev,
                          float grad_scale, tpat_stream_t stream) {
  using namespace tpat;
  TPAT_CHECK(p && g && m && v && chunks && groups && n_chunks >= 0 && (step >= 1 || step_dev != nullptr), "tpat_adamw: bad arguments");
  if (n_chunks == 0) return 0;
  const float st = (float)(step >= 1 ? step : 1);
  AdamWParams a{p, g, m, v, (__nv_bfloat16*)p_bf16, reinterpret_cast<const int4*>(chunks), reinterpret_cast<const float2*>(groups),
                lr, beta1, beta2, eps, 1.0f - powf(beta1, st), 1.0f - powf(beta2, st), grad_scale, step_dev};
  TPAT_CUDA(launch_kernel(adamw_kernel, dim3(n_chunks), dim3(256), 0, as_stream(stream), a));
  TPAT_LAUNCH_CHECK();
  return 0;
}

extern "C" int tpat_batch_sum(const float* x, float* out, int B, size_t stride, int n, int accumulate, tpat_stream_t stream) {
  using namespace tpat;
  TPAT_CHECK(x && out && B > 0 && n > 0 && stride >= (size_t)n, "tpat_batch_sum: bad arguments");
  TPAT_CUDA(launch_kernel(batch_sum_kernel, dim3((n + 255) / 256), dim3(256), 0, as_stream(stream), x, out, B, stride, n, accumulate));
  TPAT_LAUNCH_CHECK();
  return 0;
}
