// CUDA-core fused attention + importance-score partials (the fp32 parity path, also usable with
// bf16 operands as the on-device cross-check of the tcgen05 kernel).
//
// Replaces, for one (clip, head, 32-query tile): q k^T * scale, softmax over all N keys, attn @ v
// and the score slice (reference audiomae/models_vit.py:79-95,113; ast/src/models/ast_models.py:92-109,124).
// The [B,H,N,N] attention matrix never reaches HBM: the 32 x N row block of S / P lives in shared
// memory, softmax is exact (row max subtracted, normalised in fp32 before P.V, like torch.softmax).
//   COLMEAN partial: per (head, query tile) column sums of P over query rows >= extra
//   CLS_ROW partial: row 0 of P per head
// Both are reduced later in a fixed order by tpat_score_topk (no atomics -> deterministic).
#include "attention.cuh"

namespace tpat {

constexpr int AS_QT = 32;    // queries per CTA
constexpr int AS_KC = 64;    // keys per chunk
constexpr int AS_HD = 64;    // head dim

template <typename T> struct Ld4;
template <> struct Ld4<float> {
  static __device__ __forceinline__ void ld(const float* p, float (&o)[4]) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
};
template <> struct Ld4<__nv_bfloat16> {
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float (&o)[4]) {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&v.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&v.y);
    o[0] = __low2float(a); o[1] = __high2float(a); o[2] = __low2float(b); o[3] = __high2float(b);
  }
};

template <typename T>
__global__ void __launch_bounds__(256)
attention_simt_kernel(const T* __restrict__ qkv, T* __restrict__ out, float* __restrict__ score_partial, int score_mode,
                      int N, int H, int num_extra, float scale, int n_qt, float* __restrict__ lse) {
  extern __shared__ float sm[];
  pdl_trigger();
  pdl_wait();
  const int Npad = (N + AS_KC - 1) / AS_KC * AS_KC;
  const int lds = Npad + 4;
  float* Ss = sm;                              // [QT][lds]
  float* Qs = Ss + AS_QT * lds;                // [QT][HD+1]
  float* KVs = Qs + AS_QT * (AS_HD + 1);       // [64][HD+4]: K^T chunk ([d][key]) or V chunk ([key][d])
  constexpr int ldkv = AS_HD + 4;

  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x;
  const int q0 = qt * AS_QT;
  const int ldq = 3 * H * AS_HD;
  const T* base = qkv + (size_t)b * N * ldq;
  const T* Qg = base + h * AS_HD;
  const T* Kg = base + (H + h) * AS_HD;
  const T* Vg = base + (2 * H + h) * AS_HD;

  // Q tile -> smem (fp32)
  for (int i = tid; i < AS_QT * (AS_HD / 4); i += 256) {
    const int r = i / (AS_HD / 4), dv = (i % (AS_HD / 4)) * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (q0 + r < N) Ld4<T>::ld(Qg + (size_t)(q0 + r) * ldq + dv, v);
    float* d = Qs + r * (AS_HD + 1) + dv;
    d[0] = v[0]; d[1] = v[1]; d[2] = v[2]; d[3] = v[3];
  }

  const int ty = tid >> 4, tx = tid & 15;  // rows ty*2..+1, cols tx*4..+3 of a 32 x 64 chunk
  // ---- S = (Q K^T) * scale, chunk by chunk ----
  for (int k0 = 0; k0 < Npad; k0 += AS_KC) {
    __syncthreads();  // previous chunk consumed (also covers the Q store on the first pass)
    for (int i = tid; i < AS_KC * (AS_HD / 4); i += 256) {
      const int key = i % AS_KC, dv = (i / AS_KC) * 4;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (k0 + key < N) Ld4<T>::ld(Kg + (size_t)(k0 + key) * ldq + dv, v);
#pragma unroll
      for (int j = 0; j < 4; ++j) KVs[(dv + j) * ldkv + key] = v[j];
    }
    __syncthreads();
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll 8
    for (int d = 0; d < AS_HD; ++d) {
      const float a0 = Qs[(ty * 2) * (AS_HD + 1) + d], a1 = Qs[(ty * 2 + 1) * (AS_HD + 1) + d];
      const float4 kk = *reinterpret_cast<const float4*>(&KVs[d * ldkv + tx * 4]);
      acc[0][0] = fmaf(a0, kk.x, acc[0][0]); acc[0][1] = fmaf(a0, kk.y, acc[0][1]);
      acc[0][2] = fmaf(a0, kk.z, acc[0][2]); acc[0][3] = fmaf(a0, kk.w, acc[0][3]);
      acc[1][0] = fmaf(a1, kk.x, acc[1][0]); acc[1][1] = fmaf(a1, kk.y, acc[1][1]);
      acc[1][2] = fmaf(a1, kk.z, acc[1][2]); acc[1][3] = fmaf(a1, kk.w, acc[1][3]);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
      *reinterpret_cast<float4*>(&Ss[(ty * 2 + i) * lds + k0 + tx * 4]) =
          make_float4(acc[i][0] * scale, acc[i][1] * scale, acc[i][2] * scale, acc[i][3] * scale);
  }
  __syncthreads();

  // ---- exact softmax per row: warp w owns rows w*4 .. w*4+3 ----
  {
    const int w = tid >> 5, lane = tid & 31;
    for (int rr = 0; rr < 4; ++rr) {
      float* row = Ss + (w * 4 + rr) * lds;
      float mx = -INFINITY;
      for (int j = lane; j < N; j += 32) mx = fmaxf(mx, row[j]);
      mx = warp_max(mx);
      float sum = 0.f;
      for (int j = lane; j < N; j += 32) { const float e = expf(row[j] - mx); row[j] = e; sum += e; }
      sum = warp_sum(sum);
      // training: log-sum-exp of the scaled scores of this query row (the attention backward recomputes P from it)
      if (lse != nullptr && lane == 0 && q0 + w * 4 + rr < N) lse[((size_t)b * H + h) * N + q0 + w * 4 + rr] = mx + logf(sum);
      const float inv = 1.0f / sum;
      for (int j = lane; j < Npad; j += 32) row[j] = (j < N) ? row[j] * inv : 0.f;
    }
  }
  __syncthreads();

  // ---- importance-score partials from the fp32 probabilities ----
  if (score_mode == TPAT_SCORE_COLMEAN) {
    float* dst = score_partial + ((size_t)b * H * n_qt + (size_t)h * n_qt + qt) * N;
    for (int j = tid; j < N; j += 256) {
      float s = 0.f;
      for (int r = 0; r < AS_QT; ++r) {
        const int qi = q0 + r;
        if (qi >= num_extra && qi < N) s += Ss[r * lds + j];
      }
      dst[j] = s;
    }
  } else if (score_mode == TPAT_SCORE_CLS_ROW && qt == 0) {
    float* dst = score_partial + ((size_t)b * H + h) * N;
    for (int j = tid; j < N; j += 256) dst[j] = Ss[j];
  }
  if constexpr (sizeof(T) == 2) {
    // bf16 mode: P is rounded to bf16 before P.V (what the tensor-core kernel feeds the MMA)
    __syncthreads();
    for (int i = tid; i < AS_QT * Npad; i += 256) {
      float* p = Ss + (i / Npad) * lds + (i % Npad);
      *p = __bfloat162float(__float2bfloat16_rn(*p));
    }
  }

  // ---- O = P V ----
  float o[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  for (int k0 = 0; k0 < Npad; k0 += AS_KC) {
    __syncthreads();
    for (int i = tid; i < AS_KC * (AS_HD / 4); i += 256) {
      const int key = i / (AS_HD / 4), dv = (i % (AS_HD / 4)) * 4;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (k0 + key < N) Ld4<T>::ld(Vg + (size_t)(k0 + key) * ldq + dv, v);
      *reinterpret_cast<float4*>(&KVs[key * ldkv + dv]) = make_float4(v[0], v[1], v[2], v[3]);
    }
    __syncthreads();
#pragma unroll 8
    for (int key = 0; key < AS_KC; ++key) {
      const float p0 = Ss[(ty * 2) * lds + k0 + key], p1 = Ss[(ty * 2 + 1) * lds + k0 + key];
      const float4 vv = *reinterpret_cast<const float4*>(&KVs[key * ldkv + tx * 4]);
      o[0][0] = fmaf(p0, vv.x, o[0][0]); o[0][1] = fmaf(p0, vv.y, o[0][1]);
      o[0][2] = fmaf(p0, vv.z, o[0][2]); o[0][3] = fmaf(p0, vv.w, o[0][3]);
      o[1][0] = fmaf(p1, vv.x, o[1][0]); o[1][1] = fmaf(p1, vv.y, o[1][1]);
      o[1][2] = fmaf(p1, vv.z, o[1][2]); o[1][3] = fmaf(p1, vv.w, o[1][3]);
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int qi = q0 + ty * 2 + i;
    if (qi >= N) continue;
    T* dst = out + ((size_t)b * N + qi) * (H * AS_HD) + h * AS_HD + tx * 4;
    if constexpr (sizeof(T) == 4) {
      *reinterpret_cast<float4*>(dst) = make_float4(o[i][0], o[i][1], o[i][2], o[i][3]);
    } else {
      *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(o[i][0], o[i][1]), pack_bf16x2(o[i][2], o[i][3]));
    }
  }
}

int attention_simt_qtiles(int N) { return (N + AS_QT - 1) / AS_QT; }

int attention_simt(const void* qkv, void* out, int dtype, float* score_partial, int score_mode, int B, int N, int H,
                   int num_extra, float scale, cudaStream_t st, float* lse) {
  const int n_qt = (N + AS_QT - 1) / AS_QT;
  const int Npad = (N + AS_KC - 1) / AS_KC * AS_KC;
  const size_t smem = ((size_t)AS_QT * (Npad + 4) + AS_QT * (AS_HD + 1) + 64 * (AS_HD + 4)) * sizeof(float);
  TPAT_CHECK(smem <= 200 * 1024, "tpat_attention(simt): N=%d needs %zu bytes of shared memory", N, smem);
  dim3 grid(n_qt, H, B);
  if (dtype == TPAT_F32) {
    auto kern = attention_simt_kernel<float>;
    TPAT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    TPAT_CUDA(launch_kernel(kern, dim3(grid), dim3(256), smem, st, (const float*)qkv, (float*)out, score_partial, score_mode, N, H, num_extra, scale, n_qt, lse));
  } else {
    auto kern = attention_simt_kernel<__nv_bfloat16>;
    TPAT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    TPAT_CUDA(launch_kernel(kern, dim3(grid), dim3(256), smem, st, (const __nv_bfloat16*)qkv, (__nv_bfloat16*)out, score_partial, score_mode, N, H, num_extra, scale, n_qt, lse));
  }
  TPAT_LAUNCH_CHECK();
  return 0;
}

}  // namespace tpat
