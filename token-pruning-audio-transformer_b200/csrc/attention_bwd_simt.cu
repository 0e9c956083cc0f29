// CUDA-core attention backward (fp32 math; the parity path of the fine-tune step and the on-device cross-check of
// the tcgen05 backward).  Backward of q k^T * scale -> softmax -> attn @ v (reference audiomae/models_vit.py:75-95 under
// autograd; the importance score and the top-k indices carry no gradient, :113-114 are index computations):
//     P = exp(scale * Q K^T - lse),  dP = dO V^T,  delta_i = sum_d dO_id O_id,  dS = P o (dP - delta) * scale,
//     dQ = dS K,   dK = dS^T Q,   dV = P^T dO.
// P is recomputed from the saved log-sum-exp; nothing of size N x N touches HBM.  Two launches of one kernel template:
//   DKDV = false  CTA = (clip, head, 32 QUERY rows), streams the keys in chunks of 64:  dQ rows
//   DKDV = true   CTA = (clip, head, 32 KEY rows),   streams the queries in chunks of 64: dK and dV rows
// so every output row is written by exactly one CTA in a fixed order (no atomics -> deterministic).
#include "attention.cuh"

namespace tpat {

constexpr int AB_T = 32;     // tile rows
constexpr int AB_C = 64;     // streamed rows per chunk
constexpr int AB_HD = 64;
constexpr int AB_LDA = AB_HD + 1;   // tile operands   [32][65]
constexpr int AB_LDT = AB_C + 4;    // transposed chunk [d][68]
constexpr int AB_LDN = AB_HD + 4;   // natural chunk    [row][68]
constexpr int AB_LDP = AB_C + 4;    // P / dS chunk     [32][68]

template <typename T> __device__ __forceinline__ void ab_ld4(const T* p, float (&o)[4]);
template <> __device__ __forceinline__ void ab_ld4<float>(const float* p, float (&o)[4]) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <> __device__ __forceinline__ void ab_ld4<__nv_bfloat16>(const __nv_bfloat16* p, float (&o)[4]) {
  const uint2 v = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&v.x), b = *reinterpret_cast<const __nv_bfloat162*>(&v.y);
  o[0] = __low2float(a); o[1] = __high2float(a); o[2] = __low2float(b); o[3] = __high2float(b);
}

// delta[b, h, i] = sum_d dO[b, i, h, d] * O[b, i, h, d]: eight lanes per (row, head), 8 elements each, so that a warp
// instruction covers four contiguous 128-byte (bf16) head rows; shuffle reduction, lane 0 of the group writes
template <typename T>
__global__ void __launch_bounds__(256)
attn_delta_kernel(const T* __restrict__ o, const T* __restrict__ d_o, float* __restrict__ delta, int B, int N, int H) {
  pdl_trigger();
  pdl_wait();
  const size_t total = (size_t)B * N * H;
  const int sub = threadIdx.x & 7;
  for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3; i < total; i += ((size_t)gridDim.x * blockDim.x) >> 3) {
    const int h = (int)(i % H);
    const size_t row = i / H;                   // b * N + n
    const T* po = o + row * (size_t)(H * AB_HD) + h * AB_HD + sub * 8;
    const T* pd = d_o + row * (size_t)(H * AB_HD) + h * AB_HD + sub * 8;
    float a[4], g[4], a2[4], g2[4];
    ab_ld4<T>(po, a); ab_ld4<T>(pd, g); ab_ld4<T>(po + 4, a2); ab_ld4<T>(pd + 4, g2);
    float s = ((a[0] * g[0] + a[1] * g[1]) + (a[2] * g[2] + a[3] * g[3])) + ((a2[0] * g2[0] + a2[1] * g2[1]) + (a2[2] * g2[2] + a2[3] * g2[3]));
    s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (sub == 0) {
      const int b = (int)(row / N), n = (int)(row % N);
      delta[((size_t)b * H + h) * N + n] = s;
    }
  }
}

// bf16 flavour with 16-byte loads: lane = 8 elements of one (row, head), one uint4 of O and one of dO each; two
// (row, head) pairs per thread in flight
__global__ void __launch_bounds__(256)
attn_delta8_kernel(const uint4* __restrict__ o, const uint4* __restrict__ d_o, float* __restrict__ delta, int B, int N, int H) {
  pdl_trigger();
  pdl_wait();
  const size_t total8 = (size_t)B * N * H * 8;          // uint4 elements: [row][h][8]
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  auto dot8 = [](const uint4& a, const uint4& g) {
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      s0 = fmaf(__uint_as_float(aw[e] << 16), __uint_as_float(gw[e] << 16), s0);
      s1 = fmaf(__uint_as_float(aw[e] & 0xffff0000u), __uint_as_float(gw[e] & 0xffff0000u), s1);
    }
    return s0 + s1;
  };
  auto finish = [&](float s, size_t j) {
    s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 4);
    if ((j & 7) == 0 && j < total8) {
      const size_t i = j >> 3;
      const int h = (int)(i % H);
      const size_t row = i / H;
      delta[((row / N) * H + h) * (size_t)N + row % N] = s;
    }
  };
  // the loop bound is warp-uniform (the shuffles need all 32 lanes); lanes past the end load nothing
  const uint32_t lane = threadIdx.x & 31;
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j - lane < total8; j += 2 * stride) {
    const size_t j2 = j + stride;
    const bool has1 = j < total8, has2 = j2 < total8;
    uint4 a = make_uint4(0, 0, 0, 0), g = a, a2 = a, g2 = a;
    if (has1) { a = __ldg(o + j); g = __ldg(d_o + j); }
    if (has2) { a2 = __ldg(o + j2); g2 = __ldg(d_o + j2); }
    finish(dot8(a, g), has1 ? j : total8);
    finish(dot8(a2, g2), has2 ? j2 : total8);
  }
}

template <typename T, bool DKDV>
__global__ void __launch_bounds__(256)
attn_bwd_simt_kernel(const T* __restrict__ qkv, const T* __restrict__ d_o, const float* __restrict__ lse,
                     const float* __restrict__ delta, T* __restrict__ dqkv, int N, int H, float scale) {
  extern __shared__ float ab_sm[];
  pdl_trigger();
  pdl_wait();
  float* A1 = ab_sm;                       // [32][65]  tile rows of Q (dq pass) / K (dkdv pass)
  float* A2 = A1 + AB_T * AB_LDA;          // [32][65]  tile rows of dO      / V
  float* B1t = A2 + AB_T * AB_LDA;         // [64 d][68]  chunk of K / Q, transposed
  float* B2t = B1t + AB_HD * AB_LDT;       // [64 d][68]  chunk of V / dO, transposed
  float* B1n = B2t + AB_HD * AB_LDT;       // [64][68]    chunk of K / Q
  float* B2n = B1n + AB_C * AB_LDN;        // [64][68]    chunk of dO (dkdv pass only)
  float* Ps = B2n + AB_C * AB_LDN;         // [32][68]
  float* dSs = Ps + AB_T * AB_LDP;         // [32][68]

  const int tile = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x;
  const int r0 = tile * AB_T;
  const int ldq = 3 * H * AB_HD, ldo = H * AB_HD;
  const T* base = qkv + (size_t)b * N * ldq;
  const T* Qg = base + h * AB_HD;
  const T* Kg = base + (H + h) * AB_HD;
  const T* Vg = base + (2 * H + h) * AB_HD;
  const T* dOg = d_o + (size_t)b * N * ldo + h * AB_HD;
  const float* lse_bh = lse + ((size_t)b * H + h) * N;
  const float* delta_bh = delta + ((size_t)b * H + h) * N;

  // tile operands
  const T* a1g = DKDV ? Kg : Qg;   const int a1ld = ldq;
  const T* a2g = DKDV ? Vg : dOg;  const int a2ld = DKDV ? ldq : ldo;
  for (int i = tid; i < AB_T * (AB_HD / 4); i += 256) {
    const int r = i / (AB_HD / 4), dv = (i % (AB_HD / 4)) * 4;
    float u[4] = {0.f, 0.f, 0.f, 0.f}, v[4] = {0.f, 0.f, 0.f, 0.f};
    if (r0 + r < N) { ab_ld4<T>(a1g + (size_t)(r0 + r) * a1ld + dv, u); ab_ld4<T>(a2g + (size_t)(r0 + r) * a2ld + dv, v); }
#pragma unroll
    for (int j = 0; j < 4; ++j) { A1[r * AB_LDA + dv + j] = u[j]; A2[r * AB_LDA + dv + j] = v[j]; }
  }

  const int ty = tid >> 4, tx = tid & 15;   // rows ty*2..+1, cols tx*4..+3 of a 32 x 64 block
  float acc1[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};   // dQ (dq pass) / dK (dkdv pass)
  float acc2[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};   // dV (dkdv pass)
  const T* b1g = DKDV ? Qg : Kg;    const int b1ld = ldq;
  const T* b2g = DKDV ? dOg : Vg;   const int b2ld = DKDV ? ldo : ldq;

  const int Npad = (N + AB_C - 1) / AB_C * AB_C;
  for (int c0 = 0; c0 < Npad; c0 += AB_C) {
    __syncthreads();                         // previous chunk consumed (covers the tile store on the first pass)
    for (int i = tid; i < AB_C * (AB_HD / 4); i += 256) {
      const int r = i % AB_C, dv = (i / AB_C) * 4;
      float u[4] = {0.f, 0.f, 0.f, 0.f}, v[4] = {0.f, 0.f, 0.f, 0.f};
      if (c0 + r < N) { ab_ld4<T>(b1g + (size_t)(c0 + r) * b1ld + dv, u); ab_ld4<T>(b2g + (size_t)(c0 + r) * b2ld + dv, v); }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        B1t[(dv + j) * AB_LDT + r] = u[j]; B2t[(dv + j) * AB_LDT + r] = v[j];
        B1n[r * AB_LDN + dv + j] = u[j];
        if (DKDV) B2n[r * AB_LDN + dv + j] = v[j];
      }
    }
    __syncthreads();
    // S = A1 . B1^T and dP = A2 . B2^T  (32 x 64)
    float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, dp[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll 8
    for (int d = 0; d < AB_HD; ++d) {
      const float a0 = A1[(ty * 2) * AB_LDA + d], a1 = A1[(ty * 2 + 1) * AB_LDA + d];
      const float g0 = A2[(ty * 2) * AB_LDA + d], g1 = A2[(ty * 2 + 1) * AB_LDA + d];
      const float4 k4 = *reinterpret_cast<const float4*>(&B1t[d * AB_LDT + tx * 4]);
      const float4 v4 = *reinterpret_cast<const float4*>(&B2t[d * AB_LDT + tx * 4]);
      s[0][0] = fmaf(a0, k4.x, s[0][0]); s[0][1] = fmaf(a0, k4.y, s[0][1]); s[0][2] = fmaf(a0, k4.z, s[0][2]); s[0][3] = fmaf(a0, k4.w, s[0][3]);
      s[1][0] = fmaf(a1, k4.x, s[1][0]); s[1][1] = fmaf(a1, k4.y, s[1][1]); s[1][2] = fmaf(a1, k4.z, s[1][2]); s[1][3] = fmaf(a1, k4.w, s[1][3]);
      dp[0][0] = fmaf(g0, v4.x, dp[0][0]); dp[0][1] = fmaf(g0, v4.y, dp[0][1]); dp[0][2] = fmaf(g0, v4.z, dp[0][2]); dp[0][3] = fmaf(g0, v4.w, dp[0][3]);
      dp[1][0] = fmaf(g1, v4.x, dp[1][0]); dp[1][1] = fmaf(g1, v4.y, dp[1][1]); dp[1][2] = fmaf(g1, v4.z, dp[1][2]); dp[1][3] = fmaf(g1, v4.w, dp[1][3]);
    }
    // P = exp(scale * s - lse[query]),  dS = P (dP - delta[query]) scale;  query = tile row (dq pass) / chunk column (dkdv pass)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int tr = r0 + ty * 2 + i;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int cc = c0 + tx * 4 + j;
        const int qi = DKDV ? cc : tr;
        float pv = 0.f, ds = 0.f;
        if (tr < N && cc < N) {
          pv = expf(fmaf(s[i][j], scale, -lse_bh[qi]));
          ds = pv * (dp[i][j] - delta_bh[qi]) * scale;
        }
        Ps[(ty * 2 + i) * AB_LDP + tx * 4 + j] = pv;
        dSs[(ty * 2 + i) * AB_LDP + tx * 4 + j] = ds;
      }
    }
    __syncthreads();
    // acc1 += dS . B1 (natural);  dkdv pass also acc2 += P . B2 (natural)
#pragma unroll 8
    for (int c = 0; c < AB_C; ++c) {
      const float d0 = dSs[(ty * 2) * AB_LDP + c], d1 = dSs[(ty * 2 + 1) * AB_LDP + c];
      const float4 k4 = *reinterpret_cast<const float4*>(&B1n[c * AB_LDN + tx * 4]);
      acc1[0][0] = fmaf(d0, k4.x, acc1[0][0]); acc1[0][1] = fmaf(d0, k4.y, acc1[0][1]); acc1[0][2] = fmaf(d0, k4.z, acc1[0][2]); acc1[0][3] = fmaf(d0, k4.w, acc1[0][3]);
      acc1[1][0] = fmaf(d1, k4.x, acc1[1][0]); acc1[1][1] = fmaf(d1, k4.y, acc1[1][1]); acc1[1][2] = fmaf(d1, k4.z, acc1[1][2]); acc1[1][3] = fmaf(d1, k4.w, acc1[1][3]);
      if (DKDV) {
        const float p0 = Ps[(ty * 2) * AB_LDP + c], p1 = Ps[(ty * 2 + 1) * AB_LDP + c];
        const float4 v4 = *reinterpret_cast<const float4*>(&B2n[c * AB_LDN + tx * 4]);
        acc2[0][0] = fmaf(p0, v4.x, acc2[0][0]); acc2[0][1] = fmaf(p0, v4.y, acc2[0][1]); acc2[0][2] = fmaf(p0, v4.z, acc2[0][2]); acc2[0][3] = fmaf(p0, v4.w, acc2[0][3]);
        acc2[1][0] = fmaf(p1, v4.x, acc2[1][0]); acc2[1][1] = fmaf(p1, v4.y, acc2[1][1]); acc2[1][2] = fmaf(p1, v4.z, acc2[1][2]); acc2[1][3] = fmaf(p1, v4.w, acc2[1][3]);
      }
    }
  }
  // rows of dqkv: dQ -> columns [h*64, ..), dK -> [(H+h)*64, ..), dV -> [(2H+h)*64, ..)
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int tr = r0 + ty * 2 + i;
    if (tr >= N) continue;
    T* row = dqkv + ((size_t)b * N + tr) * ldq;
    auto st4 = [&](T* dst, const float (&v)[4]) {
      if constexpr (sizeof(T) == 4) *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
      else *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
    };
    if (DKDV) { st4(row + (H + h) * AB_HD + tx * 4, acc1[i]); st4(row + (2 * H + h) * AB_HD + tx * 4, acc2[i]); }
    else st4(row + h * AB_HD + tx * 4, acc1[i]);
  }
}

constexpr size_t AB_SMEM = (size_t)(2 * AB_T * AB_LDA + 2 * AB_HD * AB_LDT + 2 * AB_C * AB_LDN + 2 * AB_T * AB_LDP) * sizeof(float);

template <typename T>
static int launch_attn_bwd_simt(const void* qkv, const void* out, const void* d_out, const float* lse, void* dqkv, int B, int N, int H,
                                float scale, float* delta_ws, cudaStream_t st) {
  static DeviceOnce once;
  auto kq = attn_bwd_simt_kernel<T, false>;
  auto kkv = attn_bwd_simt_kernel<T, true>;
  if (once.first()) {
    TPAT_CUDA(cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AB_SMEM));
    TPAT_CUDA(cudaFuncSetAttribute(kkv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AB_SMEM));
    once.mark();
  }
  const size_t total = (size_t)B * N * H;
  const int dgrid = (int)((total * 8 + 255) / 256 < 8192 ? (total * 8 + 255) / 256 : 8192);
  TPAT_CUDA(launch_kernel(attn_delta_kernel<T>, dim3(dgrid), dim3(256), 0, st, (const T*)out, (const T*)d_out, delta_ws, B, N, H));
  const dim3 grid((N + AB_T - 1) / AB_T, H, B);
  TPAT_CUDA(launch_kernel(kq, dim3(grid), dim3(256), AB_SMEM, st, (const T*)qkv, (const T*)d_out, lse, (const float*)delta_ws, (T*)dqkv, N, H, scale));
  TPAT_CUDA(launch_kernel(kkv, dim3(grid), dim3(256), AB_SMEM, st, (const T*)qkv, (const T*)d_out, lse, (const float*)delta_ws, (T*)dqkv, N, H, scale));
  TPAT_LAUNCH_CHECK();
  return 0;
}

int attention_bwd_simt(const void* qkv, const void* out, const void* d_out, const float* lse, void* dqkv, int dtype, int B, int N,
                       int H, float scale, float* delta_ws, cudaStream_t st) {
  if (dtype == TPAT_F32) return launch_attn_bwd_simt<float>(qkv, out, d_out, lse, dqkv, B, N, H, scale, delta_ws, st);
  return launch_attn_bwd_simt<__nv_bfloat16>(qkv, out, d_out, lse, dqkv, B, N, H, scale, delta_ws, st);
}

}  // namespace tpat

namespace tpat {
// delta = rowsum(dO o O) for the tcgen05 backward (same kernel as the CUDA-core path)
int attention_bwd_delta_bf16(const void* out, const void* d_out, float* delta, int B, int N, int H, cudaStream_t st) {
  const size_t total = (size_t)B * N * H;
  const int dgrid = (int)((total * 4 + 255) / 256 < 8192 ? (total * 4 + 255) / 256 : 8192);     // two (row, head) pairs per 8 lanes
  TPAT_CUDA(launch_kernel(attn_delta8_kernel, dim3(dgrid), dim3(256), 0, st, (const uint4*)out, (const uint4*)d_out, delta, B, N, H));
  TPAT_LAUNCH_CHECK();
  return 0;
}
}  // namespace tpat
