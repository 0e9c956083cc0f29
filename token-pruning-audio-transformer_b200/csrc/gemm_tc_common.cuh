// Pieces shared by the 1-CTA (gemm_tc.cu) and CTA-pair (gemm_tc2.cu) tcgen05 GEMM kernels:
// tile constants, kernel parameters, the fast erf GELU and the epilogue of one 128 x 256 accumulator.
#pragma once
#include "gemm.cuh"
#include "ptx_sm100.cuh"

namespace tpat {

int encode_tmap_2d(CUtensorMap* out, const void* gptr, int elem_bytes, uint64_t rows, uint64_t cols, uint64_t pitch_bytes,
                   uint32_t box_rows, uint32_t box_cols, bool swizzle128);
int encode_tmap_2d_c32(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols, uint64_t pitch_bytes);

constexpr int TG_BM = 128, TG_BN = 256, TG_BK = 64, TG_UMMA_K = 16;
// EW epilogue warps (8 or 16): warp e -> TMEM lane quarter (e + 2) % 4 (hardware rule: warp id % 4), group e / 4;
// the EW/4 warps of a quarter interleave the eight 32-column chunks of the 256-column accumulator.
constexpr int tg_staging_bytes(int EW) { return EW * 4096; }   // one 32-row x 128 B transpose buffer per epilogue warp
constexpr int tg_threads(int EW) { return 64 + 32 * EW; }      // TMA warp + MMA warp + EW epilogue warps

struct TcGemmParams {
  int M, N, K;
  void* C; int ldc;
  const float* bias;
  const float* residual; int ldr;
  const float* pos; int P, num_extra;
  int tiles_m, tiles_n;
  int bn;           // columns per tile: 256, or 128 / 64 for small-M launches of the 1-CTA kernel (more, shorter tiles)
  int desc;         // 1 = walk the row tiles from the last one down (see g_walk_desc)
  // LayerNorm fold.  Producer side (residual GEMMs): besides C, write bf16(C) and, for every row and
  // 32-column chunk, the partial moments (sum, sum of squared deviations from the chunk mean).  Consumer side (the GEMM
  // that follows the LayerNorm): A is that bf16 copy, W carries gamma, and the epilogue applies
  //   out[m, n] = rstd[m] * (acc[m, n] - mean[m] * ln_colsum[n]) + bias[n]       (bias already holds W beta + b)
  // with mean / rstd combined from the partial moments of row m.  All pointers NULL = plain GEMM.
  __nv_bfloat16* xb; int ldxb;
  float2* part_out; int part_ld;
  const float2* ln_part; int ln_chunks;
  const float* ln_colsum;
  float ln_eps;
  // training extras (see EpiParams)
  void* dact_out; int ld_dact;
  const void* aux; int ld_aux;
  const float* row_scale; int rows_per_clip;
  int pf_l2;        // 1: tmap_r covers the tensor the epilogue reads with plain loads (residual / saved GELU derivative);
                    //    the TMA warp prefetches each tile's 128 x 256 block into L2 when it starts loading the tile
  float* cs_part;   // DGELU: partial column sums of the output, row (m0 / 32) of [.][N], written by every epilogue warp
  int red_add;      // 1 (TMA residual epilogue, C == R in place, no row scale): blocks of acc + bias are ADDED to C by TMA reduce
                    //    operations in L2 -- the residual never travels through the SM
  int tma_c;        // 1 (bf16 C, bias / bias + GELU epilogues of the CTA-pair kernel): output blocks leave through TMA stores
  int w_kn;         // 1: W is [K, N] row-major; B tiles are 64 x 64 boxes ([64 k rows][128 B of n]), MN-major descriptors
  int debug_skip;   // timing experiments only (TPAT_GEMM_DEBUG_SKIP): 1 = no TMA after the first ring fill, 2 = skip W loads
};

// row-tile index of linear tile `tile` (tiles are numbered row-tile-major)
__device__ __forceinline__ int tc_tile_m(const TcGemmParams& p, int tile) {
  const int mt = tile / p.tiles_n;
  return p.desc ? p.tiles_m - 1 - mt : mt;
}

// GELU for the bf16 tensor-core epilogue: x * sigmoid(z), z = x (c1 + c3 x^2 + c5 x^4) with the coefficients fitted
// (minimax, tools/probes/fit_gelu.py) to the erf form nn.GELU() uses: |error| <= 8.2e-5 absolute over all x, i.e.
// about one bf16 ulp at worst (x ~ -3) and far below it elsewhere -- invisible once the result is rounded to bf16.
// Evaluated as 0.5 x (1 + tanh(z / 2)): 8 instructions with ONE MUFU (tanh.approx.f32) -- libdevice erff is ~30
// instructions and made the fc1 epilogue instruction-issue bound (660 TF/s); the x * rcp(1 + ex2(.)) form (two MUFU)
// left it MUFU-bound (1215 TF/s, XU pipe 48 %); this form measures 1285 TF/s with the same error in every x range
// (tools/probes/gelu_probe.py, profiles/r01f_gelu_probe.txt).  The fp32 parity path keeps erff.
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float x2 = fminf(x * x, 81.0f);              // the fit is valid for |x| <= 9; beyond it the sigmoid is saturated
  float h = fmaf(x2, -2.8633e-4f, 0.036609f);        // (c5 x^2 + c3) / 2
  h = fmaf(h, x2, 0.797786f);                        // ... + c1 / 2
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h * x));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

// gelu_erf_fast and its derivative from ONE tanh: d/dx = 0.5 (1 + t) + 0.5 x (1 - t^2) u'(x), u = x (a0 + a1 x^2 + a2 x^4).
// The training forward stores the derivative (bf16, |gelu'| <= 1.13) instead of the pre-activation, so that the GELU
// backward inside the data-gradient GEMM's epilogue is one multiply (measured: the recompute-in-the-backward form ran
// that GEMM at 704 TF/s against 1316 TF/s without it).
__device__ __forceinline__ void gelu_erf_fast_both(float x, float& y, float& dy) {
  const float x2 = fminf(x * x, 81.0f);
  float h = fmaf(x2, -2.8633e-4f, 0.036609f);
  h = fmaf(h, x2, 0.797786f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h * x));
  const float hx = 0.5f * x;
  y = fmaf(hx, t, hx);
  float du = fmaf(x2, 5.0f * -2.8633e-4f, 3.0f * 0.036609f);
  du = fmaf(du, x2, 0.797786f);
  const float s = fmaf(-t, t, 1.0f);                 // 1 - t^2
  dy = fmaf(hx * s, du, fmaf(0.5f, t, 0.5f));
}

// (mean, rstd) of one row from the partial moments of its 32-element chunks (Chan's pairwise update, fixed order)
__device__ __forceinline__ float2 ln_row_moments(const float2* __restrict__ part, int chunks, float eps) {
  float mean = 0.f, m2 = 0.f, n = 0.f;
  for (int c0 = 0; c0 < chunks; c0 += 8) {        // eight independent loads in flight, then the (sequential) updates
    float2 pc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) pc[u] = c0 + u < chunks ? __ldg(part + c0 + u) : make_float2(0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (c0 + u < chunks) {
        const float mc = pc[u].x * (1.0f / 32.0f);
        const float delta = mc - mean, nn = n + 32.0f;
        mean = fmaf(delta, 32.0f / nn, mean);
        m2 += pc[u].y + delta * delta * (n * 32.0f / nn);
        n = nn;
      }
    }
  }
  return make_float2(mean, rsqrtf(m2 / n + eps));
}

// Per-tile epilogue state that does not depend on the accumulator: loaded BEFORE waiting for the MMA
// so that the global-load latency (bias) is off the critical path.
template <int EW> struct TcEpiPrefetch {
  static constexpr int CSTRIDE = EW / 4;                       // warps per lane quarter
  static constexpr int NCH = (8 + CSTRIDE - 1) / CSTRIDE;      // chunks per warp (the last may be dead)
  float4 bias[NCH];
  float4 colsum[NCH];   // consumer side of the LayerNorm fold
};

template <int EW>
__device__ __forceinline__ void tc_epilogue_prefetch(const TcGemmParams& p, int n0, int cg, int lane, TcEpiPrefetch<EW>& pf) {
  const int jl = lane & 7;
#pragma unroll
  for (int ci = 0; ci < TcEpiPrefetch<EW>::NCH; ++ci) {
    const int c = cg + TcEpiPrefetch<EW>::CSTRIDE * ci;
    const int ncol = n0 + c * 32 + jl * 4;
    pf.bias[ci] = (p.bias != nullptr && c < p.bn / 32 && ncol < p.N) ? __ldg(reinterpret_cast<const float4*>(p.bias + ncol)) : make_float4(0.f, 0.f, 0.f, 0.f);
    pf.colsum[ci] = (p.ln_colsum != nullptr && c < p.bn / 32 && ncol < p.N) ? __ldg(reinterpret_cast<const float4*>(p.ln_colsum + ncol)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// Epilogue of one accumulator for one epilogue warp.
//   taddr_row : TMEM address of this warp's lane quarter at column 0 of the accumulator
//   m0        : global row of this warp's first TMEM lane;  n0 : first global column of the tile
//   cg        : group of this warp within its lane quarter: it owns chunks cg, cg + EW/4, ...
//   stg       : this warp's 4 KB transpose buffer
//   release() : called once (warp-uniformly) when the last tcgen05.ld of the tile has completed
// Per chunk: tcgen05.ld (thread = row) -> XOR-swizzled transpose buffer -> re-read with lanes along
// the row, so every global access is a full 128 B (fp32) / 64 B (bf16) row segment per 8 lanes.
// Residual rows are requested before the TMEM read of the chunk so their latency overlaps it.
template <int EPI, typename OutT, int EW, bool FOLD = false, typename ReleaseFn>
__device__ __forceinline__ void tc_epilogue_tile(const TcGemmParams& p, uint32_t taddr_row, int m0, int n0, int cg,
                                                 uint8_t* stg, int lane, const TcEpiPrefetch<EW>& pf, ReleaseFn release,
                                                 const float2* rowstat = nullptr) {
  constexpr int NCH = TcEpiPrefetch<EW>::NCH, CSTRIDE = TcEpiPrefetch<EW>::CSTRIDE;
  const int jl = lane & 7, rl = lane >> 3;   // coalesced phase: lane -> (16 B piece, row within a group of 4)
  bool released = false;
  // DGELU (bf16): the saved GELU derivatives of chunk ci + 1 are requested while chunk ci is processed (packed bf16,
  // two register buffers), and the TMA warp pulls the tile's block into L2 when it starts on the tile (pf_l2).  Measured
  // (tools/probes/epilogue_probe.py, M = 32 832): loads inside the chunk 0.160 ms; this form 0.144 ms; all three chunks
  // requested before the accumulator wait 0.166 ms (register pressure: spills); no loads at all 0.100 ms.
  constexpr bool kAuxQ = EPI == TPAT_EPI_DGELU && sizeof(OutT) == 2;
  uint2 auxq[2][8];
  auto load_aux = [&](int ci_, uint2 (&dst)[8]) {
    const int c_ = cg + CSTRIDE * ci_;
    const int n_ = n0 + c_ * 32;
#ifdef TPAT_DBG_DGELU_NOAUX
    if (false) {
#else
    if (c_ < p.bn / 32 && n_ < p.N) {
#endif
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int m = m0 + it * 4 + rl;
        dst[it] = m < p.M ? __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.aux) + (size_t)m * p.ld_aux + n_ + jl * 4))
                          : make_uint2(0u, 0u);
      }
    }
  };
  if constexpr (kAuxQ) load_aux(0, auxq[0]);
#pragma unroll
  for (int ci = 0; ci < NCH; ++ci) {
    const int c = cg + CSTRIDE * ci;
    const int n = n0 + c * 32;
    const bool live = c < p.bn / 32 && n < p.N;      // warp-uniform
    const int ncol = n + jl * 4;
    float4 extra[8];
    if constexpr (kAuxQ) { if (ci + 1 < NCH) load_aux(ci + 1, auxq[(ci + 1) & 1]); }
    if constexpr (EPI == TPAT_EPI_BIAS_RESIDUAL) {
      if (live) {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int m = m0 + it * 4 + rl;
          extra[it] = m < p.M ? *reinterpret_cast<const float4*>(p.residual + (size_t)m * p.ldr + ncol) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    } else if constexpr (EPI == TPAT_EPI_DGELU && !kAuxQ) {
      if (live) {   // the saved GELU derivatives are requested before the TMEM read: their latency overlaps it
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int m = m0 + it * 4 + rl;
          if constexpr (sizeof(OutT) == 4) {
            extra[it] = m < p.M ? __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.aux) + (size_t)m * p.ld_aux + ncol)) : make_float4(0.f, 0.f, 0.f, 0.f);
          } else {
            uint2 hh = make_uint2(0u, 0u);
            if (m < p.M) hh = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.aux) + (size_t)m * p.ld_aux + ncol));
            const __nv_bfloat162 ha = *reinterpret_cast<const __nv_bfloat162*>(&hh.x), hb = *reinterpret_cast<const __nv_bfloat162*>(&hh.y);
            extra[it] = make_float4(__low2float(ha), __high2float(ha), __low2float(hb), __high2float(hb));
          }
        }
      }
    } else if constexpr (EPI == TPAT_EPI_BIAS_POS) {
      if (live) {   // the position rows are requested before the TMEM read as well (8 independent loads in flight)
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int m = m0 + it * 4 + rl;
          const int pp = m % p.P;
          extra[it] = m < p.M ? __ldg(reinterpret_cast<const float4*>(p.pos + (size_t)(p.num_extra + pp) * p.ldc + ncol)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
    uint32_t r[32];
    if (live) {
      ptx::tmem_ld_32x32b_x32(taddr_row + c * 32, r);
      ptx::tmem_ld_wait();
    }
    if (!released && (ci == NCH - 1 || c + CSTRIDE >= p.bn / 32 || n + CSTRIDE * 32 >= p.N)) {
      // last TMEM read of this tile is complete: hand the accumulator back to the MMA warp
      released = true;
      ptx::tc_fence_before();
      __syncwarp();
      release();
    }
    if (!live) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<uint4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
    __syncwarp();
    const float4 bb = pf.bias[ci];
    float4 v[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int row = it * 4 + rl;
      const float4 a = *reinterpret_cast<const float4*>(stg + row * 128 + ((jl ^ (row & 7)) << 4));
      if constexpr (FOLD && (EPI == TPAT_EPI_BIAS || EPI == TPAT_EPI_BIAS_GELU)) {   // LayerNorm fold: rstd * (acc - mean * colsum) + bias'
        const float2 rs = rowstat[row];
        const float4 cs = pf.colsum[ci];
        v[it] = make_float4(fmaf(rs.y, fmaf(-rs.x, cs.x, a.x), bb.x), fmaf(rs.y, fmaf(-rs.x, cs.y, a.y), bb.y),
                            fmaf(rs.y, fmaf(-rs.x, cs.z, a.z), bb.z), fmaf(rs.y, fmaf(-rs.x, cs.w, a.w), bb.w));
      } else {
        v[it] = make_float4(a.x + bb.x, a.y + bb.y, a.z + bb.z, a.w + bb.w);
      }
    }
    __syncwarp();   // all lanes have read the transpose buffer: the next chunk may overwrite it
    if constexpr (EPI == TPAT_EPI_BIAS_GELU) {
      if (p.dact_out != nullptr) {            // training: GELU and its derivative from one tanh; the derivative is kept
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          float4 d;
          gelu_erf_fast_both(v[it].x, v[it].x, d.x); gelu_erf_fast_both(v[it].y, v[it].y, d.y);
          gelu_erf_fast_both(v[it].z, v[it].z, d.z); gelu_erf_fast_both(v[it].w, v[it].w, d.w);
          const int m = m0 + it * 4 + rl;
          if (m >= p.M) continue;
          if constexpr (sizeof(OutT) == 4) *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.dact_out) + (size_t)m * p.ld_dact + ncol) = d;
          else *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.dact_out) + (size_t)m * p.ld_dact + ncol) =
                   make_uint2(pack_bf16x2(d.x, d.y), pack_bf16x2(d.z, d.w));
        }
      } else {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          v[it].x = gelu_erf_fast(v[it].x); v[it].y = gelu_erf_fast(v[it].y);
          v[it].z = gelu_erf_fast(v[it].z); v[it].w = gelu_erf_fast(v[it].w);
        }
      }
    } else if constexpr (EPI == TPAT_EPI_DGELU) {
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        if constexpr (kAuxQ) {
          const uint2 hh = auxq[ci & 1][it];
          const __nv_bfloat162 ha = *reinterpret_cast<const __nv_bfloat162*>(&hh.x), hb = *reinterpret_cast<const __nv_bfloat162*>(&hh.y);
          v[it].x *= __low2float(ha); v[it].y *= __high2float(ha); v[it].z *= __low2float(hb); v[it].w *= __high2float(hb);
        } else {
          v[it].x *= extra[it].x; v[it].y *= extra[it].y; v[it].z *= extra[it].z; v[it].w *= extra[it].w;
        }
      }
      if (p.cs_part != nullptr) {
        // bias gradient of the producing Linear: column sums of this warp's 32 rows (rows >= M carry aux = 0), reduced over
        // the four row groups of the coalesced layout; one 128-byte row segment per warp and chunk
        float4 cs = v[0];
#pragma unroll
        for (int it = 1; it < 8; ++it) { cs.x += v[it].x; cs.y += v[it].y; cs.z += v[it].z; cs.w += v[it].w; }
#pragma unroll
        for (int o = 8; o <= 16; o <<= 1) {
          cs.x += __shfl_xor_sync(0xffffffffu, cs.x, o); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, o);
          cs.z += __shfl_xor_sync(0xffffffffu, cs.z, o); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, o);
        }
        if (rl == 0) *reinterpret_cast<float4*>(p.cs_part + (size_t)(m0 >> 5) * p.N + ncol) = cs;
      }
    } else if constexpr (EPI == TPAT_EPI_BIAS_RESIDUAL || EPI == TPAT_EPI_BIAS_POS) {
      if (EPI == TPAT_EPI_BIAS_RESIDUAL && p.row_scale != nullptr) {     // DropPath: per-clip scale of the branch
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int m = min(m0 + it * 4 + rl, p.M - 1);
          const float sc = __ldg(p.row_scale + m / p.rows_per_clip);
          v[it].x *= sc; v[it].y *= sc; v[it].z *= sc; v[it].w *= sc;
        }
      }
#pragma unroll
      for (int it = 0; it < 8; ++it) { v[it].x += extra[it].x; v[it].y += extra[it].y; v[it].z += extra[it].z; v[it].w += extra[it].w; }
    }
    if constexpr (FOLD && EPI == TPAT_EPI_BIAS_RESIDUAL) {
      {                                  // LayerNorm fold, producer side: bf16 copy + partial moments of this chunk
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int m = m0 + it * 4 + rl;
          float sm = (v[it].x + v[it].y) + (v[it].z + v[it].w);
          sm += __shfl_xor_sync(0xffffffffu, sm, 1); sm += __shfl_xor_sync(0xffffffffu, sm, 2); sm += __shfl_xor_sync(0xffffffffu, sm, 4);
          const float mc = sm * (1.0f / 32.0f);
          const float dx = v[it].x - mc, dy = v[it].y - mc, dz = v[it].z - mc, dw = v[it].w - mc;
          float q = (dx * dx + dy * dy) + (dz * dz + dw * dw);
          q += __shfl_xor_sync(0xffffffffu, q, 1); q += __shfl_xor_sync(0xffffffffu, q, 2); q += __shfl_xor_sync(0xffffffffu, q, 4);
          if (m < p.M) {
            *reinterpret_cast<uint2*>(p.xb + (size_t)m * p.ldxb + ncol) = make_uint2(pack_bf16x2(v[it].x, v[it].y), pack_bf16x2(v[it].z, v[it].w));
            if (jl == 0) p.part_out[(size_t)m * p.part_ld + (n >> 5)] = make_float2(sm, q);
          }
        }
      }
    }
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int m = m0 + it * 4 + rl;
      if (m >= p.M) continue;
      size_t orow = (size_t)m;
      if constexpr (EPI == TPAT_EPI_BIAS_POS) {
        const int b = m / p.P, pp = m - b * p.P;
        orow = (size_t)b * (p.num_extra + p.P) + p.num_extra + pp;
      }
#ifdef TPAT_DBG_EPI_NOSTORE
      if (v[it].x == 123456.789f)      // (timing experiment: the stores are compiled in but never executed)
#endif
      if constexpr (sizeof(OutT) == 4) {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.C) + orow * p.ldc + ncol) = v[it];
      } else {
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.C) + orow * p.ldc + ncol) =
            make_uint2(pack_bf16x2(v[it].x, v[it].y), pack_bf16x2(v[it].z, v[it].w));
      }
    }
  }
}

// Epilogue of one accumulator for one epilogue warp, bf16 output through TMA stores (bias / bias + GELU; gemm_tc2.cu).
// thread = row all the way: tcgen05.ld -> + bias (, GELU) -> 16 packed bf16 pairs -> this row's 64 bytes of a
// [32 rows][64 B] block in shared memory (64B swizzle: 16-byte chunk position XOR (row / 2) % 4, conflict-free) -> ONE
// cp.async.bulk.tensor store per 32 x 32 block, issued by lane 0.  The warp's 4 KB staging buffer holds two such blocks:
// block k + 2 reuses the half of block k once its store has drained it (wait_group.read 1).  No transpose through shared
// memory, no per-lane global stores.  Motivation: with the stores compiled out the generic epilogue runs the qkv GEMM 9 %
// and fc1 + GELU 12 % faster (profiles/r02p_epilogue_probe_variants.txt).  Result (profiles/r02ac_gemm_tma_store_epilogue_ab.txt):
// no gain -- qkv 3 % slower, fc1 equal -- so what the stores cost is their L2 / HBM write traffic, not instruction issue.
// Same arithmetic, same bits; opt-in (TPAT_GEMM_TMA_STORE=1).
//   kcount : running block counter of this warp (selects the staging half)
template <int EPI, int EW, typename ReleaseFn>
__device__ __forceinline__ void tc_epilogue_tile_tma(const TcGemmParams& p, const CUtensorMap* tmap_c, uint32_t taddr_row, int m0, int n0,
                                                     int cg, uint8_t* stg, int lane, int& kcount, ReleaseFn release) {
  constexpr int NCH = TcEpiPrefetch<EW>::NCH, CSTRIDE = TcEpiPrefetch<EW>::CSTRIDE;
  bool released = false;
#pragma unroll
  for (int ci = 0; ci < NCH; ++ci) {
    const int c = cg + CSTRIDE * ci;
    const int n = n0 + c * 32;
    const bool live = c < p.bn / 32 && n < p.N;      // warp-uniform
    uint32_t r[32];
    if (live) {
      ptx::tmem_ld_32x32b_x32(taddr_row + c * 32, r);
      ptx::tmem_ld_wait();
    }
    if (!released && (ci == NCH - 1 || c + CSTRIDE >= p.bn / 32 || n + CSTRIDE * 32 >= p.N)) {
      released = true;                               // last TMEM read of this tile: hand the accumulator back
      ptx::tc_fence_before();
      __syncwarp();
      release();
    }
    if (!live) continue;
    uint8_t* buf = stg + (kcount & 1) * 2048;
    if (lane == 0) ptx::tma_store_wait_read<1>();    // the store issued two blocks ago has read this half
    __syncwarp();
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 bb = p.bias != nullptr ? __ldg(reinterpret_cast<const float4*>(p.bias + n + 4 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
      float x0 = __uint_as_float(r[4 * j]) + bb.x, x1 = __uint_as_float(r[4 * j + 1]) + bb.y;
      float x2 = __uint_as_float(r[4 * j + 2]) + bb.z, x3 = __uint_as_float(r[4 * j + 3]) + bb.w;
      if constexpr (EPI == TPAT_EPI_BIAS_GELU) { x0 = gelu_erf_fast(x0); x1 = gelu_erf_fast(x1); x2 = gelu_erf_fast(x2); x3 = gelu_erf_fast(x3); }
      pk[2 * j] = pack_bf16x2(x0, x1); pk[2 * j + 1] = pack_bf16x2(x2, x3);
    }
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4)
      *reinterpret_cast<uint4*>(buf + lane * 64 + ((c4 ^ ((lane >> 1) & 3)) << 4)) = make_uint4(pk[4 * c4], pk[4 * c4 + 1], pk[4 * c4 + 2], pk[4 * c4 + 3]);
    ptx::fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      if (m0 < p.M) ptx::tma_store_2d(tmap_c, buf, n, m0);     // rows >= M are clipped by the tensor map
      ptx::tma_store_commit();
    }
    ++kcount;
  }
}

}  // namespace tpat
