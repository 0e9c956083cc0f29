// Internal interface of the attention translation units.
#pragma once
#include "common.cuh"

namespace tpat {

int attention_simt_qtiles(int N);
// lse (optional, training): [B, H, N] fp32 = log sum_j exp(scale * q_i . k_j)
int attention_simt(const void* qkv, void* out, int dtype, float* score_partial, int score_mode, int B, int N, int H,
                   int num_extra, float scale, cudaStream_t st, float* lse = nullptr);
// backward: dqkv [B * N, 3 * H * 64] (dtype) from qkv, out (= O), d_out and lse  (CUDA-core fp32 math)
int attention_bwd_simt(const void* qkv, const void* out, const void* d_out, const float* lse, void* dqkv, int dtype, int B, int N,
                       int H, float scale, float* delta_ws, cudaStream_t st);

int attention_tc_qtiles(int N);
// qk_planes != NULL (score blocks only): q / k as split-bf16 planes [B * N, 4 * H * 64] = [q_hi k_hi | q_lo k_lo]
int attention_tc(const void* qkv, void* out, float* score_partial, int score_mode, int B, int N, int H,
                 int num_extra, float scale, cudaStream_t st, const void* qk_planes = nullptr, float* lse = nullptr);

// tcgen05 backward (bf16 operands); falls back to the CUDA-core kernels until attention_bwd_tc.cu provides it
int attention_bwd_tc(const void* qkv, const void* out, const void* d_out, const float* lse, void* dqkv, int B, int N, int H,
                     float scale, float* delta_ws, float* dbias, cudaStream_t st);
// dst[c] += sum over the nparts rows of partials[nparts][C], fixed order (backward_rows.cu)
int finish_colsum_partials(const float* partials, int nparts, int C, float* dst, cudaStream_t st);

}  // namespace tpat
