// Internal interface of the attention translation units.
#pragma once
#include "common.cuh"

namespace tpat {

int attention_simt_qtiles(int N);
int attention_simt(const void* qkv, void* out, int dtype, float* score_partial, int score_mode, int B, int N, int H,
                   int num_extra, float scale, cudaStream_t st);

int attention_tc_qtiles(int N);
// qk_planes != NULL (score blocks only): q / k as split-bf16 planes [B * N, 4 * H * 64] = [q_hi k_hi | q_lo k_lo]
int attention_tc(const void* qkv, void* out, float* score_partial, int score_mode, int B, int N, int H,
                 int num_extra, float scale, cudaStream_t st, const void* qk_planes = nullptr);

}  // namespace tpat
