// Weight-gradient GEMM on tcgen05: dW[out, in] += dY[tokens, out]^T . X[tokens, in]  (bf16 operands, fp32 accumulate).
//
// The backward of nn.Linear w.r.t. its weight (reference: autograd over audiomae/models_vit.py:41-45,76,96,246) is a
// GEMM whose reduction runs over the TOKENS, i.e. both operands are "MN-major" in their natural row-major layout.
// Instead of transposing the activations (two extra HBM passes per Linear) this kernel feeds them to the tensor
// cores as they are:
//   * TMA boxes of 64 tokens x 64 channels land in shared memory as [64 token rows][128 B] (128-byte swizzle); a CTA's
//     128 channels are two such boxes 8 KB apart (the descriptor's leading-dimension offset);
//   * tcgen05.mma.cta_group::2 with BOTH operands MN-major (instruction-descriptor major bits), M256 N256 K16, 16 tokens
//     = two 8-row groups of 1024 B per step; same CTA-pair / TMEM double-buffer structure as gemm_tc2.cu;
//   * the output is tiny ([768..3072] x [256..3072] = 9..36 tiles of 256 x 256) while the reduction is long (up to 32 832
//     tokens), so the work is split along the tokens (split-K) until it fills the 74 SM pairs, and every partial
//     accumulator is added into the fp32 gradient with TMA reduce operations (cp.reduce.async.bulk.tensor .add):
//     no partial buffers, no second pass, and `dW +=` is exactly the accumulate semantics autograd needs.
#include "gemm.cuh"
#include "ptx_sm100.cuh"

namespace tpat {

int encode_tmap_2d(CUtensorMap* out, const void* gptr, int elem_bytes, uint64_t rows, uint64_t cols, uint64_t pitch_bytes,
                   uint32_t box_rows, uint32_t box_cols, bool swizzle128);

constexpr int WG_STAGES = 4;
constexpr int WG_EPI_WARPS = 8;
constexpr int WG_THREADS = 64 + 32 * WG_EPI_WARPS;     // TMA warp, MMA warp, 8 epilogue warps
constexpr int WG_BOX = 64 * 64 * 2;                     // 8 KB: 64 tokens x 64 channels
constexpr int WG_OP_BYTES = 2 * WG_BOX;                 // 16 KB: this CTA's 128 channels of one operand
constexpr int WG_STAGE_BYTES = 2 * WG_OP_BYTES;         // A + B
constexpr int WG_STAGING = WG_EPI_WARPS * 2 * 4096;     // two 32 x 32 fp32 blocks per epilogue warp
constexpr int WG_SMEM = 1024 + WG_STAGES * WG_STAGE_BYTES + WG_STAGING + 256;

struct WgradParams {
  int tiles_m, tiles_n, splits, nkb, kb_per;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(WG_THREADS, 1)
gemm_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                     const __grid_constant__ CUtensorMap tm_c, const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + WG_STAGES * WG_OP_BYTES;
  uint8_t* staging = smem + WG_STAGES * WG_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + WG_STAGING);
  uint64_t* full_bar = bars;                    // [STAGES]  (leader)
  uint64_t* empty_bar = bars + WG_STAGES;       // [STAGES]  (both, multicast commit)
  uint64_t* acc_full = bars + 2 * WG_STAGES;    // [2]
  uint64_t* acc_empty = acc_full + 2;           // [2]       (leader, 2 x 8 remote arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int tiles = p.tiles_m * p.tiles_n;
  const int units = tiles * p.splits;           // unit u -> (tile = u % tiles, split = u / tiles)

  if (warp == 0 && lane == 0) { ptx::prefetch_tensormap(&tm_a); ptx::prefetch_tensormap(&tm_b); ptx::prefetch_tensormap(&tm_c); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < WG_STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&acc_full[a], 1); ptx::mbar_init(&acc_empty[a], 2 * WG_EPI_WARPS); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc_2cta<512>(tmem_slot);
    ptx::tmem_relinquish_2cta();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  auto kb_range = [&](int u, int& kb0, int& kb1) {
    const int split = u / tiles;
    kb0 = split * p.kb_per;
    kb1 = min(p.nkb, kb0 + p.kb_per);
  };

  if (warp == 0) {
    // ===== TMA producer (both CTAs): 64-token blocks of this CTA's 128 dY channels and 128 X channels =====
    if (ptx::elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int u = cluster_id; u < units; u += num_clusters) {
        const int tile = u % tiles;
        const int m0 = (tile / p.tiles_n) * 256 + (int)rank * 128;
        const int n0 = (tile % p.tiles_n) * 256 + (int)rank * 128;
        int kb0, kb1; kb_range(u, kb0, kb1);
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          const uint32_t full_leader = ptx::mapa_shared(ptx::smem_u32(&full_bar[stage]), 0);
          if (leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * WG_STAGE_BYTES);
          uint8_t* a_dst = smem_a + stage * WG_OP_BYTES;
          uint8_t* b_dst = smem_b + stage * WG_OP_BYTES;
          ptx::tma_load_2d_2cta(a_dst, &tm_a, full_leader, m0, kb * 64);
          ptx::tma_load_2d_2cta(a_dst + WG_BOX, &tm_a, full_leader, m0 + 64, kb * 64);
          ptx::tma_load_2d_2cta(b_dst, &tm_b, full_leader, n0, kb * 64);
          ptx::tma_load_2d_2cta(b_dst + WG_BOX, &tm_b, full_leader, n0 + 64, kb * 64);
          if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA) =====
    if (leader && ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::idesc_bf16_f32(256, 256, 1, 1);     // both operands MN-major
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int u = cluster_id; u < units; u += num_clusters) {
        int kb0, kb1; kb_range(u, kb0, kb1);
        ptx::mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smem_a + stage * WG_OP_BYTES), b_addr = ptx::smem_u32(smem_b + stage * WG_OP_BYTES);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)     // 16 tokens = two 8-row groups of 1024 B; second 64-channel box 8 KB further (LBO)
            ptx::mma_f16_ss_2cta(d_tmem, ptx::smem_desc_sw128(a_addr + ks * 2048, WG_BOX, 1024),
                                 ptx::smem_desc_sw128(b_addr + ks * 2048, WG_BOX, 1024), idesc, (kb > kb0 || ks > 0) ? 1u : 0u);
          ptx::tc_commit_2cta(&empty_bar[stage], 0b11);
          if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
        }
        ptx::tc_commit_2cta(&acc_full[acc], 0b11);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===== epilogue warps 2..9 (both CTAs): this CTA's 128 rows of the 256 x 256 partial tile -> TMA reduce-add =====
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;                     // 0 | 1: chunks cg, cg + 2, cg + 4, cg + 6
    uint8_t* stg = staging + (warp - 2) * 8192;
    int acc = 0; uint32_t acc_phase = 0;
    int buf = 0;
    for (int u = cluster_id; u < units; u += num_clusters) {
      const int tile = u % tiles;
      const int m0 = (tile / p.tiles_n) * 256 + (int)rank * 128 + q * 32;
      const int n0 = (tile % p.tiles_n) * 256;
      int kb0, kb1; kb_range(u, kb0, kb1);
      ptx::mbar_wait(&acc_full[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256;
      const uint32_t rel_leader = ptx::mapa_shared(ptx::smem_u32(&acc_empty[acc]), 0);
#pragma unroll 1
      for (int ci = 0; ci < 4; ++ci) {
        const int c = cg + 2 * ci;
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(taddr_row + c * 32, r);
        ptx::tmem_ld_wait();
        if (ci == 3) {                                  // last TMEM read of this accumulator: hand it back
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(rel_leader);
        }
        if (kb1 <= kb0) continue;                       // (an empty split adds nothing)
        if (lane == 0) ptx::tma_store_wait_read<1>();   // the reduce that last read this buffer has drained it
        __syncwarp();
        uint8_t* rowp = stg + buf * 4096 + lane * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(rowp + ((j ^ (lane & 7)) << 4)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          ptx::tma_reduce_add_2d(&tm_c, stg + buf * 4096, n0 + c * 32, m0);
          ptx::tma_store_commit();
        }
        buf ^= 1;
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) ptx::tma_store_wait<0>();
  }

  ptx::tc_fence_before();
  ptx::cluster_sync();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2cta<512>(tmem_base);
  }
}

// dW[Mo, No] += dY[K, Mo]^T X[K, No]; returns 1 (with no error set) when the shape is not supported by this kernel
int gemm_wgrad_tc(const void* dY, int ld_dy, const void* X, int ldx, float* dW, int ldw, int K, int Mo, int No, cudaStream_t st) {
  if (Mo % 256 != 0 || No % 256 != 0 || K < 1 || sm_count() < 2) return -1;
  CUtensorMap ta, tb, tc;
  if (int rc = encode_tmap_2d(&ta, dY, 2, (uint64_t)K, (uint64_t)Mo, (uint64_t)ld_dy * 2, 64, 64, true)) return rc;
  if (int rc = encode_tmap_2d(&tb, X, 2, (uint64_t)K, (uint64_t)No, (uint64_t)ldx * 2, 64, 64, true)) return rc;
  if (int rc = encode_tmap_2d(&tc, dW, 4, (uint64_t)Mo, (uint64_t)No, (uint64_t)ldw * 4, 32, 32, true)) return rc;
  WgradParams p;
  p.tiles_m = Mo / 256; p.tiles_n = No / 256;
  p.nkb = (K + 63) / 64;
  const int tiles = p.tiles_m * p.tiles_n;
  const int clusters_max = sm_count() / 2;
  // split the token range until the units fill the SM pairs: best wave efficiency over 1 .. 16 splits (each split keeps
  // at least 8 k-blocks so that the pipeline fill and the 256 KB reduce of a unit stay amortised)
  int best_s = 1; double best_eff = 0.0;
  for (int s = 1; s <= 16 && s * 8 <= p.nkb; ++s) {
    const int units = tiles * s;
    const int waves = (units + clusters_max - 1) / clusters_max;
    const double eff = (double)units / ((double)waves * clusters_max);
    if (eff > best_eff + 0.02) { best_eff = eff; best_s = s; }
  }
  p.splits = best_s;
  p.kb_per = (p.nkb + p.splits - 1) / p.splits;
  p.splits = (p.nkb + p.kb_per - 1) / p.kb_per;        // no empty trailing split
  const int units = tiles * p.splits;
  const int clusters = units < clusters_max ? units : clusters_max;
  static DeviceOnce once;
  if (once.first()) {
    TPAT_CUDA(cudaFuncSetAttribute(gemm_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
    once.mark();
  }
  TPAT_CUDA(launch_kernel(gemm_wgrad_tc_kernel, dim3(2 * clusters), dim3(WG_THREADS), (size_t)WG_SMEM, st, ta, tb, tc, p));
  TPAT_LAUNCH_CHECK();
  return 0;
}

}  // namespace tpat

extern "C" int tpat_gemm_wgrad(const void* dY, int ld_dy, const void* X, int ldx, float* dW, int ldw, int K, int Mo, int No,
                               tpat_stream_t stream) {
  using namespace tpat;
  TPAT_CHECK(dY && X && dW, "tpat_gemm_wgrad: null pointer");
  TPAT_CHECK(K > 0 && Mo > 0 && No > 0 && ld_dy >= Mo && ldx >= No && ldw >= No, "tpat_gemm_wgrad: bad sizes K=%d Mo=%d No=%d", K, Mo, No);
  TPAT_CHECK(Mo % 256 == 0 && No % 256 == 0, "tpat_gemm_wgrad: the output must be a multiple of 256 x 256 (Mo=%d No=%d)", Mo, No);
  TPAT_CHECK(aligned16(dY) && aligned16(X) && aligned16(dW) && (ld_dy * 2) % 16 == 0 && (ldx * 2) % 16 == 0 && (ldw * 4) % 16 == 0,
             "tpat_gemm_wgrad: pointers / pitches must be 16-byte aligned");
  const int rc = gemm_wgrad_tc(dY, ld_dy, X, ldx, dW, ldw, K, Mo, No, as_stream(stream));
  if (rc < 0) { set_error("tpat_gemm_wgrad: unsupported shape or device"); return 1; }
  return rc;
}
