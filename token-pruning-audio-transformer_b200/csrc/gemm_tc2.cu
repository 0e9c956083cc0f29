// CTA-pair tcgen05 GEMM (cta_group::2): C = epilogue(A W^T + bias), bf16 operands, fp32 accumulate.
//
// Same role as gemm_tc.cu (reference nn.Linear call sites: audiomae/models_vit.py:41-45,76,96,198,205)
// but two CTAs on the SM pair of one TPC cooperate on a 256 x 256 output tile:
//   * each CTA TMA-loads ITS 128 rows of A and ITS 128 of the 256 W rows per k-block (32 KB / stage
//     instead of 48 KB, so 5 stages + twelve epilogue transpose buffers fit) -- the tensor cores of the pair share the W halves, which
//     halves the shared-memory read traffic per SM, the limiter of the 1-CTA kernel (ncu: tensor
//     pipe 67 % active at 1300 TF/s; 96 B/clk UMMA reads + 96 B/clk TMA writes vs a 128 B/clk port);
//   * one thread of the leader CTA (cluster rank 0) issues tcgen05.mma.cta_group::2 (M256 N256 K16);
//     each CTA's TMEM receives its own 128 rows of the accumulator;
//   * TMA completions of both CTAs are credited to the leader's full barrier; tcgen05.commit
//     multicasts "stage free" / "accumulator ready" to both CTAs; both CTAs' epilogue warps arrive
//     remotely on the leader's accumulator-empty barrier.
// Warp roles per CTA and the epilogue are those of gemm_tc.cu (gemm_tc_common.cuh).
#include "gemm_tc_common.cuh"

#include <cstdlib>

namespace tpat {

constexpr int T2_STAGES = 5;
constexpr int T2_EPI_WARPS = 12;
constexpr int T2_THREADS = tg_threads(T2_EPI_WARPS);   // 448
constexpr int T2_A_BYTES = 128 * TG_BK * 2;    // 16 KB: this CTA's 128 rows of A
constexpr int T2_B_BYTES = 128 * TG_BK * 2;    // 16 KB: this CTA's 128 rows of W
constexpr int T2_STAGE_BYTES = T2_A_BYTES + T2_B_BYTES;
constexpr int T2_STAGING_BYTES = tg_staging_bytes(T2_EPI_WARPS);
constexpr int T2_ROWSTAT_BYTES = T2_EPI_WARPS * 32 * 8;   // LayerNorm fold, consumer side: (mean, rstd) of each warp's 32 rows
constexpr int T2_SMEM_BYTES = T2_STAGES * T2_STAGE_BYTES + T2_STAGING_BYTES + 1024 + 512 + T2_ROWSTAT_BYTES;
// residual epilogue (proj, fc2): the fp32 residual chunk is TMA-prefetched into a second 4 KB buffer per warp
constexpr int T2R_STAGES = 4;
constexpr int T2R_STAGING_BYTES = 2 * T2_STAGING_BYTES;
constexpr int T2R_SMEM_BYTES = T2R_STAGES * T2_STAGE_BYTES + T2R_STAGING_BYTES + 1024 + 512;

template <int EPI, typename OutT, bool RES_TMA, bool FOLD, bool TMA_C = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T2_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                const __grid_constant__ CUtensorMap tmap_r, const __grid_constant__ CUtensorMap tmap_c, const TcGemmParams p) {
  // TMA-fed read-modify-write epilogue: short-K residual GEMMs, and the patch-embed GEMM whose "residual" is the
  // position table (row extra + m % P) and whose output rows are shifted by the clip's extra-token rows
  constexpr bool kResTma = RES_TMA && (EPI == TPAT_EPI_BIAS_RESIDUAL || EPI == TPAT_EPI_BIAS_POS);
  constexpr int NSTAGES = kResTma ? T2R_STAGES : T2_STAGES;
  constexpr int STAGING = kResTma ? T2R_STAGING_BYTES : T2_STAGING_BYTES;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment computed as an OFFSET from the __shared__ array so that the compiler keeps the shared
  // address space (a round trip through uintptr_t turns every staging access into a generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + NSTAGES * T2_A_BYTES;
  uint8_t* staging = smem + NSTAGES * T2_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + STAGING);
  uint64_t* full_bar = bars;                    // [STAGES]  (used in the leader CTA)
  uint64_t* empty_bar = bars + NSTAGES;         // [STAGES]  (one per CTA, multicast commit)
  uint64_t* acc_full = bars + 2 * NSTAGES;      // [2]       (one per CTA, multicast commit)
  uint64_t* acc_empty = acc_full + 2;           // [2]       (used in the leader CTA, 2 x 12 remote arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  uint64_t* res_bar = acc_empty + 4;            // [EPI_WARPS][2] residual-chunk arrival (kResTma)

  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int num_tiles = p.tiles_m * p.tiles_n;   // 256 x 256 tiles
  // Tile walk: round robin over the clusters (neighbouring clusters share A through L2), or -- for the consumer of a
  // LayerNorm fold -- one contiguous range per cluster so that consecutive tiles share their row block and the row
  // moments are combined once per row block instead of once per tile.
  const bool contig = FOLD && EPI != TPAT_EPI_BIAS_RESIDUAL;
  const int per_cluster = (num_tiles + num_clusters - 1) / num_clusters;
  const int tile_first = contig ? cluster_id * per_cluster : cluster_id;
  const int tile_step = contig ? 1 : num_clusters;
  const int tile_end = contig ? min(num_tiles, tile_first + per_cluster) : num_tiles;
  const int nkb = p.K / TG_BK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_a);
    ptx::prefetch_tensormap(&tmap_w);
    if (TMA_C) ptx::prefetch_tensormap(&tmap_c);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NSTAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    if (kResTma) for (int i = 0; i < 2 * T2_EPI_WARPS; ++i) ptx::mbar_init(&res_bar[i], 1);
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&acc_full[a], 1); ptx::mbar_init(&acc_empty[a], 2 * T2_EPI_WARPS); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc_2cta<512>(tmem_slot);
    ptx::tmem_relinquish_2cta();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();          // barrier inits + TMEM allocation visible to both CTAs of the pair
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything above touched only on-chip state; global memory from here on

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own A rows + own half of the W rows, credited to the leader's barrier =====
    if (ptx::elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = tile_first; tile < tile_end; tile += tile_step) {
        const int m0 = tc_tile_m(p, tile) * 256 + (int)rank * 128;
        const int n0 = (tile % p.tiles_n) * TG_BN + (int)rank * 128;
        // epilogue operand read with plain loads (fp32 residual of fc2, saved GELU derivative of the data-gradient GEMM): pull
        // this CTA's 128 x 256 block into L2 now; the epilogue gets to it a whole main loop later (measured on the DGELU
        // GEMM: 60 us of 160 were exposed HBM latency of those loads, tools/probes/epilogue_probe.py)
        if (p.pf_l2) ptx::tma_prefetch_l2_2d(&tmap_r, (tile % p.tiles_n) * TG_BN, m0);
        for (int kb = 0; kb < nkb; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          const uint32_t full_leader = ptx::mapa_shared(ptx::smem_u32(&full_bar[stage]), 0);
          const bool first_fill = tile == tile_first && kb < NSTAGES;
          const bool load_a = p.debug_skip != 1 || first_fill;
          const bool load_b = p.debug_skip == 0 || first_fill;
          if (leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * ((load_a ? T2_A_BYTES : 0) + (load_b ? T2_B_BYTES : 0)));
          if (load_a) ptx::tma_load_2d_2cta(smem_a + stage * T2_A_BYTES, &tmap_a, full_leader, kb * TG_BK, m0);
          if (load_b) {
            if (p.w_kn) {      // W[K][N]: this CTA's 128 n-columns as two 64 x 64 boxes, 8 KB apart
              ptx::tma_load_2d_2cta(smem_b + stage * T2_B_BYTES, &tmap_w, full_leader, n0, kb * TG_BK);
              ptx::tma_load_2d_2cta(smem_b + stage * T2_B_BYTES + 8192, &tmap_w, full_leader, n0 + 64, kb * TG_BK);
            } else {
              ptx::tma_load_2d_2cta(smem_b + stage * T2_B_BYTES, &tmap_w, full_leader, kb * TG_BK, n0);
            }
          }
          if (++stage == NSTAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: leader CTA only =====
    if (leader && ptx::elect_one()) {
      const uint32_t idesc = ptx::idesc_bf16_f32(256, TG_BN, 0, p.w_kn ? 1 : 0);
      // MN-major B (w_kn): 16 k-rows per step = two 8-row groups of 1024 B (SBO), 64-column chunks 8 KB apart (LBO)
      const uint32_t b_lbo = p.w_kn ? 8192u : 16u;
      const uint64_t b_kstep = p.w_kn ? 128u : 2u;        // in the descriptor's (address >> 4) field: 2048 B | 32 B
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = tile_first; tile < tile_end; tile += tile_step) {
        ptx::mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * TG_BN;
        for (int kb = 0; kb < nkb; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint64_t a_desc = ptx::smem_desc_sw128(ptx::smem_u32(smem_a + stage * T2_A_BYTES), 16, 1024);
          const uint64_t b_desc = ptx::smem_desc_sw128(ptx::smem_u32(smem_b + stage * T2_B_BYTES), b_lbo, 1024);
#pragma unroll
          for (int k = 0; k < TG_BK / TG_UMMA_K; ++k)
            ptx::mma_f16_ss_2cta(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + b_kstep * (uint64_t)k, idesc, (kb | k) != 0);
          ptx::tc_commit_2cta(&empty_bar[stage], 0b11);   // stage free in both CTAs
          if (++stage == NSTAGES) { stage = 0; phase ^= 1; }
        }
        ptx::tc_commit_2cta(&acc_full[acc], 0b11);        // accumulator halves ready in both CTAs
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===== epilogue warps 2..13 (both CTAs): this CTA's 128 rows of the 256 x 256 tile =====
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;
    int acc = 0; uint32_t acc_phase = 0;
    if constexpr (!kResTma) {
      uint8_t* stg = staging + (warp - 2) * 4096;
      constexpr bool kTmaC = TMA_C && (EPI == TPAT_EPI_BIAS || EPI == TPAT_EPI_BIAS_GELU) && sizeof(OutT) == 2 && !FOLD;
      if constexpr (kTmaC) {
        // bf16 output through TMA stores (see tc_epilogue_tile_tma)
        {
          int kcount = 0;
          for (int tile = tile_first; tile < tile_end; tile += tile_step) {
            const int m0 = tc_tile_m(p, tile) * 256 + (int)rank * 128 + q * 32, n0 = (tile % p.tiles_n) * TG_BN;
            ptx::mbar_wait(&acc_full[acc], acc_phase);
            ptx::tc_fence_after();
            const uint32_t taddr_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * TG_BN;
            const uint32_t rel_leader = ptx::mapa_shared(ptx::smem_u32(&acc_empty[acc]), 0);
            tc_epilogue_tile_tma<EPI, T2_EPI_WARPS>(p, &tmap_c, taddr_row, m0, n0, cg, stg, lane, kcount,
                                                    [&]() { if (lane == 0) ptx::mbar_arrive_cluster(rel_leader); });
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
          }
          if (lane == 0) ptx::tma_store_wait<0>();         // all output blocks written before the CTA retires
        }
      } else {
      float2* rowstat = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(bars) + 512) + (warp - 2) * 32;
      int stat_mt = -1;                 // row block whose moments `rowstat` currently holds
      for (int tile = tile_first; tile < tile_end; tile += tile_step) {
        const int m0 = tc_tile_m(p, tile) * 256 + (int)rank * 128 + q * 32, n0 = (tile % p.tiles_n) * TG_BN;
        TcEpiPrefetch<T2_EPI_WARPS> pf;
        tc_epilogue_prefetch<T2_EPI_WARPS>(p, n0, cg, lane, pf);
        if (FOLD && EPI != TPAT_EPI_BIAS_RESIDUAL && tc_tile_m(p, tile) != stat_mt) {
          // LayerNorm fold: (mean, rstd) of this warp's 32 rows, combined while the MMA is still running
          stat_mt = tc_tile_m(p, tile);
          __syncwarp();
          rowstat[lane] = m0 + lane < p.M ? ln_row_moments(p.ln_part + (size_t)(m0 + lane) * p.ln_chunks, p.ln_chunks, p.ln_eps)
                                          : make_float2(0.f, 1.f);
          __syncwarp();
        }
        ptx::mbar_wait(&acc_full[acc], acc_phase);
        ptx::tc_fence_after();
        const uint32_t taddr_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * TG_BN;
        const uint32_t rel_leader = ptx::mapa_shared(ptx::smem_u32(&acc_empty[acc]), 0);
        tc_epilogue_tile<EPI, OutT, T2_EPI_WARPS, FOLD>(p, taddr_row, m0, n0, cg, stg, lane, pf, [&]() { if (lane == 0) ptx::mbar_arrive_cluster(rel_leader); },
                                                        rowstat);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      }
    } else {
      // Residual epilogue: C = R + acc + bias with fp32 R / C.  Each warp walks its (tile, chunk) items with two
      // 4 KB 128B-swizzled buffers: lane 0 TMA-loads the NEXT item's 32 x 32 residual block while the current one
      // is combined (thread = row: read own swizzled row, add the tcgen05.ld fragment + bias, write back) and
      // TMA-stored.  Loads do not depend on the MMA, so the first block of a tile is in flight before acc_full.
      constexpr int CS = T2_EPI_WARPS / 4, NCH = (8 + CS - 1) / CS;
      uint8_t* stg = staging + (warp - 2) * 8192;
      uint64_t* rb = res_bar + (warp - 2) * 2;
      int buf = 0; uint32_t rph0 = 0, rph1 = 0;
      auto item_cols = [&](int tile, int ci) { return (tile % p.tiles_n) * TG_BN + (cg + CS * ci) * 32; };
      auto item_row0 = [&](int tile) { return tc_tile_m(p, tile) * 256 + (int)rank * 128 + q * 32; };
      auto item_live = [&](int tile, int ci) {
        if (EPI == TPAT_EPI_BIAS_POS && item_row0(tile) >= p.M) return false;   // (M is a multiple of 32 here)
        return ci < NCH && cg + CS * ci < 8 && item_cols(tile, ci) < p.N;
      };
      // row of the block in the tensor that is read (residual: m0; pos table: extra + m0 % P) / written (shifted by the
      // extra-token rows of the clips before it)
      auto load_row = [&](int m0) { return EPI == TPAT_EPI_BIAS_POS ? p.num_extra + m0 % p.P : m0; };
      auto store_row = [&](int m0) { return EPI == TPAT_EPI_BIAS_POS ? m0 + (m0 / p.P + 1) * p.num_extra : m0; };
      // first live item at or after (tile, ci) in this warp's walk order; tile >= num_tiles when there is none
      auto next_live = [&](int& tile, int& ci) {
        while (tile < tile_end) {
          if (ci >= NCH) { tile += tile_step; ci = 0; continue; }
          if (item_live(tile, ci)) return;
          ++ci;
        }
      };
      const bool red = EPI == TPAT_EPI_BIAS_RESIDUAL && !FOLD && p.red_add != 0;     // grid-uniform
      auto issue_load = [&](int tile, int ci, int b) {   // lane 0 only
        if (red) return;
        const int m0 = tc_tile_m(p, tile) * 256 + (int)rank * 128 + q * 32;
        ptx::tma_store_wait_read<0>();                   // the store that last read this buffer has drained its smem reads
        ptx::mbar_arrive_expect_tx(&rb[b], 4096);
        ptx::tma_load_2d(stg + b * 4096, &tmap_r, &rb[b], item_cols(tile, ci), load_row(m0));
      };
      {
        int t0 = tile_first, c0 = 0;
        next_live(t0, c0);
        if (t0 < tile_end && lane == 0) issue_load(t0, c0, 0);
      }
      for (int tile = tile_first; tile < tile_end; tile += tile_step) {
        const int m0 = tc_tile_m(p, tile) * 256 + (int)rank * 128 + q * 32;
        const uint32_t taddr_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * TG_BN;
        const uint32_t rel_leader = ptx::mapa_shared(ptx::smem_u32(&acc_empty[acc]), 0);
        ptx::mbar_wait(&acc_full[acc], acc_phase);
        ptx::tc_fence_after();
        bool released = false;
#pragma unroll 1
        for (int ci = 0; ci < NCH; ++ci) {
          const bool live = item_live(tile, ci);
          uint32_t r[32];
          if (live) {
            ptx::tmem_ld_32x32b_x32(taddr_row + (cg + CS * ci) * 32, r);
            ptx::tmem_ld_wait();
          }
          bool later = false;                            // does this warp read more of this accumulator?
          for (int c2 = ci + 1; c2 < NCH; ++c2) later |= item_live(tile, c2);
          if (!released && !later) {
            released = true;
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(rel_leader);
          }
          if (!live) continue;
          if (lane == 0) {                               // prefetch the residual block of this warp's next item
            int nt = tile, nc = ci + 1;
            next_live(nt, nc);
            if (nt < tile_end) issue_load(nt, nc, buf ^ 1);
          }
          const int n = item_cols(tile, ci);
          if (red) {
            if (lane == 0) ptx::tma_store_wait_read<1>();      // the reduce that last read THIS buffer has drained it
            __syncwarp();
          } else {
            ptx::mbar_wait(&rb[buf], buf ? rph1 : rph0);
            if (buf) rph1 ^= 1; else rph0 ^= 1;
          }
          uint8_t* rowp = stg + buf * 4096 + lane * 128;
          // DropPath (training): the branch of clip b is scaled by row_scale[b] (0 or 1 / keep_prob)
          const float sc = (EPI == TPAT_EPI_BIAS_RESIDUAL && p.row_scale != nullptr) ? __ldg(p.row_scale + min(m0 + lane, p.M - 1) / p.rows_per_clip) : 1.0f;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4* cell = reinterpret_cast<float4*>(rowp + ((j ^ (lane & 7)) << 4));
            const float4 bb = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + n + 4 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 x = red ? make_float4(0.f, 0.f, 0.f, 0.f) : *cell;
            x.x = fmaf(sc, __uint_as_float(r[4 * j]) + bb.x, x.x); x.y = fmaf(sc, __uint_as_float(r[4 * j + 1]) + bb.y, x.y);
            x.z = fmaf(sc, __uint_as_float(r[4 * j + 2]) + bb.z, x.z); x.w = fmaf(sc, __uint_as_float(r[4 * j + 3]) + bb.w, x.w);
            *cell = x;
            r[4 * j] = __float_as_uint(x.x); r[4 * j + 1] = __float_as_uint(x.y);      // keep the row for the fold below
            r[4 * j + 2] = __float_as_uint(x.z); r[4 * j + 3] = __float_as_uint(x.w);
          }
          if constexpr (FOLD && EPI == TPAT_EPI_BIAS_RESIDUAL) {
            // LayerNorm fold, producer side: partial moments of this thread's 32 new values of row m0 + lane
            float sm = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) sm += __uint_as_float(r[i]);
            const float mc = sm * (1.0f / 32.0f);
            float q2 = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) { const float d = __uint_as_float(r[i]) - mc; q2 = fmaf(d, d, q2); }
            if (m0 + lane < p.M) p.part_out[(size_t)(m0 + lane) * p.part_ld + (n >> 5)] = make_float2(sm, q2);
          }
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (red) ptx::tma_reduce_add_2d(&tmap_c, stg + buf * 4096, n, store_row(m0));
            else ptx::tma_store_2d(&tmap_c, stg + buf * 4096, n, store_row(m0));   // rows >= M are clipped by the tensor map
            ptx::tma_store_commit();
          }
          if constexpr (FOLD && EPI == TPAT_EPI_BIAS_RESIDUAL) {
            // bf16 copy of the block: re-read the swizzled fp32 tile with lanes ALONG the rows (8 lanes = one 64 B row
            // segment, 4 rows per instruction) so that every global store fills whole 32 B sectors
            const int jl = lane & 7, rl = lane >> 3;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int row = it * 4 + rl;
              const float4 a = *reinterpret_cast<const float4*>(stg + buf * 4096 + row * 128 + ((jl ^ (row & 7)) << 4));
              if (m0 + row < p.M)
                *reinterpret_cast<uint2*>(p.xb + (size_t)(m0 + row) * p.ldxb + n + jl * 4) = make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w));
            }
          }
          buf ^= 1;
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (lane == 0) ptx::tma_store_wait<0>();           // all output blocks written before the CTA retires
    }
  }

  ptx::tc_fence_before();
  ptx::cluster_sync();          // neither CTA may exit (or free TMEM) while the peer can still touch it
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2cta<512>(tmem_base);
  }
}

template <int EPI, typename OutT, bool RES_TMA = false, bool FOLD = false, bool TMA_C = false>
static int launch_tc2(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& tr, const CUtensorMap& tc,
                      const TcGemmParams& p, cudaStream_t st) {
  static DeviceOnce once;
  auto kern = gemm_tc2_kernel<EPI, OutT, RES_TMA, FOLD, TMA_C>;
  constexpr int smem_bytes = (RES_TMA && (EPI == TPAT_EPI_BIAS_RESIDUAL || EPI == TPAT_EPI_BIAS_POS)) ? T2R_SMEM_BYTES : T2_SMEM_BYTES;
  if (once.first()) { TPAT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes)); once.mark(); }
  const int tiles = p.tiles_m * p.tiles_n;
  int clusters = sm_count() / 2;
  if (tiles < clusters) clusters = tiles;
  TPAT_CUDA(launch_kernel(kern, dim3(2 * clusters), dim3(T2_THREADS), smem_bytes, st, ta, tw, tr, tc, p));
  TPAT_LAUNCH_CHECK();
  return 0;
}

#ifndef TPAT_RES_REDUCE_DEFAULT
#define TPAT_RES_REDUCE_DEFAULT 0
#endif

int gemm_tc2(const void* A, int lda, const void* W, void* C, int c_dtype, int ldc, int M, int N, int K,
             const EpiParams& ep, cudaStream_t st) {
  CUtensorMap ta, tw;
  if (int rc = encode_tmap_2d(&ta, A, 2, (uint64_t)M, (uint64_t)K, (uint64_t)lda * 2, 128, TG_BK, true)) return rc;
  if (ep.w_kn) { if (int rc = encode_tmap_2d(&tw, W, 2, (uint64_t)K, (uint64_t)N, (uint64_t)N * 2, 64, 64, true)) return rc; }
  else if (int rc = encode_tmap_2d(&tw, W, 2, (uint64_t)N, (uint64_t)K, (uint64_t)K * 2, 128, TG_BK, true)) return rc;
  TcGemmParams p{};
  p.w_kn = ep.w_kn;
  p.cs_part = ep.cs_part;
  p.M = M; p.N = N; p.K = K; p.C = C; p.ldc = ldc; p.bias = ep.bias; p.residual = ep.residual; p.ldr = ep.ldr;
  p.pos = ep.pos; p.P = ep.P; p.num_extra = ep.num_extra;
  p.xb = (__nv_bfloat16*)ep.xb; p.ldxb = ep.ldxb; p.part_out = reinterpret_cast<float2*>(ep.part_out); p.part_ld = ep.part_ld;
  p.ln_part = reinterpret_cast<const float2*>(ep.ln_part); p.ln_chunks = ep.ln_chunks; p.ln_colsum = ep.ln_colsum; p.ln_eps = ep.ln_eps;
  p.dact_out = ep.dact_out; p.ld_dact = ep.ld_dact; p.aux = ep.aux; p.ld_aux = ep.ld_aux; p.row_scale = ep.row_scale; p.rows_per_clip = ep.rows_per_clip;
  p.tiles_m = (M + 255) / 256; p.tiles_n = (N + TG_BN - 1) / TG_BN;
  p.bn = TG_BN;
  p.desc = g_walk_desc;
#ifdef TPAT_DEBUG_BUILD
  { const char* e = getenv("TPAT_GEMM_DEBUG_SKIP"); p.debug_skip = e ? atoi(e) : 0; }   // timing experiments (wrong results!)
#endif
  // TPAT_GEMM_TMA_STORE=1 (OFF by default): bf16 outputs of the plain bias / bias + GELU epilogues (qkv, fc1, the bf16 data
  // gradients) leave through TMA stores instead of per-lane stores after a shared-memory transpose.  Bit-identical;
  // measured r02ac: qkv 0.0835 vs 0.0807 ms, fc1 + GELU 0.1111 vs 0.1111 ms at N = 513, whole forward within noise.
  CUtensorMap tc_out = ta;
  const char* tma_env = getenv("TPAT_GEMM_TMA_STORE");     // read per call so that tests can compare both epilogues
  const bool no_tma_c = tma_env == nullptr || tma_env[0] != '1';
  if (!no_tma_c && (ep.epilogue == TPAT_EPI_BIAS || ep.epilogue == TPAT_EPI_BIAS_GELU) && c_dtype == TPAT_BF16 && p.ln_part == nullptr &&
      p.dact_out == nullptr && N % 32 == 0 && (ldc * 2) % 16 == 0 && aligned16(C)) {
    if (int rc = encode_tmap_2d_c32(&tc_out, C, (uint64_t)M, (uint64_t)N, (uint64_t)ldc * 2)) return rc;
    p.tma_c = 1;
  }
  switch (ep.epilogue) {
    case TPAT_EPI_BIAS:
      if (p.ln_part != nullptr && c_dtype == TPAT_BF16) return launch_tc2<TPAT_EPI_BIAS, __nv_bfloat16, false, true>(ta, tw, ta, ta, p, st);
      TPAT_CHECK(p.ln_part == nullptr, "tpat_gemm_ln: the folded GEMM writes bf16");
      if (p.tma_c) return launch_tc2<TPAT_EPI_BIAS, __nv_bfloat16, false, false, true>(ta, tw, ta, tc_out, p, st);
      return c_dtype == TPAT_BF16 ? launch_tc2<TPAT_EPI_BIAS, __nv_bfloat16>(ta, tw, ta, ta, p, st) : launch_tc2<TPAT_EPI_BIAS, float>(ta, tw, ta, ta, p, st);
    case TPAT_EPI_BIAS_GELU:
      if (p.ln_part != nullptr && c_dtype == TPAT_BF16) return launch_tc2<TPAT_EPI_BIAS_GELU, __nv_bfloat16, false, true>(ta, tw, ta, ta, p, st);
      TPAT_CHECK(p.ln_part == nullptr, "tpat_gemm_ln: the folded GEMM writes bf16");
      if (p.tma_c) return launch_tc2<TPAT_EPI_BIAS_GELU, __nv_bfloat16, false, false, true>(ta, tw, ta, tc_out, p, st);
      return c_dtype == TPAT_BF16 ? launch_tc2<TPAT_EPI_BIAS_GELU, __nv_bfloat16>(ta, tw, ta, ta, p, st) : launch_tc2<TPAT_EPI_BIAS_GELU, float>(ta, tw, ta, ta, p, st);
    case TPAT_EPI_DGELU: {
      CUtensorMap tx = ta;
      static const bool no_pf = getenv("TPAT_GEMM_NO_L2_PREFETCH") != nullptr;
      if (!no_pf && N % 256 == 0 && (ep.ld_aux * dtype_size(c_dtype)) % 16 == 0) {
        if (int rc = encode_tmap_2d(&tx, ep.aux, (int)dtype_size(c_dtype), (uint64_t)M, (uint64_t)N, (uint64_t)ep.ld_aux * dtype_size(c_dtype), 128, 256, false)) return rc;
        p.pf_l2 = 1;
      }
      return c_dtype == TPAT_BF16 ? launch_tc2<TPAT_EPI_DGELU, __nv_bfloat16>(ta, tw, tx, ta, p, st) : launch_tc2<TPAT_EPI_DGELU, float>(ta, tw, tx, ta, p, st);
    }
    case TPAT_EPI_BIAS_RESIDUAL: {
      // Long-K GEMMs (fc2, K = 3072) hide the register-path epilogue behind the main loop and prefer the fifth
      // pipeline stage; short-K ones (proj, K = 768) are bound by the residual read-modify-write and use the
      // TMA-fed epilogue (measured r01: proj 0.081 -> 0.064 ms, fc2 0.129 -> 0.140 ms with it).
      // In place (C == R, the forward's residual stream) the read-modify-write can be left to L2: TMA reduce-add of
      // acc + bias blocks (TPAT_GEMM_RES_REDUCE = proj | all | 0; bit-identical: x + v is the same fp32 addition)
      static const int red_mode = [] { const char* e = getenv("TPAT_GEMM_RES_REDUCE"); return e == nullptr ? TPAT_RES_REDUCE_DEFAULT : (e[0] == 'a' ? 2 : (e[0] == 'p' ? 1 : 0)); }();
      const bool can_red = ep.residual == C && ep.ldr == ldc && p.xb == nullptr && ep.row_scale == nullptr;
      p.red_add = can_red && (red_mode == 2 || (red_mode == 1 && K <= 1536));
      if (K > 1536 && !p.red_add) {
        CUtensorMap tx = ta;
        static const bool pf_res = getenv("TPAT_GEMM_RES_L2_PREFETCH") != nullptr;     // (A/B switch)
        if (pf_res && N % 256 == 0) {
          if (int rc = encode_tmap_2d(&tx, ep.residual, 4, (uint64_t)M, (uint64_t)N, (uint64_t)ep.ldr * 4, 128, 256, false)) return rc;
          p.pf_l2 = 1;
        }
        return p.xb ? launch_tc2<TPAT_EPI_BIAS_RESIDUAL, float, false, true>(ta, tw, tx, ta, p, st)
                    : launch_tc2<TPAT_EPI_BIAS_RESIDUAL, float, false>(ta, tw, tx, ta, p, st);
      }
      // fp32 residual in / C out as 32 x 32 blocks (128 B rows, 128B swizzle)
      CUtensorMap tr, tc;
      if (int rc = encode_tmap_2d(&tr, ep.residual, 4, (uint64_t)M, (uint64_t)N, (uint64_t)ep.ldr * 4, 32, 32, true)) return rc;
      if (int rc = encode_tmap_2d(&tc, C, 4, (uint64_t)M, (uint64_t)N, (uint64_t)ldc * 4, 32, 32, true)) return rc;
      return p.xb ? launch_tc2<TPAT_EPI_BIAS_RESIDUAL, float, true, true>(ta, tw, tr, tc, p, st)
                  : launch_tc2<TPAT_EPI_BIAS_RESIDUAL, float, true>(ta, tw, tr, tc, p, st);
    }
    case TPAT_EPI_BIAS_POS: {
      // K = 256: the epilogue (fp32 rows out, position rows in) is the whole cost.  When a 32-row block never straddles
      // a clip (P % 32 == 0) it runs as a TMA read-modify-write like the residual epilogue: pos block in, C block out.
      const int rows_out = (M / ep.P) * (ep.P + ep.num_extra);
      if (ep.P % 32 != 0 || M % ep.P != 0) return launch_tc2<TPAT_EPI_BIAS_POS, float>(ta, tw, ta, ta, p, st);
      CUtensorMap tr, tc;
      if (int rc = encode_tmap_2d(&tr, ep.pos, 4, (uint64_t)(ep.P + ep.num_extra), (uint64_t)N, (uint64_t)ldc * 4, 32, 32, true)) return rc;
      if (int rc = encode_tmap_2d(&tc, C, 4, (uint64_t)rows_out, (uint64_t)N, (uint64_t)ldc * 4, 32, 32, true)) return rc;
      return launch_tc2<TPAT_EPI_BIAS_POS, float, true>(ta, tw, tr, tc, p, st);
    }
  }
  set_error("tpat_gemm(tc2): bad epilogue %d", ep.epilogue);
  return 1;
}

}  // namespace tpat
