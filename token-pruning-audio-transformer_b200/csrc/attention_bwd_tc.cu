// tcgen05 attention backward for sm_100a (bf16 operands, fp32 accumulate / softmax math).
//
// Backward of q k^T * scale -> softmax -> attn @ v (reference audiomae/models_vit.py:75-95 under autograd):
//     P = exp(scale S - lse),  dP = dO V^T,  dS = P o (dP - delta) scale,  dV = P^T dO,  dK = dS^T Q,  dQ = dS K
// with S recomputed from Q, K and the forward's log-sum-exp; nothing of size N x N reaches HBM.
//
// One CTA per (clip, head, 128-KEY tile j); it walks the 128-query tiles i.  dV_j and dK_j accumulate in tensor
// memory over the whole walk; the dQ_i contribution of this key tile is added into an fp32 accumulator in HBM with
// a TMA reduce (cp.reduce.async.bulk.tensor ... .add: no per-thread atomics; <= 5 key tiles add into each element).
//   TMEM (all 512 columns):  S [0,128)  dP [128,256)  dV [256,320)  dK [320,384)  dQ ping [384,448) pong [448,512)
//   smem (194 KB):  K_j, V_j (16 KB each, loaded once) | Q_i, dO_i (2 stages x 16 KB each) | P, dS (32 KB each, bf16,
//                   [query][key] in two 64-key column blocks) | dQ staging (8 x 4 KB)
//   warp 0   TMA producer;  warp 1   MMA issuer (one elected thread);  warps 2-9  softmax / epilogue:
//            thread = (query row = TMEM lane, 64-key half)
// Five products per (i, j), all tcgen05.mma kind::f16 from shared memory:
//   S  = Q_i K_j^T   (M128 N128 K64, both K-major)        dP = dO_i V_j^T   (same)
//   dV += P^T dO_i   (M128 N64 K128: A = P MN-major, two 64-key blocks 16 KB apart (LBO); B = dO_i MN-major)
//   dK += dS^T Q_i   (same shapes with dS, Q_i)           dQ = dS K_j      (M128 N64 K128: A = dS K-major, B = K_j MN-major)
// The same [query][key] bf16 tile therefore serves as an MN-major A operand (dV, dK) and a K-major one (dQ).
#include "attention.cuh"
#include "ptx_sm100.cuh"

#include <cstdlib>

namespace tpat {

int encode_tmap_3d(CUtensorMap* out, const void* gptr, int elem_bytes, int B, int N, int ld, int box_rows, int box_cols);

constexpr int BT_M = 128;                 // queries per tile = keys per tile
constexpr int BT_HD = 64;
constexpr int BT_TILE = BT_M * BT_HD * 2; // 16 KB: one [128][64] bf16 operand tile
constexpr int BT_PS = 2 * BT_TILE;        // 32 KB: P or dS, two 64-key column blocks of [128 queries][128 B]
// CW = softmax / epilogue warps: 8 (thread = query row x 64 keys; default) or 16 (x 32 keys: four warps per SM sub-partition,
// TPAT_ATTN_BWD_WARPS=16).  Measured equal (0.471 vs 0.474 ms at N = 513, profiles/r02y_attn_bwd_trace_16warps.txt): the warps
// move through their phases in lock-step (exponentials: MUFU-bound, 1 024 cycles per tile whatever the warp count; P / dS
// stores: shared-memory-bandwidth-bound, 512 cycles) because the single P / dS buffer is handed over by CTA-wide barriers, so
// more warps do not overlap one phase with another.
constexpr int bt_threads(int cw) { return 64 + 32 * cw; }
constexpr int bt_smem(int cw) { return 1024 + 2 * BT_TILE + 4 * BT_TILE + 2 * BT_PS + cw * 4096 + 256; }
constexpr uint32_t BT_S = 0, BT_DP = 128, BT_DV = 256, BT_DK = 320, BT_DQ = 384;   // dQ: two buffers, [384,448) and [448,512)

#ifdef TPAT_ATTN_BWD_TRACE
#define BWD_TRACE(slot) do { if (tracing && tn < 250) p.trace[tn++] = (clock64() & 0xFFFFFFFFFFFFll) | ((long long)(slot) << 48); } while (0)
#else
#define BWD_TRACE(slot) do { } while (0)
#endif

struct AttnBwdTcParams {
  long long* trace;      // debug builds (TPAT_ATTN_BWD_TRACE): clock stamps of one softmax thread and of the MMA thread
  const float* lse;      // [B, H, N] natural log
  const float* delta;    // [B, H, N]
  float* bias_part;      // optional [B * n_t * 2][3 * H * 64]: column sums of this CTA's dK / dV rows (two 64-row halves)
  int N, H, n_t;
  float scale, scale_log2;
};

// 32 values of one query row -> bf16, into the 128B-swizzled [row][64 keys] block: 16-byte pieces piece0 .. piece0 + 3
__device__ __forceinline__ void bt_store_row32(uint8_t* block, int r_local, int piece0, const float (&v)[32]) {
#pragma unroll
  for (int g = 0; g < 4; ++g)
    *reinterpret_cast<uint4*>(block + r_local * 128 + (((piece0 + g) ^ (r_local & 7)) * 16)) =
        make_uint4(pack_bf16x2(v[g * 8 + 0], v[g * 8 + 1]), pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]),
                   pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]), pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]));
}

template <int CW>
__global__ void __launch_bounds__(bt_threads(CW), 1)
attention_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                        const __grid_constant__ CUtensorMap tm_dq, const __grid_constant__ CUtensorMap tm_dkv,
                        const AttnBwdTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* k_s = smem;
  uint8_t* v_s = k_s + BT_TILE;
  uint8_t* q_s = v_s + BT_TILE;            // 2 stages
  uint8_t* do_s = q_s + 2 * BT_TILE;       // 2 stages
  uint8_t* p_s = do_s + 2 * BT_TILE;
  uint8_t* ds_s = p_s + BT_PS;
  uint8_t* stg = ds_s + BT_PS;             // CW x 4 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg + CW * 4096);
  constexpr int NPART = CW / 4;            // column parts of the 128-key tile per query row: 2 | 4
  constexpr int COLS = BT_M / NPART;       // keys per thread: 64 | 32
  constexpr int NCH = COLS / 32;           // 32-column TMEM chunks per thread
  uint64_t* kv_full = bars;                // [1]
  uint64_t* q_full = bars + 1;             // [2]
  uint64_t* qdo_empty = bars + 3;          // [2]
  uint64_t* s_full = bars + 5;             // [1]
  uint64_t* pds_full = bars + 6;           // [1] CW arrivals
  uint64_t* dq_full = bars + 7;            // [1]
  uint64_t* dkv_full = bars + 8;           // [1]
  uint64_t* sdp_free = bars + 9;           // [1] CW arrivals: S / dP of the current tile are in registers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int jt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int n_t = p.n_t;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_qkv); ptx::prefetch_tensormap(&tm_do);
    ptx::prefetch_tensormap(&tm_dq); ptx::prefetch_tensormap(&tm_dkv);
  }
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(&q_full[s], 1); ptx::mbar_init(&qdo_empty[s], 1); }
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(pds_full, CW);
    ptx::mbar_init(dq_full, 1);
    ptx::mbar_init(dkv_full, 1);
    ptx::mbar_init(sdp_free, CW);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc<512>(tmem_slot);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();

  const int col_q = h * BT_HD, col_k = (p.H + h) * BT_HD, col_v = (2 * p.H + h) * BT_HD;

  if (warp == 0) {
    // ===== TMA producer =====
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(kv_full, 2 * BT_TILE);
      ptx::tma_load_3d(k_s, &tm_qkv, kv_full, col_k, jt * BT_M, b);
      ptx::tma_load_3d(v_s, &tm_qkv, kv_full, col_v, jt * BT_M, b);
      for (int i = 0; i < n_t; ++i) {
        const int st = i & 1;
        ptx::mbar_wait(&qdo_empty[st], ((i >> 1) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&q_full[st], 2 * BT_TILE);
        ptx::tma_load_3d(q_s + st * BT_TILE, &tm_qkv, &q_full[st], col_q, i * BT_M, b);
        ptx::tma_load_3d(do_s + st * BT_TILE, &tm_do, &q_full[st], col_q, i * BT_M, b);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (ptx::elect_one()) {
#ifdef TPAT_ATTN_BWD_TRACE
      const bool tracing = p.trace != nullptr && blockIdx.x == 1 && blockIdx.y == 3 && blockIdx.z == (gridDim.z >> 1);
      int tn = 256;
#define BWD_TRACE_M(slot) do { if (tracing && tn < 500) p.trace[tn++] = (clock64() & 0xFFFFFFFFFFFFll) | ((long long)(slot) << 48); } while (0)
#else
#define BWD_TRACE_M(slot) do { } while (0)
#endif
      constexpr uint32_t idesc_s = ptx::idesc_bf16_f32(128, 128, 0, 0);    // Q / dO (K-major) x K / V (K-major)
      constexpr uint32_t idesc_kv = ptx::idesc_bf16_f32(128, 64, 1, 1);    // P^T / dS^T (MN-major) x dO / Q (MN-major)
      constexpr uint32_t idesc_q = ptx::idesc_bf16_f32(128, 64, 0, 1);     // dS (K-major) x K (MN-major)
      const uint64_t k_desc = ptx::smem_desc_sw128(ptx::smem_u32(k_s), 16, 1024);
      const uint64_t v_desc = ptx::smem_desc_sw128(ptx::smem_u32(v_s), 16, 1024);
      auto issue_sdp = [&](int i) {
        const int st = i & 1;
        ptx::mbar_wait(&q_full[st], (i >> 1) & 1);
        ptx::tc_fence_after();
        const uint64_t q_desc = ptx::smem_desc_sw128(ptx::smem_u32(q_s + st * BT_TILE), 16, 1024);
        const uint64_t do_desc = ptx::smem_desc_sw128(ptx::smem_u32(do_s + st * BT_TILE), 16, 1024);
#pragma unroll
        for (int k = 0; k < BT_HD / 16; ++k)
          ptx::mma_f16_ss(tmem + BT_S, q_desc + (uint64_t)(2 * k), k_desc + (uint64_t)(2 * k), idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < BT_HD / 16; ++k)
          ptx::mma_f16_ss(tmem + BT_DP, do_desc + (uint64_t)(2 * k), v_desc + (uint64_t)(2 * k), idesc_s, k != 0);
        ptx::tc_commit(s_full);
      };
      ptx::mbar_wait(kv_full, 0);
      issue_sdp(0);
      for (int i = 0; i < n_t; ++i) {
        const int st = i & 1;
        // S / dP of the NEXT query tile as soon as the softmax warps hold tile i's in registers (about half way through
        // their exponentials): the results are ready long before the softmax warps come back for them
        ptx::mbar_wait(sdp_free, i & 1);
        ptx::tc_fence_after();
        if (i + 1 < n_t) issue_sdp(i + 1);
        BWD_TRACE_M(20);
        ptx::mbar_wait(pds_full, i & 1);            // P and dS of tile i are in shared memory
        ptx::tc_fence_after();
        BWD_TRACE_M(21);
        const uint32_t p_a = ptx::smem_u32(p_s), ds_a = ptx::smem_u32(ds_s);
        const uint32_t do_a = ptx::smem_u32(do_s + st * BT_TILE), q_a = ptx::smem_u32(q_s + st * BT_TILE), k_a = ptx::smem_u32(k_s);
        // K index = query row: 16 rows = two 8-row groups of 1024 B; the second 64-key block of P / dS is 16 KB further (LBO)
#pragma unroll
        for (int ks = 0; ks < BT_M / 16; ++ks)
          ptx::mma_f16_ss(tmem + BT_DV, ptx::smem_desc_sw128(p_a + ks * 2048, BT_TILE, 1024),
                          ptx::smem_desc_sw128(do_a + ks * 2048, 16, 1024), idesc_kv, (i | ks) != 0);
#pragma unroll
        for (int ks = 0; ks < BT_M / 16; ++ks)
          ptx::mma_f16_ss(tmem + BT_DK, ptx::smem_desc_sw128(ds_a + ks * 2048, BT_TILE, 1024),
                          ptx::smem_desc_sw128(q_a + ks * 2048, 16, 1024), idesc_kv, (i | ks) != 0);
        // dQ = dS K_j: K index = key: dS K-major (64-key block ks / 4, 32 B per step), K_j MN-major (2048 B per 16 keys)
#pragma unroll
        for (int ks = 0; ks < BT_M / 16; ++ks)
          ptx::mma_f16_ss(tmem + BT_DQ + (uint32_t)(i & 1) * 64, ptx::smem_desc_sw128(ds_a + (ks >> 2) * BT_TILE + (ks & 3) * 32, 16, 1024),
                          ptx::smem_desc_sw128(k_a + ks * 2048, 16, 1024), idesc_q, ks != 0);
        ptx::tc_commit(&qdo_empty[st]);
        ptx::tc_commit(dq_full);                    // also: P / dS shared memory is free again
        BWD_TRACE_M(22);
      }
#ifdef TPAT_ATTN_BWD_TRACE
      if (tracing) p.trace[511] = tn;
#endif
      ptx::tc_commit(dkv_full);
    }
  } else {
    // ===== softmax / epilogue warps =====
    const int quarter = warp & 3;
    const int part = (warp - 2) >> 2;      // which COLS keys of the tile; half = which 32 columns in the dQ / dK / dV epilogues
    const int half = part & 1;
    const int r_local = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const float* lse_bh = p.lse + ((size_t)b * p.H + h) * p.N;
    const float* delta_bh = p.delta + ((size_t)b * p.H + h) * p.N;
    uint8_t* my_stg = stg + (warp - 2) * 4096;
    const float c = p.scale_log2;
    constexpr float LOG2E = 1.4426950408889634f;
    // Software pipeline over the query tiles: while the tensor core runs the 24 MMAs of tile i, these warps already
    // turn S / dP of tile i + 1 into P / dS (kept in registers as bf16 pairs); the stores into the shared P / dS tile wait
    // for dq_full(i) (= those MMAs have retired), and the dQ_i epilogue runs after the hand-off, next to MMAs(i + 1).
#ifdef TPAT_ATTN_BWD_TRACE
    const bool tracing = p.trace != nullptr && blockIdx.x == 1 && blockIdx.y == 3 && blockIdx.z == (gridDim.z >> 1) && threadIdx.x == 64;
    int tn = 0;
    BWD_TRACE(0);
#endif
    uint32_t pk_p[COLS / 2], pk_ds[COLS / 2];      // this thread's keys of P and dS, packed bf16 pairs
    // log-sum-exp and delta of this thread's row of tile i are requested one tile ahead (their global-load latency was
    // ~850 cycles per tile pair on the critical path, profiles/r02m_attn_bwd_trace_v3.txt)
    float lse_nx = 0.f, dlt_nx = 0.f;
    auto prefetch_row = [&](int i) {
      const int row = i * BT_M + r_local;
      const bool ok = i < n_t && row < p.N;
      lse_nx = ok ? __ldg(lse_bh + row) : 0.f;
      dlt_nx = ok ? __ldg(delta_bh + row) : 0.f;
    };
    auto compute = [&](int i, float lse_row, float dlt) {
      const int row = i * BT_M + r_local;
      const bool row_ok = row < p.N;
      BWD_TRACE(10);
      const float lse2 = lse_row * LOG2E;
#ifdef TPAT_ATTN_BWD_TRACE
      asm volatile("" :: "f"(lse2), "f"(dlt));      // the scoreboard wait for the prefetched row values lands here
#endif
      BWD_TRACE(1);
      ptx::mbar_wait(s_full, i & 1);
      ptx::tc_fence_after();
      BWD_TRACE(2);
      // dS = P (dP - delta) scale = P * fma(dP, scale, -delta * scale).  Fully valid tile pairs (the common case) skip the
      // per-element bounds tests.
      const float nds = -dlt * p.scale;
      const bool all_ok = (i * BT_M + BT_M <= p.N) && (jt * BT_M + BT_M <= p.N);     // CTA-uniform
#pragma unroll
      for (int cc = 0; cc < NCH; ++cc) {
        const int col0 = part * COLS + cc * 32;             // column of the 128-key tile
        uint32_t rs[32], rd[32];
        ptx::tmem_ld_32x32b_x32(tmem + lane_off + BT_S + col0, rs);
        ptx::tmem_ld_32x32b_x32(tmem + lane_off + BT_DP + col0, rd);
        ptx::tmem_ld_wait();
        if (cc == NCH - 1) {                 // all chunks are out of tensor memory: S / dP may be overwritten
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(sdp_free);
        }
        if (all_ok) {
#pragma unroll
          for (int t = 0; t < 32; t += 2) {
            const float p0 = ptx::ex2_ftz(fmaf(__uint_as_float(rs[t]), c, -lse2));
            const float p1 = ptx::ex2_ftz(fmaf(__uint_as_float(rs[t + 1]), c, -lse2));
            pk_p[cc * 16 + (t >> 1)] = pack_bf16x2(p0, p1);
            pk_ds[cc * 16 + (t >> 1)] = pack_bf16x2(p0 * fmaf(__uint_as_float(rd[t]), p.scale, nds), p1 * fmaf(__uint_as_float(rd[t + 1]), p.scale, nds));
          }
        } else {
          const int key0 = jt * BT_M + col0;
#pragma unroll
          for (int t = 0; t < 32; t += 2) {
            const bool ok0 = row_ok && (key0 + t < p.N), ok1 = row_ok && (key0 + t + 1 < p.N);
            const float p0 = ok0 ? ptx::ex2_ftz(fmaf(__uint_as_float(rs[t]), c, -lse2)) : 0.f;
            const float p1 = ok1 ? ptx::ex2_ftz(fmaf(__uint_as_float(rs[t + 1]), c, -lse2)) : 0.f;
            pk_p[cc * 16 + (t >> 1)] = pack_bf16x2(p0, p1);
            pk_ds[cc * 16 + (t >> 1)] = pack_bf16x2(p0 * fmaf(__uint_as_float(rd[t]), p.scale, nds), p1 * fmaf(__uint_as_float(rd[t + 1]), p.scale, nds));
          }
        }
      }
      ptx::tc_fence_before();              // the TMEM reads above are complete
      BWD_TRACE(3);
    };
    auto store_pds = [&]() {
      // the tile is two 64-key column blocks of [128 rows][128 B]; this thread owns COLS / 8 16-byte pieces of one of them
      const int blk = (part * COLS) >> 6, pc0 = ((part * COLS) & 63) >> 3;
      uint8_t* prow = p_s + blk * BT_TILE + r_local * 128;
      uint8_t* drow = ds_s + blk * BT_TILE + r_local * 128;
#pragma unroll
      for (int g = 0; g < COLS / 8; ++g) {
        const int sw = ((pc0 + g) ^ (r_local & 7)) * 16;
        *reinterpret_cast<uint4*>(prow + sw) = make_uint4(pk_p[4 * g], pk_p[4 * g + 1], pk_p[4 * g + 2], pk_p[4 * g + 3]);
        *reinterpret_cast<uint4*>(drow + sw) = make_uint4(pk_ds[4 * g], pk_ds[4 * g + 1], pk_ds[4 * g + 2], pk_ds[4 * g + 3]);
      }
    };
    auto dq_epilogue = [&](int i) {        // dQ_i contribution of this key tile: TMEM -> swizzled fp32 staging -> TMA reduce-add
      uint32_t r[32];
      ptx::tmem_ld_32x32b_x32(tmem + lane_off + BT_DQ + (uint32_t)(i & 1) * 64 + half * 32, r);
      ptx::tmem_ld_wait();
      if (lane == 0) ptx::tma_store_wait_read<0>();          // the previous reduce has drained this staging buffer
      __syncwarp();
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4)
        *reinterpret_cast<uint4*>(my_stg + lane * 128 + ((j4 ^ (lane & 7)) << 4)) = make_uint4(r[4 * j4], r[4 * j4 + 1], r[4 * j4 + 2], r[4 * j4 + 3]);
      ptx::tc_fence_before();
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0 && i * BT_M + quarter * 32 < p.N) {
        ptx::tma_reduce_add_3d(&tm_dq, my_stg, col_q + half * 32, i * BT_M + quarter * 32, b);   // rows >= N are dropped
        ptx::tma_store_commit();
      }
    };
    prefetch_row(0);
    compute(0, lse_nx, dlt_nx);
    prefetch_row(1);
    for (int i = 0; i < n_t; ++i) {
      const float lse_i1 = lse_nx, dlt_i1 = dlt_nx;      // tile i + 1's values (requested a whole iteration ago)
      prefetch_row(i + 2);
      BWD_TRACE(9);
      if (i > 0) {                         // MMAs(i - 1) have retired: P / dS shared memory is free, dQ(i - 1) is complete
        BWD_TRACE(4);
        ptx::mbar_wait(dq_full, (i - 1) & 1);
        ptx::tc_fence_after();
        BWD_TRACE(5);
      }
      store_pds();
      ptx::fence_proxy_async_smem();       // P / dS visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(pds_full);
      BWD_TRACE(6);
      // (16 warps: the two warp groups take the dQ epilogue of alternate tiles)
      if (i > 0 && (NPART == 2 || (part >> 1) == ((i - 1) & 1))) dq_epilogue(i - 1);
      BWD_TRACE(7);
      if (i + 1 < n_t) compute(i + 1, lse_i1, dlt_i1);
    }
    ptx::mbar_wait(dq_full, (n_t - 1) & 1);
    ptx::tc_fence_after();
    if (NPART == 2 || (part >> 1) == ((n_t - 1) & 1)) dq_epilogue(n_t - 1);
    BWD_TRACE(8);
#ifdef TPAT_ATTN_BWD_TRACE
    if (tracing) p.trace[254] = tn;
#endif
    // ---- dK_j, dV_j: TMEM -> bf16 -> the (dead) K / V tiles -> two TMA stores ----
    if (part < 2) {                        // (16 warps: the first eight finish the CTA; thread = key row x 32 of the 64 columns)
    ptx::mbar_wait(dkv_full, 0);
    ptx::tc_fence_after();
    {
      uint32_t r[32];
      float v[32];
      ptx::tmem_ld_32x32b_x32(tmem + lane_off + BT_DK + half * 32, r);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int t = 0; t < 32; ++t) v[t] = __uint_as_float(r[t]);
      bt_store_row32(k_s, r_local, half * 4, v);
      ptx::tmem_ld_32x32b_x32(tmem + lane_off + BT_DV + half * 32, r);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int t = 0; t < 32; ++t) v[t] = __uint_as_float(r[t]);
      bt_store_row32(v_s, r_local, half * 4, v);
    }
    ptx::fence_proxy_async_smem();
    asm volatile("bar.sync 1, 256;\n" ::: "memory");
    if (warp == 2 && lane == 0) {
      ptx::tma_store_3d(&tm_dkv, k_s, col_k, jt * BT_M, b);      // key rows >= N are clipped by the tensor map
      ptx::tma_store_3d(&tm_dkv, v_s, col_v, jt * BT_M, b);
      ptx::tma_store_commit();
    }
    if (p.bias_part != nullptr) {
      // bias gradient of the qkv projection, K / V columns: column sums of the bf16 tiles just staged (valid key rows
      // only), thread = (row half, tensor, column); a warp reads 64 contiguous bytes of one swizzled row
      const int t = threadIdx.x - 64;
      const int c = t & 63, tsel = (t >> 6) & 1, rh = t >> 7;
      const uint8_t* blk = tsel ? v_s : k_s;
      const int r_end = min(rh * 64 + 64, min(BT_M, p.N - jt * BT_M));
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      auto at = [&](int r) {
        const uint16_t u = *reinterpret_cast<const uint16_t*>(blk + r * 128 + ((((c >> 3) ^ (r & 7)) << 4) | ((c & 7) << 1)));
        return __uint_as_float((uint32_t)u << 16);
      };
      int r = rh * 64;
      for (; r + 3 < r_end; r += 4) { s0 += at(r); s1 += at(r + 1); s2 += at(r + 2); s3 += at(r + 3); }
      for (; r < r_end; ++r) s0 += at(r);
      p.bias_part[((size_t)(b * p.n_t + jt) * 2 + rh) * (size_t)(3 * p.H * BT_HD) + (tsel ? col_v : col_k) + c] = (s0 + s1) + (s2 + s3);
    }
    }
    if (lane == 0) ptx::tma_store_wait<0>();                    // reduces / stores complete before the CTA retires
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem);
  }
}

// dq accumulator fp32 [B * N, H * 64] -> bf16 into the q columns of dqkv [B * N, 3 * H * 64]
__global__ void __launch_bounds__(256)
dq_convert_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dqkv, size_t rows, int HD) {
  pdl_trigger();
  pdl_wait();
  const int c4n = HD / 4;
  const size_t total = rows * c4n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / c4n; const int c4 = (int)(i - r * c4n);
    const float4 v = __ldg(reinterpret_cast<const float4*>(acc) + i);
    reinterpret_cast<uint2*>(dqkv + r * 3 * HD)[c4] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}

// the same per (clip, 64-row half of a 128-row tile) with the column sums of the fp32 rows (q columns of the qkv bias
// gradient): thread = (16-byte column group, row phase), the two phases combined through shared memory; partial rows as
// attention_bwd_tc_kernel writes them (one per tile half)
__global__ void __launch_bounds__(1024)
dq_convert_sum_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dqkv, float* __restrict__ part, int N, int HD, int n_t) {
  __shared__ float4 red[512];
  pdl_trigger();
  pdl_wait();
  const int jt = blockIdx.x, b = blockIdx.y, rh = blockIdx.z;
  const int c4n = HD / 4;
  const int c4 = threadIdx.x % c4n, ph = threadIdx.x / c4n;
  const int r1 = min(N, jt * 128 + rh * 64 + 64);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  int r = jt * 128 + rh * 64 + ph;
  auto one = [&](int rr, const float4& v) {
    reinterpret_cast<uint2*>(dqkv + ((size_t)b * N + rr) * 3 * HD)[c4] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  };
  for (; r + 6 < r1; r += 8) {
    float4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = __ldg(reinterpret_cast<const float4*>(acc + ((size_t)b * N + r + 2 * k) * HD) + c4);
#pragma unroll
    for (int k = 0; k < 4; ++k) one(r + 2 * k, v[k]);
  }
  for (; r < r1; r += 2) one(r, __ldg(reinterpret_cast<const float4*>(acc + ((size_t)b * N + r) * HD) + c4));
  if (ph == 1) red[c4] = s;
  __syncthreads();
  if (ph == 0) {
    const float4 o = red[c4];
    reinterpret_cast<float4*>(part + ((size_t)(b * n_t + jt) * 2 + rh) * (size_t)(3 * HD))[c4] = make_float4(s.x + o.x, s.y + o.y, s.z + o.z, s.w + o.w);
  }
}

int attention_bwd_delta_bf16(const void* out, const void* d_out, float* delta, int B, int N, int H, cudaStream_t st);

int attention_bwd_tc(const void* qkv, const void* out, const void* d_out, const float* lse, void* dqkv, int B, int N, int H,
                     float scale, float* delta_ws, float* dbias, cudaStream_t st) {
  static const bool force_simt = getenv("TPAT_ATTN_BWD_SIMT") != nullptr;
  if (force_simt) {
    TPAT_CHECK(dbias == nullptr, "tpat_attention_bwd: dbias needs the tcgen05 kernel (TPAT_ATTN_BWD_SIMT is set)");
    return attention_bwd_simt(qkv, out, d_out, lse, dqkv, TPAT_BF16, B, N, H, scale, delta_ws, st);
  }
  // workspace: delta [B * H * N] then the fp32 dQ accumulator [B * N, H * 64] (256-byte aligned)
  float* delta = delta_ws;
  const size_t dq_off = ((size_t)B * H * N + 63) / 64 * 64;
  float* dq_acc = delta_ws + dq_off;
  if (int rc = attention_bwd_delta_bf16(out, d_out, delta, B, N, H, st)) return rc;
  TPAT_CUDA(cudaMemsetAsync(dq_acc, 0, (size_t)B * N * H * BT_HD * sizeof(float), st));
  CUtensorMap tm_qkv, tm_do, tm_dq, tm_dkv;
  if (int rc = encode_tmap_3d(&tm_qkv, qkv, 2, B, N, 3 * H * BT_HD, BT_M, 64)) return rc;
  if (int rc = encode_tmap_3d(&tm_do, d_out, 2, B, N, H * BT_HD, BT_M, 64)) return rc;
  if (int rc = encode_tmap_3d(&tm_dq, dq_acc, 4, B, N, H * BT_HD, 32, 32)) return rc;
  if (int rc = encode_tmap_3d(&tm_dkv, dqkv, 2, B, N, 3 * H * BT_HD, BT_M, 64)) return rc;
  AttnBwdTcParams p;
  p.trace = nullptr;
#ifdef TPAT_ATTN_BWD_TRACE
  { static long long* dbg = nullptr; if (!dbg) cudaMalloc(&dbg, 512 * sizeof(long long)); cudaMemsetAsync(dbg, 0, 512 * sizeof(long long), st);
    p.trace = dbg; extern long long* g_attn_bwd_trace_buf; g_attn_bwd_trace_buf = dbg; }
#endif
  p.lse = lse; p.delta = delta; p.N = N; p.H = H; p.n_t = (N + BT_M - 1) / BT_M;
  // (column-sum partials behind the dQ accumulator; every entry is written by this launch pair, no memset)
  p.bias_part = dbias != nullptr ? dq_acc + (size_t)B * N * H * BT_HD : nullptr;
  p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  static const int cw = [] { const char* e = getenv("TPAT_ATTN_BWD_WARPS"); return e != nullptr && atoi(e) == 16 ? 16 : 8; }();
  if (cw == 8) {
    static DeviceOnce once;
    if (once.first()) {
      TPAT_CUDA(cudaFuncSetAttribute(attention_bwd_tc_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, bt_smem(8)));
      once.mark();
    }
    TPAT_CUDA(launch_kernel(attention_bwd_tc_kernel<8>, dim3(p.n_t, H, B), dim3(bt_threads(8)), (size_t)bt_smem(8), st, tm_qkv, tm_do, tm_dq, tm_dkv, p));
  } else {
    static DeviceOnce once;
    if (once.first()) {
      TPAT_CUDA(cudaFuncSetAttribute(attention_bwd_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, bt_smem(16)));
      once.mark();
    }
    TPAT_CUDA(launch_kernel(attention_bwd_tc_kernel<16>, dim3(p.n_t, H, B), dim3(bt_threads(16)), (size_t)bt_smem(16), st, tm_qkv, tm_do, tm_dq, tm_dkv, p));
  }
  const size_t rows = (size_t)B * N;
  const size_t total = rows * (H * BT_HD / 4);
  const int grid = (int)((total + 255) / 256 < (size_t)sm_count() * 16 ? (total + 255) / 256 : (size_t)sm_count() * 16);
  if (dbias != nullptr) {
    TPAT_CHECK(H * BT_HD / 4 * 2 <= 1024, "tpat_attention_bwd: dbias supports up to 32 heads");
    TPAT_CUDA(launch_kernel(dq_convert_sum_kernel, dim3(p.n_t, B, 2), dim3(H * BT_HD / 4 * 2), 0, st, (const float*)dq_acc, (__nv_bfloat16*)dqkv,
                            p.bias_part, N, H * BT_HD, p.n_t));
    TPAT_LAUNCH_CHECK();
    return finish_colsum_partials(p.bias_part, B * p.n_t * 2, 3 * H * BT_HD, dbias, st);
  }
  TPAT_CUDA(launch_kernel(dq_convert_kernel, dim3(grid), dim3(256), 0, st, (const float*)dq_acc, (__nv_bfloat16*)dqkv, rows, H * BT_HD));
  TPAT_LAUNCH_CHECK();
  return 0;
}

}  // namespace tpat

#ifdef TPAT_ATTN_BWD_TRACE
namespace tpat { long long* g_attn_bwd_trace_buf = nullptr; }
extern "C" int tpat_debug_attn_bwd_trace(long long* host_out) {   // debug builds only: copy the 512 stamps to the host
  if (!tpat::g_attn_bwd_trace_buf) return 1;
  return cudaMemcpy(host_out, tpat::g_attn_bwd_trace_buf, 512 * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : 2;
}
#endif

extern "C" size_t tpat_attention_bwd_ws_floats(int B, int N, int H, int hd) {
  return ((size_t)B * H * N + 63) / 64 * 64 + (size_t)B * N * H * hd + (size_t)B * ((N + 127) / 128) * 2 * 3 * H * hd;
}
