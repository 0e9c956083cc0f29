// tcgen05 fused attention, third structure ("two tiles, one thread per row"): the tiles that do not feed an importance
// score (9 of 12 blocks in inference, every block's attention in training).
//
// Same math as attention_tc.cu's single-pass instantiation (q k^T * scale, lazily rescaled online softmax, attn @ v;
// reference audiomae/models_vit.py:79-95) -- different schedule, after the r01 / r02 traces: per
// 64-key block the old kernel spends ~2000 cycles of which only 600-800 are exponentials; the rest is fixed latency
// (barrier round trips, tcgen05.ld, the max / flag exchange between the two partner warps of a row).  Here
//   * ONE CTA per SM owns TWO 128-query tiles of one (clip, head) and all 512 TMEM columns:
//         S_A [0,128)  S_B [128,256)  O_A [256,320)  O_B [320,384)  P_A [384,448)  P_B [448,512)
//     K / V tiles are loaded once for both query tiles (half the L2 -> shared traffic, one prologue per two tiles);
//   * key blocks are 128 wide (M128 N128 K64 for S; M128 N64 K128 for P.V with P read from tensor memory): half as many
//     barrier round trips per key;
//   * one thread owns one full query row (4 warps = 128 rows per tile): no partner warp, no shared-memory exchange, no
//     named barrier inside the loop; the lazy-rescale decision is thread-local (warp-uniform vote only because
//     tcgen05.ld / st are warp-collective);
//   * the two tiles' softmax warp groups ping-pong: while group A exponentiates S_A(j), the tensor core runs P_B.V /
//     S_B(j+1), and vice versa, so MUFU.EX2 -- the binding unit (16 / clk / SM) -- always has a group feeding it.
// warp 0: TMA producer (Q_A, Q_B once; K_j, V_j into a 4-slot ring of 16 KB tiles).  warp 1: MMA issuer.
// warps 2-5: softmax group A, warps 6-9: softmax group B (TMEM lane quarter = warp % 4).
#include "attention.cuh"
#include "ptx_sm100.cuh"

#include <cstdlib>

namespace tpat {

int encode_tmap_3d(CUtensorMap* out, const void* gptr, int elem_bytes, int B, int N, int ld, int box_rows, int box_cols);

constexpr int A3_BM = 128, A3_BK = 128, A3_HD = 64;
constexpr int A3_TILE = 128 * 64 * 2;      // 16 KB: Q tile, K tile, V tile
constexpr int A3_SLOTS = 4;
constexpr int A3_THREADS = 320;
constexpr int A3_SMEM = 1024 + 2 * A3_TILE + A3_SLOTS * A3_TILE + 512;
constexpr float A3_RESCALE_LOG2 = 64.0f;

struct Attn3Params {
  float* lse;            // optional [B, H, N] natural-log sum of exp (training)
  int N, H, n_qt, nb, qt_offset;
  int desc;
  float scale_log2;
};

__device__ __forceinline__ void a3_store_row32(uint8_t* tile_row, int hf, int r_local, const float (&v)[32]) {
#pragma unroll
  for (int g = 0; g < 4; ++g)
    *reinterpret_cast<uint4*>(tile_row + (((hf * 4 + g) ^ (r_local & 7)) * 16)) =
        make_uint4(pack_bf16x2(v[g * 8 + 0], v[g * 8 + 1]), pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]),
                   pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]), pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]));
}

__global__ void __launch_bounds__(A3_THREADS, 1)
attention_tc3_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_o, const Attn3Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* q_s = smem;                           // 2 x 16 KB (later: O staging)
  uint8_t* kv_s = q_s + 2 * A3_TILE;             // A3_SLOTS x 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(kv_s + A3_SLOTS * A3_TILE);
  uint64_t* q_full = bars;                       // [1]
  uint64_t* kv_full = bars + 1;                  // [SLOTS]
  uint64_t* kv_empty = kv_full + A3_SLOTS;       // [SLOTS]
  uint64_t* s_full = kv_empty + A3_SLOTS;        // [2]
  uint64_t* s_empty = s_full + 2;                // [2]  4 arrivals (the tile's softmax warps)
  uint64_t* p_full = s_empty + 2;                // [2]  4 arrivals
  uint64_t* o_full = p_full + 2;                 // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y;
  const int b = p.desc ? (int)(gridDim.z - 1 - blockIdx.z) : (int)blockIdx.z;
  const int qt0 = p.qt_offset + 2 * blockIdx.x;              // tiles qt0 (A) and qt0 + 1 (B)
  const bool has_b = qt0 + 1 < p.n_qt;
  const int nb = p.nb;

  if (warp == 0 && lane == 0) { ptx::prefetch_tensormap(&tm_qkv); ptx::prefetch_tensormap(&tm_o); }
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(q_full, 1);
    for (int s = 0; s < A3_SLOTS; ++s) { ptx::mbar_init(&kv_full[s], 1); ptx::mbar_init(&kv_empty[s], 1); }
    for (int x = 0; x < 2; ++x) {
      ptx::mbar_init(&s_full[x], 1); ptx::mbar_init(&s_empty[x], 4);
      ptx::mbar_init(&p_full[x], 4); ptx::mbar_init(&o_full[x], 1);
    }
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

  const int col_q = h * A3_HD, col_k = (p.H + h) * A3_HD, col_v = (2 * p.H + h) * A3_HD;

  if (warp == 0) {
    // ===== TMA producer =====
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(q_full, has_b ? 2 * A3_TILE : A3_TILE);
      ptx::tma_load_3d(q_s, &tm_qkv, q_full, col_q, qt0 * A3_BM, b);
      if (has_b) ptx::tma_load_3d(q_s + A3_TILE, &tm_qkv, q_full, col_q, (qt0 + 1) * A3_BM, b);
      int slot = 0; uint32_t phase = 0;
      for (int j = 0; j < nb; ++j) {
#pragma unroll 1
        for (int kv = 0; kv < 2; ++kv) {          // K_j then V_j: the order the MMA warp first touches them
          ptx::mbar_wait(&kv_empty[slot], phase ^ 1);
          ptx::mbar_arrive_expect_tx(&kv_full[slot], A3_TILE);
          ptx::tma_load_3d(kv_s + slot * A3_TILE, &tm_qkv, &kv_full[slot], kv ? col_v : col_k, j * A3_BK, b);
          if (++slot == A3_SLOTS) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (ptx::elect_one()) {
      constexpr uint32_t idesc_s = ptx::idesc_bf16_f32(128, 128, 0, 0);   // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_o = ptx::idesc_bf16_f32(128, A3_HD, 0, 1); // P (TMEM) x V (MN-major)
      auto slot_of = [](int idx) { return idx % A3_SLOTS; };
      auto phase_of = [](int idx) { return (uint32_t)((idx / A3_SLOTS) & 1); };
      auto issue_s = [&](int x, int j) {
        const uint64_t qd = ptx::smem_desc_sw128(ptx::smem_u32(q_s + x * A3_TILE), 16, 1024);
        const uint64_t kd = ptx::smem_desc_sw128(ptx::smem_u32(kv_s + slot_of(2 * j) * A3_TILE), 16, 1024);
#pragma unroll
        for (int k = 0; k < A3_HD / 16; ++k)
          ptx::mma_f16_ss(tmem + x * 128, qd + (uint64_t)(2 * k), kd + (uint64_t)(2 * k), idesc_s, k != 0);
        ptx::tc_commit(&s_full[x]);
      };
      auto issue_pv = [&](int x, int j) {
        const uint32_t v_addr = ptx::smem_u32(kv_s + slot_of(2 * j + 1) * A3_TILE);
        const int valid = min(A3_BK, p.N - j * A3_BK);
        const int ksteps = (valid + 15) >> 4;             // P is zero beyond `valid`; V rows beyond N are zero-filled
        for (int k = 0; k < ksteps; ++k)
          ptx::mma_f16_ts(tmem + 256 + x * 64, tmem + 384 + x * 64 + k * 8, ptx::smem_desc_sw128(v_addr + k * 2048, 16, 1024),
                          idesc_o, (j | k) != 0);
      };
      ptx::mbar_wait(q_full, 0);
      ptx::mbar_wait(&kv_full[slot_of(0)], phase_of(0));
      ptx::tc_fence_after();
      issue_s(0, 0);
      if (has_b) issue_s(1, 0);
      ptx::tc_commit(&kv_empty[slot_of(0)]);
      for (int j = 0; j < nb; ++j) {
        const int iv = 2 * j + 1, ik = 2 * j + 2;
        ptx::mbar_wait(&kv_full[slot_of(iv)], phase_of(iv));
        if (j + 1 < nb) ptx::mbar_wait(&kv_full[slot_of(ik)], phase_of(ik));
        // tile A: P_A(j) . V_j, then S_A(j + 1)
        ptx::mbar_wait(&p_full[0], j & 1);
        ptx::tc_fence_after();
        issue_pv(0, j);
        if (j + 1 < nb) {
          ptx::mbar_wait(&s_empty[0], j & 1);
          ptx::tc_fence_after();
          issue_s(0, j + 1);                              // (its commit also covers P_A(j) . V_j: P_A / O_A are free when s_full fires)
        }
        if (has_b) {
          ptx::mbar_wait(&p_full[1], j & 1);
          ptx::tc_fence_after();
          issue_pv(1, j);
        }
        ptx::tc_commit(&kv_empty[slot_of(iv)]);           // V_j consumed by both tiles
        if (j + 1 < nb) {
          if (has_b) {
            ptx::mbar_wait(&s_empty[1], j & 1);
            ptx::tc_fence_after();
            issue_s(1, j + 1);
          }
          ptx::tc_commit(&kv_empty[slot_of(ik)]);         // K_{j+1} consumed by both tiles
        }
      }
      ptx::tc_commit(&o_full[0]);
      if (has_b) ptx::tc_commit(&o_full[1]);
    }
  } else {
    // ===== softmax groups: warps 2-5 tile A, warps 6-9 tile B; thread = one query row =====
    const int x = (warp - 2) >> 2;
    if (x == 0 || has_b) {
      const int quarter = warp & 3;
      const int r_local = quarter * 32 + lane;
      const int q0 = (qt0 + x) * A3_BM;
      const int row = q0 + r_local;
      const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
      const uint32_t t_s = tmem + lane_off + x * 128, t_o = tmem + lane_off + 256 + x * 64, t_p = tmem + lane_off + 384 + x * 64;
      const float c = p.scale_log2;
      const bool warp_live = q0 + quarter * 32 < p.N;     // else: only keeps the barrier protocol going
      float m_ref = -INFINITY, l = 0.f;

      auto mask_chunk = [&](int j, int ch, uint32_t (&r)[32]) {   // columns >= N of this row: -inf
        const int v = p.N - j * A3_BK - ch * 32;
        if (v < 32) {
#pragma unroll
          for (int i = 0; i < 32; ++i) if (i >= v) r[i] = 0xff800000u;
        }
      };
      auto ld_chunk = [&](int j, int ch, uint32_t (&r)[32]) {     // 32 scores of this row (synchronous)
        ptx::tmem_ld_32x32b_x32(t_s + ch * 32, r);
        ptx::tmem_ld_wait();
        mask_chunk(j, ch, r);
      };
      auto max32 = [&](const uint32_t (&r)[32]) {
        float a0 = -INFINITY, a1 = -INFINITY, a2 = -INFINITY, a3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          a0 = fmaxf(a0, __uint_as_float(r[i])); a1 = fmaxf(a1, __uint_as_float(r[i + 1]));
          a2 = fmaxf(a2, __uint_as_float(r[i + 2])); a3 = fmaxf(a3, __uint_as_float(r[i + 3]));
        }
        return fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
      };

      for (int j = 0; j < nb; ++j) {
        ptx::mbar_wait(&s_full[x], j & 1);
        ptx::tc_fence_after();
        const int valid = min(A3_BK, p.N - j * A3_BK);             // > 0
        const int nch = (valid + 31) >> 5;                         // chunks that hold at least one key
        const int kch = (((valid + 15) & ~15) + 31) >> 5;          // chunks inside the P.V K range (zero-filled when dead)
        if (warp_live) {
          if (j == 0) {                                            // first block fixes the reference max of the row
            float mx = -INFINITY;
            for (int ch = 0; ch < nch; ++ch) { uint32_t r[32]; ld_chunk(0, ch, r); mx = fmaxf(mx, max32(r)); }
            m_ref = mx;
          }
          float l_in = l;
          bool redo = false;
          do {
            const float off = m_ref * c;
            float la = 0.f, lb = 0.f, lc = 0.f, ld = 0.f, mx_row = -INFINITY;
            // (a two-buffer variant that keeps the tcgen05.ld of chunk ch + 1 in flight measured 7 % slower: r02u)
#pragma unroll 1
            for (int ch = 0; ch < kch; ++ch) {
              uint32_t pk[16];
              if (ch < nch) {
                uint32_t r[32];
                ld_chunk(j, ch, r);
                if (j > 0 && !redo) mx_row = fmaxf(mx_row, max32(r));
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                  const float e0 = ptx::ex2_ftz(fmaf(__uint_as_float(r[i]), c, -off)), e1 = ptx::ex2_ftz(fmaf(__uint_as_float(r[i + 1]), c, -off));
                  const float e2 = ptx::ex2_ftz(fmaf(__uint_as_float(r[i + 2]), c, -off)), e3 = ptx::ex2_ftz(fmaf(__uint_as_float(r[i + 3]), c, -off));
                  la += e0; lb += e1; lc += e2; ld += e3;
                  pk[i >> 1] = pack_bf16x2(e0, e1); pk[(i >> 1) + 1] = pack_bf16x2(e2, e3);
                }
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) pk[i] = 0u;
              }
              ptx::tmem_st_32x32b_x16(t_p + ch * 16, pk);
            }
            // lazy rescale: only when some row of this warp exceeds its reference by more than 2^64 (tcgen05.ld / st are
            // warp-collective, so the warp goes through the slow path together; rows that do not need it use f = 1)
            const bool need = !redo && j > 0 && (mx_row - m_ref) * c > A3_RESCALE_LOG2;
            if (!redo && __any_sync(0xffffffffu, need)) {
              const float f = need ? ptx::ex2_ftz((m_ref - mx_row) * c) : 1.0f;
              if (need) m_ref = mx_row;
              l_in *= f;
              // every P.V issued so far retired before s_full(j) fired (tcgen05 ops retire in issue order): O is quiescent
              uint32_t o0[32];
#pragma unroll 1
              for (int hf = 0; hf < 2; ++hf) {
                ptx::tmem_ld_32x32b_x32(t_o + hf * 32, o0);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o0[i] = __float_as_uint(__uint_as_float(o0[i]) * f);
                ptx::tmem_st_32x32b_x32(t_o + hf * 32, o0);
              }
              ptx::tmem_st_wait();
              redo = true;                                         // exponentiate this block again against the new reference
            } else {
              l = l_in + ((la + lb) + (lc + ld));
              redo = false;
              break;
            }
          } while (true);
        }
        ptx::tmem_st_wait();                 // P is in tensor memory before the MMA thread is told so
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) { ptx::mbar_arrive(&s_empty[x]); ptx::mbar_arrive(&p_full[x]); }
      }
      const float o_scale = warp_live ? 1.0f / l : 0.f;
      if (p.lse != nullptr && warp_live && row < p.N)
        p.lse[((size_t)b * p.H + h) * p.N + row] = fmaf(m_ref, c, __log2f(l)) * 0.69314718055994531f;
      // ---- epilogue: O (TMEM) -> bf16 -> swizzled smem tile (the dead Q tile) -> one TMA store per tile ----
      ptx::mbar_wait(&o_full[x], 0);
      ptx::tc_fence_after();
      if (warp_live) {
#pragma unroll 1
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t r0[32];
          ptx::tmem_ld_32x32b_x32(t_o + hf * 32, r0);
          ptx::tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r0[i]) * o_scale;
          a3_store_row32(q_s + x * A3_TILE + r_local * 128, hf, r_local, v);
        }
      }
      ptx::fence_proxy_async_smem();
      asm volatile("bar.sync %0, 128;\n" ::"r"(1 + x) : "memory");
      if ((warp == 2 || warp == 6) && lane == 0) {
        ptx::tma_store_3d(&tm_o, q_s + x * A3_TILE, h * A3_HD, q0, b);     // rows >= N are clipped by the tensor map
        ptx::tma_store_commit();
        ptx::tma_store_wait_read<0>();
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem);
  }
}

int attention_tc3(const void* qkv, void* out, int B, int N, int H, float scale, int qt_offset, float* lse, cudaStream_t st) {
  CUtensorMap tm_qkv, tm_o;
  if (int rc = encode_tmap_3d(&tm_qkv, qkv, 2, B, N, 3 * H * A3_HD, 128, 64)) return rc;
  if (int rc = encode_tmap_3d(&tm_o, out, 2, B, N, H * A3_HD, 128, 64)) return rc;
  Attn3Params p;
  p.lse = lse; p.N = N; p.H = H;
  p.n_qt = (N + A3_BM - 1) / A3_BM;
  p.nb = (N + A3_BK - 1) / A3_BK;
  p.qt_offset = qt_offset;
  p.desc = g_walk_desc;
  p.scale_log2 = scale * 1.4426950408889634f;
  const int tiles = p.n_qt - qt_offset;
  if (tiles <= 0) return 0;
  static DeviceOnce once;
  if (once.first()) {
    TPAT_CUDA(cudaFuncSetAttribute(attention_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A3_SMEM));
    once.mark();
  }
  TPAT_CUDA(launch_kernel(attention_tc3_kernel, dim3((tiles + 1) / 2, H, B), dim3(A3_THREADS), (size_t)A3_SMEM, st, tm_qkv, tm_o, p));
  TPAT_LAUNCH_CHECK();
  return 0;
}

}  // namespace tpat
