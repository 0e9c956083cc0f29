// tcgen05 fused attention, fourth structure ("independent key halves"): the tiles that do not feed an importance score
// (9 of 12 blocks in inference, every block's attention in training).
//
// Same math as attention_tc.cu's single-pass instantiation (q k^T * scale, lazily rescaled online softmax, attn @ v;
// reference audiomae/models_vit.py:79-95) and the same occupancy (one CTA per (clip, head, 128-query tile), two CTAs per
// SM, 8 softmax warps, thread = (query row, 32-key half of every 64-key block)) -- but the two partner warps of a row
// never talk to each other inside the key loop.  The r01 / r02 traces put ~300 of the ~2000 cycles per
// 64-key block on the partners' exchange (row max of the own columns, warp vote, flag through shared memory, 64-thread
// named barrier, flag read) and it also locks the two warps of a sub-partition into the same phase, so that they queue
// for MUFU.EX2 together and leave it idle together.  Here
//   * every half keeps its OWN reference max, row sum and OUTPUT ACCUMULATOR: O_a += P[:, 0:32] V[0:32], O_b += P[:, 32:64]
//     V[32:64] (two K = 32 products instead of one K = 64 product: the same four tcgen05.mma per block); the halves are
//     merged once, in the epilogue: O = (w_a O_a + w_b O_b) / (w_a l_a + w_b l_b), w_x = 2^(c (m_x - max(m_a, m_b)));
//   * the lazy-rescale test needs no row max: the block's sum of exponentials (already computed) exceeding 2^64 is the
//     trigger (some element > 2^59); only then is the max taken, the half's O and sum rescaled and the block redone;
//   * P_j is written over the first 16 columns of the thread's own 32 score columns (S_j is in registers by then), so the
//     256 TMEM columns hold S0 S1 | O_a | O_b; tcgen05.mma executes in issue order and S(j+2) is issued after P(j).V(j),
//     so the S buffers need no "empty" barrier at all.
// Per block a softmax warp now does: wait s_full, tcgen05.ld, 32 x (FFMA, EX2, FADD), pack, tcgen05.st, arrive.
#include "attention.cuh"
#include "ptx_sm100.cuh"

#include <cstdlib>

namespace tpat {

int encode_tmap_3d_qkv(CUtensorMap* out, const void* gptr, int B, int N, int ld, int box_rows);

constexpr int A4_BM = 128, A4_BK = 64, A4_HD = 64;
constexpr int A4_SLOTS = 6;                         // K / V ring slots
constexpr int A4_Q_BYTES = A4_BM * A4_HD * 2;       // 16 KB (also the O staging tile of the epilogue)
constexpr int A4_KV_BYTES = A4_BK * A4_HD * 2;      // 8 KB
constexpr int A4_THREADS = 320;                     // TMA warp, MMA warp, 8 softmax warps
constexpr int A4_TMEM_COLS = 256;                   // S0 S1 [0, 128) (P_j over S_j), O_a [128, 192), O_b [192, 256)
constexpr int A4_SMEM = 1024 + A4_Q_BYTES + A4_SLOTS * A4_KV_BYTES + 256 + 2 * A4_BM * (int)sizeof(float2) + 64;
constexpr float A4_RESCALE_SUM = 18446744073709551616.0f;   // 2^64: a block sum above it raises the half's reference max

#ifdef TPAT_ATTN_TRACE
// debug builds only: clock stamps of one softmax thread ([0, 120), count at [127]) and of the MMA thread ([128, 250), count at [255])
#define A4_TRACE(slot) do { if (tracing && trace_n < 120) p.trace[trace_base + trace_n++] = clock64() - t_start + ((long long)(slot) << 48); } while (0)
#else
#define A4_TRACE(slot) do { } while (0)
#endif

struct Attn4Params {
  long long* trace;      // TPAT_ATTN_TRACE builds only
  float* lse;            // optional [B, H, N] natural-log sum of exp(scale * s) per query row (training)
  int N, H, nb, qt_offset;
  int desc;              // 1 = clips are visited from the last one down (g_walk_desc)
  float scale_log2;      // scale * log2(e)
};

__device__ __forceinline__ void a4_store_row32(uint8_t* tile_row, int hf, int r_local, const float (&v)[32]) {
#pragma unroll
  for (int g = 0; g < 4; ++g)
    *reinterpret_cast<uint4*>(tile_row + (((hf * 4 + g) ^ (r_local & 7)) * 16)) =
        make_uint4(pack_bf16x2(v[g * 8 + 0], v[g * 8 + 1]), pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]),
                   pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]), pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]));
}

__global__ void __launch_bounds__(A4_THREADS, 2)
attention_tc4_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                     const __grid_constant__ CUtensorMap tmap_o, const Attn4Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* q_s = smem;                                   // 16 KB
  uint8_t* kv_s = q_s + A4_Q_BYTES;                      // A4_SLOTS x 8 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(kv_s + A4_SLOTS * A4_KV_BYTES);
  uint64_t* q_full = bars;                   // [1]
  uint64_t* kv_full = bars + 1;              // [SLOTS]
  uint64_t* kv_empty = kv_full + A4_SLOTS;   // [SLOTS]
  uint64_t* s_full = kv_empty + A4_SLOTS;    // [2]
  uint64_t* p_full = s_full + 2;             // [2]  8 arrivals (one per softmax warp)
  uint64_t* pv_done = p_full + 2;            // [2]  P(j).V(j) retired (only the rare rescale path waits on it)
  uint64_t* o_full = pv_done + 2;            // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);
  float2* pair_s = reinterpret_cast<float2*>(bars + 32);   // [2 halves][128 rows]: (reference max, row sum), epilogue only

  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x + p.qt_offset, h = blockIdx.y;
  const int b = p.desc ? (int)(gridDim.z - 1 - blockIdx.z) : (int)blockIdx.z;
  const int q0 = qt * A4_BM;
  const int nb = p.nb;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_q);
    ptx::prefetch_tensormap(&tmap_kv);
    ptx::prefetch_tensormap(&tmap_o);
  }
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(q_full, 1);
    for (int s = 0; s < A4_SLOTS; ++s) { ptx::mbar_init(&kv_full[s], 1); ptx::mbar_init(&kv_empty[s], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&s_full[i], 1); ptx::mbar_init(&p_full[i], 8); ptx::mbar_init(&pv_done[i], 1); }
    ptx::mbar_init(o_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc<A4_TMEM_COLS>(tmem_slot);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 2 * A4_BK;     // O_a; O_b = tmem_o + 64
  pdl_wait();   // everything above touched only on-chip state; global memory from here on

  const int col_q = h * A4_HD, col_k = (p.H + h) * A4_HD, col_v = (2 * p.H + h) * A4_HD;

  if (warp == 0) {
    // ===== TMA producer: Q once, then K_0, K_1, V_0, K_2, V_1, ..., V_{nb-1} (the MMA thread's consumption order) =====
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(q_full, A4_Q_BYTES);
      ptx::tma_load_3d(q_s, &tmap_q, q_full, col_q, q0, b);
      int slot = 0; uint32_t phase = 0;
      auto load_tile = [&](int col, int key0) {
        ptx::mbar_wait(&kv_empty[slot], phase ^ 1);
        ptx::mbar_arrive_expect_tx(&kv_full[slot], A4_KV_BYTES);
        ptx::tma_load_3d(kv_s + slot * A4_KV_BYTES, &tmap_kv, &kv_full[slot], col, key0, b);
        if (++slot == A4_SLOTS) { slot = 0; phase ^= 1; }
      };
      load_tile(col_k, 0);
      for (int j = 0; j < nb; ++j) {
        if (j + 1 < nb) load_tile(col_k, (j + 1) * A4_BK);
        load_tile(col_v, j * A4_BK);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (ptx::elect_one()) {
      constexpr uint32_t idesc_s = ptx::idesc_bf16_f32(128, A4_BK, 0, 0);  // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_o = ptx::idesc_bf16_f32(128, A4_HD, 0, 1);  // P (TMEM, K-major) x V (MN-major)
      int slot = 0; uint32_t phase = 0;
      const uint64_t q_desc = ptx::smem_desc_sw128(ptx::smem_u32(q_s), 16, 1024);
      // S(jj) goes into buffer jj & 1, whose last content P(jj-2) was read by P(jj-2).V(jj-2): issued earlier, and
      // tcgen05.mma executes in issue order -- no barrier needed.  The softmax warps finished reading S(jj-2) before
      // they arrived on p_full(jj-2), which this thread waited for before that product.
      auto issue_s = [&](int jj) {
        const int sb = jj & 1;
        ptx::mbar_wait(&kv_full[slot], phase);
        ptx::tc_fence_after();
        const uint64_t k_desc = ptx::smem_desc_sw128(ptx::smem_u32(kv_s + slot * A4_KV_BYTES), 16, 1024);
#pragma unroll
        for (int k = 0; k < A4_HD / 16; ++k)
          ptx::mma_f16_ss(tmem_base + sb * A4_BK, q_desc + (uint64_t)(2 * k), k_desc + (uint64_t)(2 * k), idesc_s, k != 0);
        ptx::tc_commit(&kv_empty[slot]);
        ptx::tc_commit(&s_full[sb]);
        if (++slot == A4_SLOTS) { slot = 0; phase ^= 1; }
      };
#ifdef TPAT_ATTN_TRACE
      const bool tracing = p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 3 && blockIdx.z == (gridDim.z >> 1);
      int trace_n = 0; const int trace_base = 128;
      const long long t_start = clock64();
      if (tracing) p.trace[254] = t_start;
#endif
      ptx::mbar_wait(q_full, 0);
      A4_TRACE(20);
      issue_s(0);
      for (int j = 0; j < nb; ++j) {
        if (j + 1 < nb) issue_s(j + 1);                  // S(j+1) overlaps the softmax of block j
        A4_TRACE(21);
        const int sb = j & 1;
        ptx::mbar_wait(&kv_full[slot], phase);           // V_j
        ptx::mbar_wait(&p_full[sb], (j >> 1) & 1);       // P_j written by the softmax warps (over S_j)
        ptx::tc_fence_after();
        A4_TRACE(22);
        const uint32_t v_addr = ptx::smem_u32(kv_s + slot * A4_KV_BYTES);
        const int valid = min(A4_BK, p.N - j * A4_BK);   // keys of this block that exist
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int vh = min(32, valid - hf * 32);       // keys of this half that exist
          const int ksteps = vh > 0 ? (vh + 15) >> 4 : 0;   // P is zero beyond vh, V rows beyond N are zero-filled
          for (int k = 0; k < ksteps; ++k) {
            // A: this half's P from TMEM, 16 keys = 8 columns;  B (MN-major): 16 keys = two 8-row groups of 1024 B
            const uint64_t b_desc = ptx::smem_desc_sw128(v_addr + (hf * 2 + k) * 2048, 16, 1024);
            ptx::mma_f16_ts(tmem_o + hf * A4_HD, tmem_base + sb * A4_BK + hf * 32 + k * 8, b_desc, idesc_o, (j | k) != 0);
          }
        }
        ptx::tc_commit(&kv_empty[slot]);
        ptx::tc_commit(&pv_done[sb]);
        if (++slot == A4_SLOTS) { slot = 0; phase ^= 1; }
      }
      ptx::tc_commit(o_full);
#ifdef TPAT_ATTN_TRACE
      if (tracing) p.trace[255] = trace_n;
#endif
    }
  } else {
    // ===== softmax / epilogue warps: TMEM lane quarter = warp % 4, thread = (query row, 32-key half) =====
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r_local = quarter * 32 + lane;
    const int row = q0 + r_local;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const float c = p.scale_log2;
    // A warp whose 32 query rows all lie beyond N only keeps the barrier protocol going.
    const bool warp_live = q0 + quarter * 32 < p.N;
    float m_ref = -INFINITY;                             // this half's reference max (raw score units)
    float la = 0.f, lb = 0.f, lc = 0.f, ld = 0.f;        // this half's row sum, four partial accumulators

    auto max32 = [&](const uint32_t (&r)[32]) {
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(r[i])); mx1 = fmaxf(mx1, __uint_as_float(r[i + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(r[i + 2])); mx3 = fmaxf(mx3, __uint_as_float(r[i + 3]));
      }
      return fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
    };

#ifdef TPAT_ATTN_TRACE
    const bool tracing = p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 3 && blockIdx.z == (gridDim.z >> 1) && threadIdx.x == 64;
    int trace_n = 0; const int trace_base = 0;
    long long t_start = 0;
    if (tracing) { t_start = clock64(); p.trace[126] = t_start; }
    A4_TRACE(1);
#endif
    for (int j = 0; j < nb; ++j) {
      const int sb = j & 1;
      A4_TRACE(2);
      ptx::mbar_wait(&s_full[sb], (j >> 1) & 1);
      ptx::tc_fence_after();
      A4_TRACE(3);
      const int vh = p.N - j * A4_BK - half * 32;        // valid columns in this thread's half (may be <= 0: no MMA k-step reads them)
      if (warp_live && vh > 0) {
        const uint32_t t_sp = tmem_base + lane_off + sb * A4_BK + half * 32;   // own 32 score columns; P over the first 16
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(t_sp, r);
        ptx::tmem_ld_wait();
        A4_TRACE(4);
        if (vh < 32) {
#pragma unroll
          for (int i = 0; i < 32; ++i) if (i >= vh) r[i] = 0xff800000u;        // -inf -> probability 0
        }
        if (j == 0) m_ref = max32(r);                    // finite: vh > 0
        float la_in = la, lb_in = lb, lc_in = lc, ld_in = ld;
        float bsum;
        auto emit = [&]() {
          const float off = m_ref * c;
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
          uint32_t pk[16];                               // bf16 pairs: the K-major A operand of P.V, straight into TMEM
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float e0 = ptx::ex2_ftz(fmaf(__uint_as_float(r[i]), c, -off)), e1 = ptx::ex2_ftz(fmaf(__uint_as_float(r[i + 1]), c, -off));
            const float e2 = ptx::ex2_ftz(fmaf(__uint_as_float(r[i + 2]), c, -off)), e3 = ptx::ex2_ftz(fmaf(__uint_as_float(r[i + 3]), c, -off));
            a0 += e0; a1 += e1; a2 += e2; a3 += e3;
            pk[i >> 1] = pack_bf16x2(e0, e1); pk[(i >> 1) + 1] = pack_bf16x2(e2, e3);
          }
          ptx::tmem_st_32x32b_x16(t_sp, pk);
          la = la_in + a0; lb = lb_in + a1; lc = lc_in + a2; ld = ld_in + a3;
          bsum = (a0 + a1) + (a2 + a3);
        };
        emit();
        A4_TRACE(5);
        // Lazy rescale: a block sum above 2^64 (or inf) means some probability left the comfortable range.  Rare; the
        // warp goes through the slow path together (tcgen05.ld / st are warp-collective), rows that do not need it use
        // f = 1.  No row max is evaluated on the fast path.
        const bool need = j > 0 && !(bsum <= A4_RESCALE_SUM);
        if (__any_sync(0xffffffffu, need)) {
          const float mx = fmaxf(max32(r), m_ref);
          const float f = need ? ptx::ex2_ftz((m_ref - mx) * c) : 1.0f;
          if (need) m_ref = mx;
          la_in *= f; lb_in *= f; lc_in *= f; ld_in *= f;
          // every P.V issued so far (up to block j-1) must have retired before O is touched
          ptx::mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
          ptx::tc_fence_after();
#pragma unroll 1
          for (int ch = 0; ch < 2; ++ch) {               // this half's own accumulator: all 64 output columns of the row
            uint32_t o0[32];
            ptx::tmem_ld_32x32b_x32(tmem_o + lane_off + half * A4_HD + ch * 32, o0);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o0[i] = __float_as_uint(__uint_as_float(o0[i]) * f);
            ptx::tmem_st_32x32b_x32(tmem_o + lane_off + half * A4_HD + ch * 32, o0);
          }
          emit();                                        // the block again, against the new reference (overwrites P_j)
        }
        ptx::tmem_st_wait();             // P is in tensor memory before the MMA thread is told so
        A4_TRACE(6);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&p_full[sb]);
      A4_TRACE(7);
    }

    // ---- merge the two halves of every row: one exchange per tile ----
    float w_own = 0.f, w_oth = 0.f, inv_l = 0.f;
    const bool oth_has = p.N > (half ^ 1) * 32;          // the partner's half saw at least one key (block 0)
    const bool own_has = p.N > half * 32;
    if (warp_live) {
      const float l_own = (la + lb) + (lc + ld);
      pair_s[half * A4_BM + r_local] = make_float2(m_ref, l_own);
      asm volatile("bar.sync %0, 64;\n" ::"r"(2 + quarter) : "memory");
      const float2 o = pair_s[(half ^ 1) * A4_BM + r_local];
      const float m = fmaxf(m_ref, o.x);                 // finite: key 0 exists (half 0, block 0)
      w_own = own_has ? ptx::ex2_ftz((m_ref - m) * c) : 0.f;
      w_oth = oth_has ? ptx::ex2_ftz((o.x - m) * c) : 0.f;
      const float l_tot = l_own * w_own + o.y * w_oth;
      inv_l = 1.0f / l_tot;
      if (p.lse != nullptr && half == 0 && row < p.N)
        p.lse[((size_t)b * p.H + h) * p.N + row] = fmaf(m, c, __log2f(l_tot)) * 0.69314718055994531f;
    }
    // ---- epilogue: (w_a O_a + w_b O_b) / l -> bf16 -> swizzled smem tile (the dead Q tile) -> one TMA store ----
    A4_TRACE(9);
    ptx::mbar_wait(o_full, 0);           // every P.V retired
    ptx::tc_fence_after();
    A4_TRACE(10);
    if (warp_live) {
      // this thread writes output columns [32 half, 32 half + 32) of its row: the same columns of both accumulators
      const float w_a = (half == 0 ? w_own : w_oth) * inv_l, w_b = (half == 0 ? w_oth : w_own) * inv_l;
      float v[32];
      {
        uint32_t r0[32];
        ptx::tmem_ld_32x32b_x32(tmem_o + lane_off + half * 32, r0);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r0[i]) * w_a;
      }
      if (p.N > 32) {                    // O_b was never written when no key reaches the second half
        uint32_t r1[32];
        ptx::tmem_ld_32x32b_x32(tmem_o + lane_off + A4_HD + half * 32, r1);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaf(__uint_as_float(r1[i]), w_b, v[i]);
      }
      a4_store_row32(q_s + r_local * 128, half, r_local, v);     // Q is dead: its tile stages O
    }
    ptx::fence_proxy_async_smem();
    asm volatile("bar.sync 1, 256;\n" ::: "memory");
    if (warp == 2 && lane == 0) {
      ptx::tma_store_3d(&tmap_o, q_s, h * A4_HD, q0, b);   // rows >= N are clipped by the tensor map
      ptx::tma_store_commit();
      ptx::tma_store_wait_read<0>();                       // smem must outlive the bulk store's reads
    }
    A4_TRACE(11);
#ifdef TPAT_ATTN_TRACE
    if (tracing) p.trace[127] = trace_n;
#endif
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<A4_TMEM_COLS>(tmem_base);
  }
}

int attention_tc4(const void* qkv, void* out, int B, int N, int H, float scale, int qt_offset, float* lse, cudaStream_t st) {
  CUtensorMap tm_q, tm_kv, tm_o;
  if (int rc = encode_tmap_3d_qkv(&tm_q, qkv, B, N, 3 * H * A4_HD, A4_BM)) return rc;
  if (int rc = encode_tmap_3d_qkv(&tm_kv, qkv, B, N, 3 * H * A4_HD, A4_BK)) return rc;
  if (int rc = encode_tmap_3d_qkv(&tm_o, out, B, N, H * A4_HD, A4_BM)) return rc;
  Attn4Params p;
  p.trace = nullptr;
#ifdef TPAT_ATTN_TRACE
  { extern long long* g_attn_trace_buf; p.trace = g_attn_trace_buf; }
#endif
  p.lse = lse; p.N = N; p.H = H;
  p.nb = (N + A4_BK - 1) / A4_BK;
  p.qt_offset = qt_offset;
  p.desc = g_walk_desc;
  p.scale_log2 = scale * 1.4426950408889634f;
  const int tiles = (N + A4_BM - 1) / A4_BM - qt_offset;
  if (tiles <= 0) return 0;
  static DeviceOnce once;
  if (once.first()) {
    TPAT_CUDA(cudaFuncSetAttribute(attention_tc4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A4_SMEM));
    TPAT_CUDA(cudaFuncSetAttribute(attention_tc4_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    once.mark();
  }
  TPAT_CUDA(launch_kernel(attention_tc4_kernel, dim3(tiles, H, B), dim3(A4_THREADS), (size_t)A4_SMEM, st, tm_q, tm_kv, tm_o, p));
  TPAT_LAUNCH_CHECK();
  return 0;
}

}  // namespace tpat
