// tcgen05 fused attention, fifth structure: attention_tc4.cu's independent key halves in a PERSISTENT kernel.
//
// Same math and the same per-block work as attention_tc4.cu (q k^T * scale, lazily rescaled online softmax per 32-key
// half with its own reference max / row sum / output accumulator, attn @ v; reference audiomae/models_vit.py:79-95).
// The r02aa trace of that kernel (profiles/r02aa_attention_v4_trace_n513.txt) shows a steady 1300 cycles per 64-key
// block (MUFU.EX2 ~80 % busy) but 1300 cycles before the first scores arrive, 1300 after the last block (O out of
// tensor memory, staging, TMA store) and the launch / TMEM allocation / barrier set-up of a CTA around that: a third of a
// tile's life at N = 513, more at the pruned token counts.  Here two CTAs per SM stay resident and walk the
// (clip, head, query tile) items round robin:
//   * the TMA warp runs ahead: the next item's Q tile (two Q buffers) and first K tile are in shared memory before the
//     current item's last block is done; the K / V ring never drains between items;
//   * the MMA thread issues S(0) of the next item before the last P.V of the current one, so the softmax warps find
//     scores waiting when they come out of the epilogue; the first P.V of an item waits until every softmax warp has read
//     the previous item's accumulators (o_empty);
//   * barriers, TMEM and tensor-map prefetch are set up once per CTA; S-buffer / P-buffer parities run on a global block
//     counter, so the "S(g+2) is issued after P(g).V(g)" ordering argument of attention_tc4.cu holds across items.
#include "attention.cuh"
#include "ptx_sm100.cuh"

#include <cstdlib>

namespace tpat {

int encode_tmap_3d_qkv(CUtensorMap* out, const void* gptr, int B, int N, int ld, int box_rows);

constexpr int A5_BM = 128, A5_BK = 64, A5_HD = 64;
constexpr int A5_SLOTS = 6;                         // K / V ring slots
constexpr int A5_Q_BYTES = A5_BM * A5_HD * 2;       // 16 KB per Q buffer (also the O staging tile of the item's epilogue)
constexpr int A5_KV_BYTES = A5_BK * A5_HD * 2;      // 8 KB
constexpr int A5_THREADS = 320;                     // TMA warp, MMA warp, 8 softmax warps
constexpr int A5_TMEM_COLS = 256;                   // S0 S1 [0, 128) (P_g over S_g), O_a [128, 192), O_b [192, 256)
constexpr int A5_SMEM = 1024 + 2 * A5_Q_BYTES + A5_SLOTS * A5_KV_BYTES + 256 + 2 * A5_BM * (int)sizeof(float2) + 64;
constexpr float A5_RESCALE_SUM = 18446744073709551616.0f;   // 2^64: a block sum above it raises the half's reference max

// Per-SM ticket counter (module global, zero at load, never reset): the two CTAs that share an SM draw alternating
// parities, and the odd one starts half an item late.  Persistent CTAs of one launch would otherwise run in lock-step
// (same start, same item sizes): both in their exponentials at once, both in their epilogues -- MUFU idle -- at once.
// The non-persistent kernel gets that de-phasing for free from the block scheduler.
__device__ unsigned int g_a5_ticket[1024];

#ifdef TPAT_ATTN_TRACE
// debug builds only: clock stamps of one softmax thread ([0, 120), count at [127]) and of the MMA thread ([128, 250), count at [255])
#define A5_TRACE(slot) do { if (tracing && trace_n < 120) p.trace[trace_base + trace_n++] = clock64() - t_start + ((long long)(slot) << 48); } while (0)
#else
#define A5_TRACE(slot) do { } while (0)
#endif

struct Attn5Params {
  long long* trace;      // TPAT_ATTN_TRACE builds only
  float* lse;            // optional [B, H, N] natural-log sum of exp(scale * s) per query row (training)
  int B, N, H, nb, qt_offset, n_tiles, n_items;
  int stagger;           // cycles the second CTA of an SM waits before its first item (0 = off)
  int desc;              // 1 = clips are visited from the last one down (g_walk_desc)
  float scale_log2;      // scale * log2(e)
};

struct A5Item { int b, h, q0; };
__device__ __forceinline__ A5Item a5_item(const Attn5Params& p, int it) {
  // query tile fastest: the tiles of one (clip, head) run at the same time on neighbouring CTAs and share K / V in L2
  const int qt = it % p.n_tiles, bh = it / p.n_tiles;
  const int b = bh / p.H;
  return A5Item{p.desc ? p.B - 1 - b : b, bh % p.H, (qt + p.qt_offset) * A5_BM};
}

__device__ __forceinline__ void a5_store_row32(uint8_t* tile_row, int hf, int r_local, const float (&v)[32]) {
#pragma unroll
  for (int g = 0; g < 4; ++g)
    *reinterpret_cast<uint4*>(tile_row + (((hf * 4 + g) ^ (r_local & 7)) * 16)) =
        make_uint4(pack_bf16x2(v[g * 8 + 0], v[g * 8 + 1]), pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]),
                   pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]), pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]));
}

__global__ void __launch_bounds__(A5_THREADS, 2)
attention_tc5_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                     const __grid_constant__ CUtensorMap tmap_o, const Attn5Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* q_s = smem;                                   // 2 x 16 KB
  uint8_t* kv_s = q_s + 2 * A5_Q_BYTES;                  // A5_SLOTS x 8 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(kv_s + A5_SLOTS * A5_KV_BYTES);
  uint64_t* q_full = bars;                   // [2]
  uint64_t* q_empty = bars + 2;              // [2]  the item's O store has drained the buffer
  uint64_t* kv_full = bars + 4;              // [SLOTS]
  uint64_t* kv_empty = kv_full + A5_SLOTS;   // [SLOTS]
  uint64_t* s_full = kv_empty + A5_SLOTS;    // [2]
  uint64_t* p_full = s_full + 2;             // [2]  8 arrivals (one per softmax warp)
  uint64_t* pv_done = p_full + 2;            // [2]  P(g).V(g) retired (only the rare rescale path waits on it)
  uint64_t* o_full = pv_done + 2;            // [1]  every P.V of the item retired
  uint64_t* o_empty = o_full + 1;            // [1]  8 arrivals: the item's accumulators are in registers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 1);
  float2* pair_s = reinterpret_cast<float2*>(bars + 32);   // [2 halves][128 rows]: (reference max, row sum), epilogue only

  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb = p.nb;
  const int it0 = blockIdx.x, it_step = gridDim.x;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_q);
    ptx::prefetch_tensormap(&tmap_kv);
    ptx::prefetch_tensormap(&tmap_o);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&q_full[i], 1); ptx::mbar_init(&q_empty[i], 1);
      ptx::mbar_init(&s_full[i], 1); ptx::mbar_init(&p_full[i], 8); ptx::mbar_init(&pv_done[i], 1);
    }
    for (int s = 0; s < A5_SLOTS; ++s) { ptx::mbar_init(&kv_full[s], 1); ptx::mbar_init(&kv_empty[s], 1); }
    ptx::mbar_init(o_full, 1);
    ptx::mbar_init(o_empty, 8);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc<A5_TMEM_COLS>(tmem_slot);
    ptx::tmem_relinquish();
  }
  if (warp == 3 && lane == 0 && p.stagger > 0) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    tmem_slot[1] = atomicAdd(&g_a5_ticket[smid & 1023], 1u) & 1u;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 2 * A5_BK;     // O_a; O_b = tmem_o + 64
  pdl_wait();   // everything above touched only on-chip state; global memory from here on
  if (p.stagger > 0 && tmem_slot[1] != 0 && warp >= 2) {         // softmax warps of the SM's second CTA: start late
    const long long t0 = clock64();
    while (clock64() - t0 < (long long)p.stagger) __nanosleep(200);
  }

  if (warp == 0) {
    // ===== TMA producer.  Per item: Q (one item ahead of the MMA thread), then the K / V tiles in the MMA thread's
    // consumption order: K_0 | K_1 V_0 | K_2 V_1 | ... | K_0(next item) V_{nb-1} =====
    if (ptx::elect_one()) {
      int slot = 0; uint32_t phase = 0;
      auto load_tile = [&](int col, int key0, int b) {
        ptx::mbar_wait(&kv_empty[slot], phase ^ 1);
        ptx::mbar_arrive_expect_tx(&kv_full[slot], A5_KV_BYTES);
        ptx::tma_load_3d(kv_s + slot * A5_KV_BYTES, &tmap_kv, &kv_full[slot], col, key0, b);
        if (++slot == A5_SLOTS) { slot = 0; phase ^= 1; }
      };
      auto load_q = [&](int n, const A5Item& im) {
        const int qb = n & 1;
        ptx::mbar_wait(&q_empty[qb], ((n >> 1) & 1) ^ 1);      // the store of item n - 2 has drained this buffer
        ptx::mbar_arrive_expect_tx(&q_full[qb], A5_Q_BYTES);
        ptx::tma_load_3d(q_s + qb * A5_Q_BYTES, &tmap_q, &q_full[qb], im.h * A5_HD, im.q0, im.b);
      };
      if (it0 < p.n_items) {
        A5Item im = a5_item(p, it0);
        load_q(0, im);
        load_tile((p.H + im.h) * A5_HD, 0, im.b);
        int n = 0;
        for (int it = it0; it < p.n_items; it += it_step, ++n) {
          const bool has_next = it + it_step < p.n_items;
          const A5Item nx = has_next ? a5_item(p, it + it_step) : im;
          const int col_k = (p.H + im.h) * A5_HD, col_v = (2 * p.H + im.h) * A5_HD;
          for (int j = 0; j < nb; ++j) {
            if (j + 1 < nb) load_tile(col_k, (j + 1) * A5_BK, im.b);
            else if (has_next) { load_q(n + 1, nx); load_tile((p.H + nx.h) * A5_HD, 0, nx.b); }
            load_tile(col_v, j * A5_BK, im.b);
          }
          im = nx;
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (ptx::elect_one() && it0 < p.n_items) {
      constexpr uint32_t idesc_s = ptx::idesc_bf16_f32(128, A5_BK, 0, 0);  // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_o = ptx::idesc_bf16_f32(128, A5_HD, 0, 1);  // P (TMEM, K-major) x V (MN-major)
      int slot = 0; uint32_t phase = 0;
      int g = 0;                                 // global block counter: S / P buffer = g & 1, use count = g >> 1
      // S(gg) goes into buffer gg & 1, whose last content P(gg-2) was read by P(gg-2).V(gg-2): issued earlier, and
      // tcgen05.mma executes in issue order.  The softmax warps finished reading S(gg-2) before they arrived on
      // p_full(gg-2), which this thread waited for before that product.  True across items: gg runs on.
      auto issue_s = [&](int gg, int qb) {
        const int sb = gg & 1;
        ptx::mbar_wait(&kv_full[slot], phase);
        ptx::tc_fence_after();
        const uint64_t q_desc = ptx::smem_desc_sw128(ptx::smem_u32(q_s + qb * A5_Q_BYTES), 16, 1024);
        const uint64_t k_desc = ptx::smem_desc_sw128(ptx::smem_u32(kv_s + slot * A5_KV_BYTES), 16, 1024);
#pragma unroll
        for (int k = 0; k < A5_HD / 16; ++k)
          ptx::mma_f16_ss(tmem_base + sb * A5_BK, q_desc + (uint64_t)(2 * k), k_desc + (uint64_t)(2 * k), idesc_s, k != 0);
        ptx::tc_commit(&kv_empty[slot]);
        ptx::tc_commit(&s_full[sb]);
        if (++slot == A5_SLOTS) { slot = 0; phase ^= 1; }
      };
#ifdef TPAT_ATTN_TRACE
      const bool tracing = p.trace != nullptr && blockIdx.x == 100;
      int trace_n = 0; const int trace_base = 128;
      const long long t_start = clock64();
      if (tracing) p.trace[254] = t_start;
#endif
      ptx::mbar_wait(&q_full[0], 0);
      A5_TRACE(20);
      issue_s(0, 0);
      int n = 0;
      for (int it = it0; it < p.n_items; it += it_step, ++n) {
        const bool has_next = it + it_step < p.n_items;
        for (int j = 0; j < nb; ++j, ++g) {
          if (j + 1 < nb) issue_s(g + 1, n & 1);                 // S(j+1) overlaps the softmax of block j
          else if (has_next) {                                   // ... or the next item's S(0)
            ptx::mbar_wait(&q_full[(n + 1) & 1], ((n + 1) >> 1) & 1);
            issue_s(g + 1, (n + 1) & 1);
          }
          const int sb = g & 1;
          A5_TRACE(21);
          ptx::mbar_wait(&kv_full[slot], phase);                 // V_j
          ptx::mbar_wait(&p_full[sb], (g >> 1) & 1);             // P_j written by the softmax warps (over S_j)
          A5_TRACE(22);
          if (j == 0 && n > 0) { ptx::mbar_wait(o_empty, (n - 1) & 1); A5_TRACE(23); }   // the previous item's accumulators were read out
          ptx::tc_fence_after();
          const uint32_t v_addr = ptx::smem_u32(kv_s + slot * A5_KV_BYTES);
          const int valid = min(A5_BK, p.N - j * A5_BK);         // keys of this block that exist
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int vh = min(32, valid - hf * 32);             // keys of this half that exist
            const int ksteps = vh > 0 ? (vh + 15) >> 4 : 0;      // P is zero beyond vh, V rows beyond N are zero-filled
            for (int k = 0; k < ksteps; ++k) {
              // A: this half's P from TMEM, 16 keys = 8 columns;  B (MN-major): 16 keys = two 8-row groups of 1024 B
              const uint64_t b_desc = ptx::smem_desc_sw128(v_addr + (hf * 2 + k) * 2048, 16, 1024);
              ptx::mma_f16_ts(tmem_o + hf * A5_HD, tmem_base + sb * A5_BK + hf * 32 + k * 8, b_desc, idesc_o, (j | k) != 0);
            }
          }
          ptx::tc_commit(&kv_empty[slot]);
          ptx::tc_commit(&pv_done[sb]);
          if (j == nb - 1) ptx::tc_commit(o_full);
          if (++slot == A5_SLOTS) { slot = 0; phase ^= 1; }
        }
      }
#ifdef TPAT_ATTN_TRACE
      if (tracing) p.trace[255] = trace_n;
#endif
    }
  } else {
    // ===== softmax / epilogue warps: TMEM lane quarter = warp % 4, thread = (query row, 32-key half) =====
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r_local = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const float c = p.scale_log2;
    const bool oth_has = p.N > (half ^ 1) * 32;          // the partner's half sees at least one key (block 0)
    const bool own_has = p.N > half * 32;
    const bool storer = warp == 2 && lane == 0;

    auto max32 = [&](const uint32_t (&r)[32]) {
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(r[i])); mx1 = fmaxf(mx1, __uint_as_float(r[i + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(r[i + 2])); mx3 = fmaxf(mx3, __uint_as_float(r[i + 3]));
      }
      return fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
    };
    auto retire_store = [&](int qb) {                    // storer only: the bulk store has read its staging tile (Q buffer qb)
      ptx::tma_store_wait_read<0>();
      ptx::mbar_arrive(&q_empty[qb]);
    };

#ifdef TPAT_ATTN_TRACE
    const bool tracing = p.trace != nullptr && blockIdx.x == 100 && threadIdx.x == 64;
    int trace_n = 0; const int trace_base = 0;
    long long t_start = 0;
    if (tracing) { t_start = clock64(); p.trace[126] = t_start; }
#endif
    int g = 0, n = 0;
    for (int it = it0; it < p.n_items; it += it_step, ++n) {
      A5_TRACE(1);
      // A warp whose 32 query rows all lie beyond N only keeps the barrier protocol going.
      const bool warp_live = a5_item(p, it).q0 + quarter * 32 < p.N;
      float m_ref = -INFINITY;                           // this half's reference max (raw score units)
      float la = 0.f, lb = 0.f, lc = 0.f, ld = 0.f;      // this half's row sum, four partial accumulators

      for (int j = 0; j < nb; ++j, ++g) {
        const int sb = g & 1;
        ptx::mbar_wait(&s_full[sb], (g >> 1) & 1);
        ptx::tc_fence_after();
        A5_TRACE(3);
        const int vh = p.N - j * A5_BK - half * 32;      // valid columns in this thread's half (may be <= 0: no MMA k-step reads them)
        if (warp_live && vh > 0) {
          const uint32_t t_sp = tmem_base + lane_off + sb * A5_BK + half * 32;   // own 32 score columns; P over the first 16
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(t_sp, r);
          ptx::tmem_ld_wait();
          if (vh < 32) {
#pragma unroll
            for (int i = 0; i < 32; ++i) if (i >= vh) r[i] = 0xff800000u;        // -inf -> probability 0
          }
          if (j == 0) m_ref = max32(r);                  // finite: vh > 0
          float la_in = la, lb_in = lb, lc_in = lc, ld_in = ld;
          float bsum;
          auto emit = [&]() {
            const float off = m_ref * c;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            uint32_t pk[16];                             // bf16 pairs: the K-major A operand of P.V, straight into TMEM
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
          A5_TRACE(5);
          // Lazy rescale (see attention_tc4.cu): sum-triggered, warp-collective slow path, f = 1 for rows that do not need it
          const bool need = j > 0 && !(bsum <= A5_RESCALE_SUM);
          if (__any_sync(0xffffffffu, need)) {
            const float mx = fmaxf(max32(r), m_ref);
            const float f = need ? ptx::ex2_ftz((m_ref - mx) * c) : 1.0f;
            if (need) m_ref = mx;
            la_in *= f; lb_in *= f; lc_in *= f; ld_in *= f;
            // every P.V issued so far (up to block g-1, same item since j > 0) must have retired before O is touched
            ptx::mbar_wait(&pv_done[(g - 1) & 1], ((g - 1) >> 1) & 1);
            ptx::tc_fence_after();
#pragma unroll 1
            for (int ch = 0; ch < 2; ++ch) {             // this half's own accumulator: all 64 output columns of the row
              uint32_t o0[32];
              ptx::tmem_ld_32x32b_x32(tmem_o + lane_off + half * A5_HD + ch * 32, o0);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o0[i] = __float_as_uint(__uint_as_float(o0[i]) * f);
              ptx::tmem_st_32x32b_x32(tmem_o + lane_off + half * A5_HD + ch * 32, o0);
            }
            emit();                                      // the block again, against the new reference (overwrites P_j)
          }
          ptx::tmem_st_wait();             // P is in tensor memory before the MMA thread is told so
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&p_full[sb]);
        A5_TRACE(7);
        if (storer && j == 0 && n > 0) retire_store((n - 1) & 1);   // the previous item's O store has long finished reading its tile
      }

      // ---- merge the two halves of every row: one exchange per item ----
      int it_again = it;
      asm volatile("" : "+r"(it_again));                 // (decoded again: keeps b / h / q0 out of the block loop's live set)
      const A5Item im = a5_item(p, it_again);
      const int row = im.q0 + r_local;
      float w_own = 0.f, w_oth = 0.f, inv_l = 0.f;
      if (warp_live) {
        const float l_own = (la + lb) + (lc + ld);
        pair_s[half * A5_BM + r_local] = make_float2(m_ref, l_own);
        asm volatile("bar.sync %0, 64;\n" ::"r"(2 + quarter) : "memory");
        const float2 o = pair_s[(half ^ 1) * A5_BM + r_local];
        const float m = fmaxf(m_ref, o.x);               // finite: key 0 exists (half 0, block 0)
        w_own = own_has ? ptx::ex2_ftz((m_ref - m) * c) : 0.f;
        w_oth = oth_has ? ptx::ex2_ftz((o.x - m) * c) : 0.f;
        const float l_tot = l_own * w_own + o.y * w_oth;
        inv_l = 1.0f / l_tot;
        if (p.lse != nullptr && half == 0 && row < p.N)
          p.lse[((size_t)im.b * p.H + im.h) * p.N + row] = fmaf(m, c, __log2f(l_tot)) * 0.69314718055994531f;
      }
      // ---- epilogue: (w_a O_a + w_b O_b) / l -> bf16 -> swizzled smem tile (the item's dead Q tile) -> one TMA store ----
      A5_TRACE(9);
      ptx::mbar_wait(o_full, n & 1);       // every P.V of the item retired (and every S: the Q tile is dead)
      ptx::tc_fence_after();
      A5_TRACE(10);
      float v[32];
      if (warp_live) {
        // this thread writes output columns [32 half, 32 half + 32) of its row: the same columns of both accumulators
        const float w_a = (half == 0 ? w_own : w_oth) * inv_l, w_b = (half == 0 ? w_oth : w_own) * inv_l;
        {
          uint32_t r0[32];
          ptx::tmem_ld_32x32b_x32(tmem_o + lane_off + half * 32, r0);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r0[i]) * w_a;
        }
        if (p.N > 32) {                    // O_b is never written when no key reaches the second half
          uint32_t r1[32];
          ptx::tmem_ld_32x32b_x32(tmem_o + lane_off + A5_HD + half * 32, r1);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaf(__uint_as_float(r1[i]), w_b, v[i]);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(o_empty);          // the next item's first P.V may overwrite the accumulators
      A5_TRACE(12);
      uint8_t* stage = q_s + (n & 1) * A5_Q_BYTES;
      if (warp_live) a5_store_row32(stage + r_local * 128, half, r_local, v);
      ptx::fence_proxy_async_smem();
      asm volatile("bar.sync 1, 256;\n" ::: "memory");
      A5_TRACE(13);
      if (storer) {
        ptx::tma_store_3d(&tmap_o, stage, im.h * A5_HD, im.q0, im.b);   // rows >= N are clipped by the tensor map
        ptx::tma_store_commit();
      }
      A5_TRACE(11);
    }
#ifdef TPAT_ATTN_TRACE
    if (tracing) p.trace[127] = trace_n;
#endif
    if (storer && n > 0) retire_store((n - 1) & 1);      // smem must outlive the last bulk store's reads
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<A5_TMEM_COLS>(tmem_base);
  }
}

int attention_tc5(const void* qkv, void* out, int B, int N, int H, float scale, int qt_offset, float* lse, cudaStream_t st) {
  CUtensorMap tm_q, tm_kv, tm_o;
  if (int rc = encode_tmap_3d_qkv(&tm_q, qkv, B, N, 3 * H * A5_HD, A5_BM)) return rc;
  if (int rc = encode_tmap_3d_qkv(&tm_kv, qkv, B, N, 3 * H * A5_HD, A5_BK)) return rc;
  if (int rc = encode_tmap_3d_qkv(&tm_o, out, B, N, H * A5_HD, A5_BM)) return rc;
  Attn5Params p;
  p.trace = nullptr;
#ifdef TPAT_ATTN_TRACE
  { extern long long* g_attn_trace_buf; p.trace = g_attn_trace_buf; }
#endif
  p.lse = lse; p.B = B; p.N = N; p.H = H;
  p.nb = (N + A5_BK - 1) / A5_BK;
  p.qt_offset = qt_offset;
  p.n_tiles = (N + A5_BM - 1) / A5_BM - qt_offset;
  if (p.n_tiles <= 0) return 0;
  p.n_items = p.n_tiles * H * B;
  p.desc = g_walk_desc;
  p.scale_log2 = scale * 1.4426950408889634f;
  int grid = 2 * sm_count();
  if (grid > p.n_items) grid = p.n_items;
  // start offset of every SM's second CTA: TPAT_A5_STAGGER_PCT percent (default 0 = off: measured no effect) of one item's ~1300 (nb + 1) cycles
  static const int stagger_pct = [] { const char* e = getenv("TPAT_A5_STAGGER_PCT"); return e ? atoi(e) : 0; }();
  p.stagger = p.n_items > grid ? 13 * (p.nb + 1) * stagger_pct : 0;
  static DeviceOnce once;
  if (once.first()) {
    TPAT_CUDA(cudaFuncSetAttribute(attention_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A5_SMEM));
    TPAT_CUDA(cudaFuncSetAttribute(attention_tc5_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    once.mark();
  }
  TPAT_CUDA(launch_kernel(attention_tc5_kernel, dim3(grid), dim3(A5_THREADS), (size_t)A5_SMEM, st, tm_q, tm_kv, tm_o, p));
  TPAT_LAUNCH_CHECK();
  return 0;
}

}  // namespace tpat
