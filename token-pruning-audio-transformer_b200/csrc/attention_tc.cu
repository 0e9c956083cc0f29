// tcgen05 fused attention + importance-score partials for sm_100a (bf16 operands, fp32 softmax).
//
// Replaces q k^T * scale, softmax, attn @ v and the score slice/mean over the materialised
// [B,H,N,N] matrix (reference audiomae/models_vit.py:79-95,113; ast/src/models/ast_models.py:92-109,124).
// One CTA per (clip, head, 128-query tile), two CTAs resident per SM (~68 KB smem, 256 TMEM
// columns each: S0 S1 | O | P0 P1) so one CTA's softmax hides the other's prologue / MMA round trips.  320 threads:
//   warp 0   TMA producer: Q tile once, then K / V 64-key tiles into a 6-slot 8 KB ring
//            (3-D tensor maps over qkv[B][N][3*H*64]: rows >= N are zero-filled, never the next clip)
//   warp 1   MMA issuer (one elected thread): S = Q K^T (M=128,N=64,K=64) into one of two TMEM
//            buffers; O += P V (M=128,N=64,K<=64) with P read from TENSOR MEMORY (tcgen05.mma, A operand in TMEM:
//            the softmax warps tcgen05.st their bf16 pairs there), V as MN-major B from the TMA tile
//   warps 2-9 softmax: thread = (query row = TMEM lane, 32-key half); warps w and w+4 share a row quarter and
//            split each 64-key block by columns (four softmax warps per SM sub-partition with two CTAs resident).
//
// Two instantiations:
//  TWO_PASS = true  (tiles that must emit NORMALISED probabilities for the importance score; the
//            score needs the final row normaliser before any column can be accumulated)
//            pass 1: S -> running row max and sum of exp;  pass 2: S recomputed ->
//            P = 2^(s*c - m*c - log2 l) (one FFMA + one MUFU.EX2 per element) -> score partials from
//            the fp32 probabilities (warp transpose-reduce, fixed order, no atomics) -> bf16 P -> O += P V.
//  TWO_PASS = false (every other tile) single pass with a lazily rescaled online softmax: tile 0 fixes
//            the reference max m_ref; later tiles exponentiate against the current m_ref at once (their own
//            max is evaluated off the critical path) and only when a row's max exceeds m_ref by more than
//            2^64 is m_ref raised, O (TMEM) and the running sum rescaled by the softmax warp itself and the
//            tile redone.  P stays <= 2^64 (exact in bf16 / fp32 range).  O is divided by the row sum at the
//            end.  One exp and one QK^T per element, K streamed once.
// The N x N matrix never leaves the SM; O leaves through one TMA store staged in the (dead) Q tile.
#include "attention.cuh"
#include "ptx_sm100.cuh"

#include <cstdlib>

namespace tpat {

int encode_tmap_3d_qkv(CUtensorMap* out, const void* gptr, int B, int N, int ld, int box_rows);

constexpr int AT_BM = 128;          // queries per CTA
constexpr int AT_BK = 64;           // keys per block
constexpr int AT_HD = 64;
#define AT_CTAS_PER_SM 2            // resident CTAs per SM (256 of the 512 TMEM columns each)
constexpr int AT_SBUF = 2;          // S accumulators in TMEM
constexpr int AT_SLOTS = 6;         // K/V ring slots
constexpr int AT_Q_BYTES = AT_BM * AT_HD * 2;    // 16 KB (also the O staging tile of the epilogue)
constexpr int AT_KV_BYTES = AT_BK * AT_HD * 2;   // 8 KB
constexpr int AT_P_COLS = AT_BK / 2;             // one P buffer in TMEM: 64 bf16 keys = 32 columns of 32 bits
constexpr int AT_THREADS = 320;          // TMA warp, MMA warp, 8 softmax warps
constexpr int AT_TMEM_COLS = 256;   // S0, S1 [0, 128), O [128, 192), P0, P1 (bf16 pairs) [192, 256)
constexpr int AT_SMEM_LIMIT = 227 * 1024;   // hard cap; <= 113 KB keeps two CTAs per SM (N <= 960 with the column-sum buffer)
constexpr float AT_RESCALE_LOG2 = 64.0f;  // online softmax: the reference max is only raised past 2^64 (then the tile is redone)

#ifdef TPAT_ATTN_TRACE
#define ATTN_TRACE(slot) do { if (tracing && trace_n < 120) p.trace[trace_n++] = clock64() - t_start + ((long long)(slot) << 48); } while (0)
#else
#define ATTN_TRACE(slot) do { } while (0)
#endif

struct AttnTcParams {
  long long* trace;   // debug only (TPAT_ATTN_TRACE builds): clock stamps of one softmax thread
  float* score_partial;
  float* lse;         // training: [B, H, N] natural-log sum of exp(scale * s) per query row, or NULL
  int score_mode;
  int N, H, num_extra, n_qt, nb, qt_offset;
  int desc;          // 1 = clips are visited from the last one down (see g_walk_desc)
  int lo_off;        // SPLIT: column distance between the hi and lo planes of q / k in the plane buffer (2 * H * 64)
  float scale_log2;  // scale * log2(e)
};

// bf16 row segment (32 probabilities) -> 128B-swizzled K-major tile row
__device__ __forceinline__ void store_p_half(uint8_t* p_row, int hf, int r_local, const float (&v)[32]) {
#pragma unroll
  for (int g = 0; g < 4; ++g)
    *reinterpret_cast<uint4*>(p_row + (((hf * 4 + g) ^ (r_local & 7)) * 16)) =
        make_uint4(pack_bf16x2(v[g * 8 + 0], v[g * 8 + 1]), pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]),
                   pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]), pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]));
}

// SPLIT (score blocks of the "bf16+score32" precision mode, two-pass tiles only): Q and K arrive as split-bf16 planes
// (hi = bf16(x), lo = bf16(x - hi); tmap_q / tmap_kv cover the plane buffer [B][N][q_hi k_hi | q_lo k_lo], tmap_v the
// ordinary qkv buffer) and S = Q_hi K_hi^T + Q_hi K_lo^T + Q_lo K_hi^T: three tcgen05.mma per K = 16 step instead of one,
// i.e. scores exact to ~2^-16 relative instead of 2^-8, on the 3 of 12 blocks whose scores decide which tokens survive.
template <bool TWO_PASS, bool SPLIT>
__global__ void __launch_bounds__(AT_THREADS, SPLIT ? 1 : AT_CTAS_PER_SM)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                    const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_o,
                    const AttnTcParams p) {
  constexpr int Q_BYTES = SPLIT ? 2 * AT_Q_BYTES : AT_Q_BYTES;        // hi (, lo) query tiles; the hi tile later stages O
  constexpr int SLOT_BYTES = SPLIT ? 2 * AT_KV_BYTES : AT_KV_BYTES;   // K tile: hi (, lo); V tile: first 8 KB
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment computed as an OFFSET from the __shared__ array so that the compiler keeps the shared
  // address space (a round trip through uintptr_t turns every staging access into a generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* q_s = smem;                                   // 16 KB
  uint8_t* kv_s = q_s + Q_BYTES;                         // AT_SLOTS x 8 KB (16 KB when SPLIT)
  uint64_t* bars = reinterpret_cast<uint64_t*>(kv_s + AT_SLOTS * SLOT_BYTES);
  uint64_t* q_full = bars;                 // [1]
  uint64_t* kv_full = bars + 1;            // [SLOTS]
  uint64_t* kv_empty = kv_full + AT_SLOTS; // [SLOTS]
  uint64_t* s_full = kv_empty + AT_SLOTS;  // [2] (AT_SBUF used)
  uint64_t* s_empty = s_full + 2;          // [2]
  uint64_t* p_full = s_empty + 2;          // [2]
  uint64_t* p_empty = p_full + 2;          // [2]
  uint64_t* o_full = p_empty + 2;          // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);
  float2* pair_s = reinterpret_cast<float2*>(bars + 32);   // [2][128]: row statistics exchanged between partner warps
  int* flag_s = reinterpret_cast<int*>(pair_s + 2 * AT_BM);  // [2 tile parities][4 quarters][2 halves]: 'reference max must rise'
  float* colsum_s = reinterpret_cast<float*>(flag_s + 16);   // [4][nb*64] when COLMEAN

  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x + p.qt_offset, h = blockIdx.y;
  const int b = p.desc ? (int)(gridDim.z - 1 - blockIdx.z) : (int)blockIdx.z;
  const int q0 = qt * AT_BM;
  const int nb = p.nb;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_q);
    ptx::prefetch_tensormap(&tmap_kv);
    ptx::prefetch_tensormap(&tmap_o);
  }
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(q_full, 1);
    for (int s = 0; s < AT_SLOTS; ++s) { ptx::mbar_init(&kv_full[s], 1); ptx::mbar_init(&kv_empty[s], 1); }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&s_full[i], 1); ptx::mbar_init(&s_empty[i], 8);
      ptx::mbar_init(&p_full[i], 8); ptx::mbar_init(&p_empty[i], 1);
    }
    ptx::mbar_init(o_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc<AT_TMEM_COLS>(tmem_slot);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + AT_SBUF * AT_BK;
  const uint32_t tmem_p = tmem_o + AT_HD;
  pdl_wait();   // everything above touched only on-chip state; global memory from here on

  const int col_q = h * AT_HD, col_k = (p.H + h) * AT_HD, col_v = (2 * p.H + h) * AT_HD;

  if (warp == 0) {
    // ===== TMA producer =====
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(q_full, Q_BYTES);
      ptx::tma_load_3d(q_s, &tmap_q, q_full, col_q, q0, b);
      if (SPLIT) ptx::tma_load_3d(q_s + AT_Q_BYTES, &tmap_q, q_full, p.lo_off + col_q, q0, b);
      int slot = 0; uint32_t phase = 0;
      auto load_tile = [&](int col, int key0) {
        ptx::mbar_wait(&kv_empty[slot], phase ^ 1);
        const bool is_v = col == col_v;
        ptx::mbar_arrive_expect_tx(&kv_full[slot], (SPLIT && !is_v) ? 2 * AT_KV_BYTES : AT_KV_BYTES);
        ptx::tma_load_3d(kv_s + slot * SLOT_BYTES, is_v ? &tmap_v : &tmap_kv, &kv_full[slot], col, key0, b);
        if (SPLIT && !is_v) ptx::tma_load_3d(kv_s + slot * SLOT_BYTES + AT_KV_BYTES, &tmap_kv, &kv_full[slot], p.lo_off + col, key0, b);
        if (++slot == AT_SLOTS) { slot = 0; phase ^= 1; }
      };
      if (TWO_PASS)
        for (int j = 0; j < nb; ++j) load_tile(col_k, j * AT_BK);    // pass 1: K_0 .. K_{nb-1}
      // main pass, in the order the MMA warp consumes them: K_0, K_1, V_0, K_2, V_1, ..., V_{nb-1}
      load_tile(col_k, 0);
      for (int j = 0; j < nb; ++j) {
        if (j + 1 < nb) load_tile(col_k, (j + 1) * AT_BK);
        load_tile(col_v, j * AT_BK);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (ptx::elect_one()) {
      constexpr uint32_t idesc_s = ptx::idesc_bf16_f32(128, AT_BK, 0, 0);  // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_o = ptx::idesc_bf16_f32(128, AT_HD, 0, 1);  // P (K-major) x V (MN-major)
      int slot = 0; uint32_t phase = 0;
      int sidx = 0;  // running S-tile counter: buffer = sidx % AT_SBUF, use count = sidx / AT_SBUF
      const uint64_t q_desc = ptx::smem_desc_sw128(ptx::smem_u32(q_s), 16, 1024);
      auto issue_s = [&]() {
        const int sb = sidx % AT_SBUF;
        ptx::mbar_wait(&kv_full[slot], phase);
        // Main pass of the two-pass tiles, from its third block on: S(j) reuses the buffer of S(j-2), which every softmax
        // warp had read before it arrived on p_full(j-2) -- and this thread waited for that before P(j-2).V(j-2), i.e.
        // before it got here.  The softmax warps therefore do not signal s_empty in the main pass at all; the first two
        // main-pass tiles wait for the last two pass-1 reads.
        if (!TWO_PASS || sidx < nb + 2) ptx::mbar_wait(&s_empty[sb], ((sidx / AT_SBUF) & 1) ^ 1);
        ptx::tc_fence_after();
        const uint64_t k_desc = ptx::smem_desc_sw128(ptx::smem_u32(kv_s + slot * SLOT_BYTES), 16, 1024);
#pragma unroll
        for (int k = 0; k < AT_HD / 16; ++k)
          ptx::mma_f16_ss(tmem_base + sb * AT_BK, q_desc + (uint64_t)(2 * k), k_desc + (uint64_t)(2 * k), idesc_s, k != 0);
        if (SPLIT) {
          // the two cross terms; lo tiles sit 16 KB (Q) / 8 KB (K) behind the hi tiles: descriptor address field is >> 4
          constexpr uint64_t q_lo = AT_Q_BYTES >> 4, k_lo = AT_KV_BYTES >> 4;
#pragma unroll
          for (int k = 0; k < AT_HD / 16; ++k)
            ptx::mma_f16_ss(tmem_base + sb * AT_BK, q_desc + (uint64_t)(2 * k), k_desc + k_lo + (uint64_t)(2 * k), idesc_s, 1);
#pragma unroll
          for (int k = 0; k < AT_HD / 16; ++k)
            ptx::mma_f16_ss(tmem_base + sb * AT_BK, q_desc + q_lo + (uint64_t)(2 * k), k_desc + (uint64_t)(2 * k), idesc_s, 1);
        }
        ptx::tc_commit(&kv_empty[slot]);
        ptx::tc_commit(&s_full[sb]);
        if (++slot == AT_SLOTS) { slot = 0; phase ^= 1; }
        ++sidx;
      };
      ptx::mbar_wait(q_full, 0);
      if (TWO_PASS)
        for (int j = 0; j < nb; ++j) issue_s();          // pass 1
      issue_s();                                         // main pass: S(0)
      for (int j = 0; j < nb; ++j) {
        if (j + 1 < nb) issue_s();                       // S(j+1) overlaps the softmax of block j
        const int pb = j & 1;
        ptx::mbar_wait(&kv_full[slot], phase);           // V_j
        ptx::mbar_wait(&p_full[pb], (j >> 1) & 1);       // P_j written by the softmax warps
        ptx::tc_fence_after();
        const uint32_t v_addr = ptx::smem_u32(kv_s + slot * SLOT_BYTES);
        const int valid = min(AT_BK, p.N - j * AT_BK);   // keys of this block that exist
        const int ksteps = (valid + 15) >> 4;            // P is zero beyond `valid`, V rows beyond N are zero-filled
        for (int k = 0; k < ksteps; ++k) {
          // A: P from TMEM, 16 keys = 8 columns;  B (MN-major): 16 keys = two 8-row groups of 1024 B
          const uint64_t b_desc = ptx::smem_desc_sw128(v_addr + k * 2048, 16, 1024);
          ptx::mma_f16_ts(tmem_o, tmem_p + pb * AT_P_COLS + k * 8, b_desc, idesc_o, (j | k) != 0);
        }
        ptx::tc_commit(&kv_empty[slot]);
        ptx::tc_commit(&p_empty[pb]);
        if (++slot == AT_SLOTS) { slot = 0; phase ^= 1; }
      }
      ptx::tc_commit(o_full);
    }
  } else {
    // ===== softmax / epilogue warps: 8 warps, TMEM lane quarter = warp % 4, thread = (query row, 32-key half) =====
    // Two warps share each 32-row quarter and split every 64-key tile by columns, so an SM sub-partition holds
    // four softmax warps (two per resident CTA): twice the latency hiding around the MUFU.EX2 bursts for the same
    // number of exponentials.  In the single-pass mode both partners read the FULL score row and reduce its max
    // (a few FMNMX on the ALU pipe): the lazy-rescale decision below is then a pure function of identical data and
    // the partners take it together without exchanging a word.
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r_local = quarter * 32 + lane;
    const int row = q0 + r_local;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const float c = p.scale_log2;
    float m_run = -INFINITY, l_run = 0.f;
    int sidx = 0;
    // A warp whose 32 query rows all lie beyond N (tail tile of e.g. N = 513: one valid row in 128) only keeps the
    // barrier protocol going: no TMEM reads, no exps, no P / O writes.  Its P rows stay whatever is in smem and its
    // O rows are garbage, but rows are independent and rows >= N are never stored.
    const bool warp_live = q0 + quarter * 32 < p.N;
    float2* my_x = pair_s + half * AT_BM + r_local;
    const float2* other_x = pair_s + (half ^ 1) * AT_BM + r_local;
#ifdef TPAT_ATTN_TRACE
    const bool tracing = p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 3 && blockIdx.z == (gridDim.z >> 1) && threadIdx.x == 64;
    int trace_n = 0;
    long long t_start = 0;
    if (tracing) { t_start = clock64(); }
    ATTN_TRACE(1);
#endif
    auto release_s = [&](int sb) {
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&s_empty[sb]);
    };
    auto mask_tail = [&](int v, uint32_t (&r)[32]) {     // columns >= v of this half do not exist: -inf
      if (v < 32) {
#pragma unroll
        for (int i = 0; i < 32; ++i) if (i >= v) r[i] = 0xff800000u;
      }
    };
    auto max32 = [&](const uint32_t (&r)[32]) {
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(r[i])); mx1 = fmaxf(mx1, __uint_as_float(r[i + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(r[i + 2])); mx3 = fmaxf(mx3, __uint_as_float(r[i + 3]));
      }
      return fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
    };
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;\n" ::"r"(2 + quarter) : "memory"); };

    if (TWO_PASS) {
      // ---- pass 1: max and sum of exp over this thread's columns, merged with the partner's at the end ----
      for (int j = 0; j < nb; ++j, ++sidx) {
        const int sb = sidx % AT_SBUF;
        ptx::mbar_wait(&s_full[sb], (sidx / AT_SBUF) & 1);
        ptx::tc_fence_after();
        const int vh = p.N - j * AT_BK - half * 32;      // valid columns in this thread's half (may be <= 0)
        uint32_t r[32];
        if (warp_live && vh > 0) {
          ptx::tmem_ld_32x32b_x32(tmem_base + lane_off + sb * AT_BK + half * 32, r);
          ptx::tmem_ld_wait();
        }
        release_s(sb);
        if (!warp_live || vh <= 0) continue;
        mask_tail(vh, r);
        const float mx = fmaxf(m_run, max32(r));
        const float mc = mx * c;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          a0 += ptx::ex2_ftz(fmaf(__uint_as_float(r[i]), c, -mc));
          a1 += ptx::ex2_ftz(fmaf(__uint_as_float(r[i + 1]), c, -mc));
          a2 += ptx::ex2_ftz(fmaf(__uint_as_float(r[i + 2]), c, -mc));
          a3 += ptx::ex2_ftz(fmaf(__uint_as_float(r[i + 3]), c, -mc));
        }
        l_run = l_run * ptx::ex2_ftz((m_run - mx) * c) + ((a0 + a1) + (a2 + a3));
        m_run = mx;
      }
      if (warp_live) {                                   // both partner warps are live together
        *my_x = make_float2(m_run, l_run);
        pair_sync();
        const float2 o = *other_x;
        const float m = fmaxf(m_run, o.x);               // finite: key 0 always exists
        l_run = l_run * ptx::ex2_ftz((m_run - m) * c) + o.y * ptx::ex2_ftz((o.x - m) * c);
        m_run = m;
        pair_sync();                                     // pair_s is reused for the row sums below
      }
    }
    // exponent offset: p = 2^(s*c - off).  Two-pass tiles fold log2(l) in (normalised probabilities).
    float off = TWO_PASS ? fmaf(m_run, c, __log2f(l_run)) : 0.f;
    if (TWO_PASS && p.lse != nullptr && half == 0 && warp_live && row < p.N)
      p.lse[((size_t)b * p.H + h) * p.N + row] = off * 0.69314718055994531f;
    const float row_w = (row >= p.num_extra && row < p.N) ? 1.0f : 0.f;
    const bool mask_rows = q0 + quarter * 32 < p.num_extra || q0 + quarter * 32 + 32 > p.N;
    const bool cls_writer = TWO_PASS && (p.score_mode == TPAT_SCORE_CLS_ROW) && (row == 0);
    float* colsum_w = colsum_s + (size_t)quarter * nb * AT_BK;
    float l2a = 0.f, l2b = 0.f, l2c = 0.f, l2d = 0.f;
    // ---- main pass: probabilities, score partials, P -> smem ----
    for (int j = 0; j < nb; ++j, ++sidx) {
      const int sb = sidx % AT_SBUF, pb = j & 1;
      ATTN_TRACE(2);
      ptx::mbar_wait(&s_full[sb], (sidx / AT_SBUF) & 1);
      ptx::tc_fence_after();
      ATTN_TRACE(3);
      const int valid = p.N - j * AT_BK;                 // > 0
      const int vh = valid - half * 32;                  // valid columns in this thread's half (may be <= 0)
      if (!warp_live) {
        if (TWO_PASS && p.score_mode == TPAT_SCORE_COLMEAN)     // this warp's rows contribute nothing to the column sums
          colsum_w[j * AT_BK + half * 32 + lane] = 0.f;
        if (!TWO_PASS) release_s(sb);                    // (fence::before_thread_sync + __syncwarp inside)
        else { ptx::tc_fence_before(); __syncwarp(); }
        if (lane == 0) ptx::mbar_arrive(&p_full[pb]);
        continue;
      }
      uint32_t r[32];
      if (vh > 0) { ptx::tmem_ld_32x32b_x32(tmem_base + lane_off + sb * AT_BK + half * 32, r); ptx::tmem_ld_wait(); }
      if (!TWO_PASS) release_s(sb);                      // (two-pass main pass: p_full doubles as "S read", see issue_s)
      if (vh > 0) mask_tail(vh, r);
      ATTN_TRACE(4);
      if (!TWO_PASS) {
        // Online softmax with a LAZY reference max: tile 0 fixes m_ref = its row max; later tiles compute their
        // probabilities against the current m_ref straight away (their own max is evaluated off the critical
        // path) and only if some row's max exceeds m_ref by more than 2^64 is the tile redone after raising m_ref
        // and rescaling O (TMEM) and the running sum.  Probabilities stay <= 2^64: exact in bf16 / fp32 range.
        if (j == 0) {                                    // the row max of tile 0: one exchange with the partner
          const float mx_own = vh > 0 ? max32(r) : -INFINITY;
          my_x->x = mx_own;
          pair_sync();
          m_run = fmaxf(mx_own, other_x->x);
        }
        off = m_run * c;
      }
      ATTN_TRACE(5);
      // P buffer pb was last read by PV(j-2), which the MMA thread issued BEFORE S(j); tcgen05 operations retire
      // in issue order and s_full(j) is a commit of everything issued before it, so the buffer is already free.
      const uint32_t p_tm = tmem_p + lane_off + pb * AT_P_COLS + half * 16;   // this thread's 32 keys = 16 columns
      const float l2a_in = l2a, l2b_in = l2b, l2c_in = l2c, l2d_in = l2d;
      auto emit_half = [&]() {
        const int col0 = j * AT_BK + half * 32;
        if (vh <= 0) {
          if (half * 32 < ((valid + 15) & ~15)) {        // still inside the MMA's K range: zero it
            uint32_t z[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) z[i] = 0u;
            ptx::tmem_st_32x32b_x16(p_tm, z);
          }
          return;
        }
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = ptx::ex2_ftz(fmaf(__uint_as_float(r[i]), c, -off));   // -inf -> 0
        if (!TWO_PASS) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) { l2a += v[i]; l2b += v[i + 1]; l2c += v[i + 2]; l2d += v[i + 3]; }
        }
        {
          uint32_t pk[16];                               // bf16 pairs: the K-major A operand of P.V, straight into TMEM
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
          ptx::tmem_st_32x32b_x16(p_tm, pk);
        }
        if (TWO_PASS) {
          if (p.score_mode == TPAT_SCORE_COLMEAN) {
            // column sums over this warp's 32 rows: butterfly transpose-reduce, lane i ends with column col0+i
            if (mask_rows) {                              // warp-uniform: only warps that hold a cls row or rows >= N
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] *= row_w;
            }
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) {
              const bool upper = (lane & o) != 0;
#pragma unroll
              for (int i = 0; i < o; ++i) {
                const float send = upper ? v[i] : v[i + o];
                const float keep = upper ? v[i + o] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
              }
            }
            colsum_w[col0 + lane] = v[0];
          } else if (cls_writer) {
            float* dst = p.score_partial + ((size_t)b * p.H + h) * p.N;
            for (int i = 0; i < 32; ++i)
              if (col0 + i < p.N) dst[col0 + i] = v[i];
          }
        }
      };
      emit_half();
      ATTN_TRACE(12);
      if (!TWO_PASS && j > 0) {
        // does any row of this quarter need a higher reference?  Each partner knows its own columns only: the
        // warp-level verdicts are swapped through shared memory (slot = tile parity) around a 64-thread barrier
        const float mx_own = vh > 0 ? max32(r) : -INFINITY;
        const bool need_w = __any_sync(0xffffffffu, (mx_own - m_run) * c > AT_RESCALE_LOG2);
        if (lane == 0) flag_s[(j & 1) * 8 + quarter * 2 + half] = need_w ? 1 : 0;
        ATTN_TRACE(13);
        pair_sync();
        ATTN_TRACE(14);
        if (need_w || flag_s[(j & 1) * 8 + quarter * 2 + (half ^ 1)] != 0) {
          my_x->x = mx_own;
          pair_sync();
          const float mx = fmaxf(mx_own, other_x->x);    // the full row's max: same value in both partners
          pair_sync();
          const bool need = (mx - m_run) * c > AT_RESCALE_LOG2;
          // rare: raise m_ref, rescale what was accumulated before this tile, redo the tile
          const float f = need ? ptx::ex2_ftz((m_run - mx) * c) : 1.0f;
          if (need) m_run = mx;
          off = m_run * c;
          l2a = l2a_in * f; l2b = l2b_in * f; l2c = l2c_in * f; l2d = l2d_in * f;
          // every PV issued so far (up to block j-1) must have retired before O is touched
          ptx::mbar_wait(&p_empty[(j - 1) & 1], ((j - 1) >> 1) & 1);
          ptx::tc_fence_after();
          {
            uint32_t o0[32];                             // this thread's 32 of the row's 64 output columns
            ptx::tmem_ld_32x32b_x32(tmem_o + lane_off + half * 32, o0);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o0[i] = __float_as_uint(__uint_as_float(o0[i]) * f);
            ptx::tmem_st_32x32b_x32(tmem_o + lane_off + half * 32, o0);
          }
          ptx::tmem_st_wait();
          ptx::tc_fence_before();
          emit_half();
        }
      }
      ATTN_TRACE(7);
      ptx::tmem_st_wait();             // P is in tensor memory before the MMA thread is told so
      ATTN_TRACE(6);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&p_full[pb]);
      ATTN_TRACE(8);
    }
    // the row sum is split between the partners
    float o_scale = 1.0f;
    if (!TWO_PASS && warp_live) {
      const float l_own = (l2a + l2b) + (l2c + l2d);
      my_x->y = l_own;
      pair_sync();
      const float l_tot = l_own + other_x->y;
      o_scale = 1.0f / l_tot;
      if (p.lse != nullptr && half == 0 && row < p.N)
        p.lse[((size_t)b * p.H + h) * p.N + row] = fmaf(m_run, c, __log2f(l_tot)) * 0.69314718055994531f;
    }
    // ---- epilogue: O (TMEM) -> bf16 -> swizzled smem tile -> one TMA store ----
    ATTN_TRACE(9);
    ptx::mbar_wait(o_full, 0);         // every PV MMA retired: the P buffers are free as well
    ptx::tc_fence_after();
    ATTN_TRACE(10);
    if (warp_live) {
      uint32_t r0[32];
      ptx::tmem_ld_32x32b_x32(tmem_o + lane_off + half * 32, r0);
      ptx::tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r0[i]) * o_scale;
      store_p_half(q_s + r_local * 128, half, r_local, v);     // Q is dead: its tile stages O
    }
    ptx::fence_proxy_async_smem();
    asm volatile("bar.sync 1, 256;\n" ::: "memory");
    if (warp == 2 && lane == 0) {
      ptx::tma_store_3d(&tmap_o, q_s, h * AT_HD, q0, b);   // rows >= N are clipped by the tensor map
      ptx::tma_store_commit();
    }
    if (TWO_PASS && p.score_mode == TPAT_SCORE_COLMEAN) {
      // sum the four quarters' column sums in a fixed order and publish this tile's partial row
      float* dstp = p.score_partial + ((size_t)b * p.H * p.n_qt + (size_t)h * p.n_qt + qt) * p.N;
      const int ldc = nb * AT_BK;
      for (int jcol = threadIdx.x - 64; jcol < p.N; jcol += 256)
        dstp[jcol] = ((colsum_s[jcol] + colsum_s[ldc + jcol]) + colsum_s[2 * ldc + jcol]) + colsum_s[3 * ldc + jcol];
    }
    if (warp == 2 && lane == 0) ptx::tma_store_wait_read<0>();   // smem must outlive the bulk store's reads
    ATTN_TRACE(11);
#ifdef TPAT_ATTN_TRACE
    if (tracing) p.trace[127] = trace_n;
#endif
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<AT_TMEM_COLS>(tmem_base);
  }
}

#ifdef TPAT_ATTN_TRACE
long long* g_attn_trace_buf = nullptr;
extern "C" int tpat_debug_attn_trace(long long* host_out) {   // debug builds only: copy the 128 stamps to the host
  if (!g_attn_trace_buf) return 1;
  return cudaMemcpy(host_out, g_attn_trace_buf, 256 * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : 2;
}
#endif

int attention_tc_qtiles(int N) { return (N + AT_BM - 1) / AT_BM; }

// attention_tc3.cu: two query tiles per CTA, one thread per row, 128-key blocks (tiles that feed no importance score)
int attention_tc3(const void* qkv, void* out, int B, int N, int H, float scale, int qt_offset, float* lse, cudaStream_t st);
static bool use_v3() {
  const char* e = getenv("TPAT_ATTN_V3");      // read per call so that tests can A/B both kernels
  return e != nullptr && e[0] == '1';
}
// attention_tc4.cu: same tiling as the single-pass instantiation above, but the two 32-key halves of a row are fully
// independent (own reference max, row sum and output accumulator; no partner exchange inside the key loop)
int attention_tc4(const void* qkv, void* out, int B, int N, int H, float scale, int qt_offset, float* lse, cudaStream_t st);
// attention_tc5.cu: the same kernel made persistent (two resident CTAs per SM walk the (clip, head, tile) items)
int attention_tc5(const void* qkv, void* out, int B, int N, int H, float scale, int qt_offset, float* lse, cudaStream_t st);
static bool use_v5() {
  const char* e = getenv("TPAT_ATTN_V5");
  return e != nullptr && e[0] == '1';
}
static bool use_v4() {
  const char* e = getenv("TPAT_ATTN_V4");      // default ON since r02aa (x1.05 - 1.16 on the single-pass tiles); "0" = the kernel above
  return e == nullptr || e[0] != '0';
}

template <bool TWO_PASS, bool SPLIT = false>
static int launch_attn(const CUtensorMap& tq, const CUtensorMap& tkv, const CUtensorMap& tv, const CUtensorMap& to,
                       const AttnTcParams& p, dim3 grid, size_t smem, cudaStream_t st) {
  static DeviceOnce once;
  auto kern = attention_tc_kernel<TWO_PASS, SPLIT>;
  if (once.first()) {
    TPAT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM_LIMIT));
    TPAT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    if (getenv("TPAT_DEBUG")) {
      int nblk = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nblk, kern, AT_THREADS, smem);
      fprintf(stderr, "[tpat] attention_tc_kernel<%d>: %d CTAs/SM at %zu B smem\n", (int)TWO_PASS, nblk, smem);
    }
    once.mark();
  }
  TPAT_CUDA(launch_kernel(kern, dim3(grid), dim3(AT_THREADS), smem, st, tq, tkv, tv, to, p));
  TPAT_LAUNCH_CHECK();
  return 0;
}

int attention_tc(const void* qkv, void* out, float* score_partial, int score_mode, int B, int N, int H,
                 int num_extra, float scale, cudaStream_t st, const void* qk_planes, float* lse) {
  TPAT_CHECK(N <= 4096, "tpat_attention(tc): N=%d too large (max 4096)", N);
  TPAT_CHECK(qk_planes == nullptr || score_mode != TPAT_SCORE_NONE, "tpat_attention(tc): split q / k planes are for score blocks only");
  CUtensorMap tm_q, tm_kv, tm_o, tm_qs, tm_ks;
  if (qk_planes != nullptr) {
    if (int rc = encode_tmap_3d_qkv(&tm_qs, qk_planes, B, N, 4 * H * AT_HD, AT_BM)) return rc;
    if (int rc = encode_tmap_3d_qkv(&tm_ks, qk_planes, B, N, 4 * H * AT_HD, AT_BK)) return rc;
  }
  if (int rc = encode_tmap_3d_qkv(&tm_q, qkv, B, N, 3 * H * AT_HD, AT_BM)) return rc;
  if (int rc = encode_tmap_3d_qkv(&tm_kv, qkv, B, N, 3 * H * AT_HD, AT_BK)) return rc;
  if (int rc = encode_tmap_3d_qkv(&tm_o, out, B, N, H * AT_HD, AT_BM)) return rc;
  AttnTcParams p;
  p.trace = nullptr;
#ifdef TPAT_ATTN_TRACE
  { static long long* dbg = nullptr; if (!dbg) { cudaMalloc(&dbg, 256 * sizeof(long long)); } cudaMemsetAsync(dbg, 0, 256 * sizeof(long long), st); p.trace = dbg;
    extern long long* g_attn_trace_buf; g_attn_trace_buf = dbg; }
#endif
  p.score_partial = score_partial;
  p.lse = lse;
  p.score_mode = score_mode;
  p.N = N; p.H = H; p.num_extra = num_extra;
  p.n_qt = attention_tc_qtiles(N);
  p.nb = (N + AT_BK - 1) / AT_BK;
  p.qt_offset = 0;
  p.desc = g_walk_desc;
  p.lo_off = 2 * H * AT_HD;
  p.scale_log2 = scale * 1.4426950408889634f;
  const size_t split_extra = qk_planes ? (size_t)AT_Q_BYTES + AT_SLOTS * AT_KV_BYTES : 0;
  const size_t base_smem = 1024 + AT_Q_BYTES + AT_SLOTS * AT_KV_BYTES + 256 + 2 * AT_BM * sizeof(float2) + 64;
  const size_t colsum_bytes = (size_t)4 * p.nb * AT_BK * sizeof(float);
  TPAT_CHECK(base_smem + (score_mode == TPAT_SCORE_COLMEAN ? colsum_bytes : 0) <= (size_t)AT_SMEM_LIMIT,
             "tpat_attention(tc): N=%d needs %zu bytes of shared memory", N, base_smem + colsum_bytes);
  TPAT_CHECK(base_smem + split_extra + (score_mode == TPAT_SCORE_COLMEAN ? colsum_bytes : 0) <= (size_t)AT_SMEM_LIMIT,
             "tpat_attention(tc, split): N=%d needs too much shared memory", N);
  if (score_mode == TPAT_SCORE_COLMEAN) {     // every tile contributes normalised column sums
    if (qk_planes) return launch_attn<true, true>(tm_qs, tm_ks, tm_kv, tm_o, p, dim3(p.n_qt, H, B), base_smem + split_extra + colsum_bytes, st);
    return launch_attn<true>(tm_q, tm_kv, tm_kv, tm_o, p, dim3(p.n_qt, H, B), base_smem + colsum_bytes, st);
  }
  if (score_mode == TPAT_SCORE_CLS_ROW) {     // only the tile holding query row 0 must normalise
    if (qk_planes) { if (int rc = launch_attn<true, true>(tm_qs, tm_ks, tm_kv, tm_o, p, dim3(1, H, B), base_smem + split_extra, st)) return rc; }
    else if (int rc = launch_attn<true>(tm_q, tm_kv, tm_kv, tm_o, p, dim3(1, H, B), base_smem, st)) return rc;
    if (p.n_qt == 1) return 0;
    if (use_v3()) return attention_tc3(qkv, out, B, N, H, scale, 1, lse, st);
    if (use_v5()) return attention_tc5(qkv, out, B, N, H, scale, 1, lse, st);
    if (use_v4()) return attention_tc4(qkv, out, B, N, H, scale, 1, lse, st);
    p.qt_offset = 1;
    return launch_attn<false>(tm_q, tm_kv, tm_kv, tm_o, p, dim3(p.n_qt - 1, H, B), base_smem, st);
  }
  if (use_v3()) return attention_tc3(qkv, out, B, N, H, scale, 0, lse, st);
  if (use_v5()) return attention_tc5(qkv, out, B, N, H, scale, 0, lse, st);
  if (use_v4()) return attention_tc4(qkv, out, B, N, H, scale, 0, lse, st);
  return launch_attn<false>(tm_q, tm_kv, tm_kv, tm_o, p, dim3(p.n_qt, H, B), base_smem, st);
}

}  // namespace tpat
