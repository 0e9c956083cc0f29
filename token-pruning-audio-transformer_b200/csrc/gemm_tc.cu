// tcgen05 GEMM for sm_100a: C = epilogue(A W^T + bias), bf16 operands, fp32 accumulation in TMEM.
//
// Replaces cuBLASLt + the unfused elementwise kernels behind nn.Linear in the reference's block
// (audiomae/models_vit.py:41-45 fc1/GELU/fc2, :76 qkv, :96,198 proj + residual, :205 MLP residual)
// and the patch-embed conv restated as a GEMM (:246, pos add :358).
//
// Design (one CTA per SM, persistent over output tiles, 192 threads):
//   warp 0   TMA producer: cp.async.bulk.tensor loads of the A [128 x 64] and W [256 x 64] bf16
//            k-blocks into a 4-stage 128B-swizzled shared-memory ring (48 KB / stage)
//   warp 1   MMA issuer: one elected thread issues tcgen05.mma.cta_group::1.kind::f16
//            (M=128, N=256, K=16) x 4 per k-block; tcgen05.commit frees the stage / publishes
//            the accumulator
//   warps 2-5 epilogue: tcgen05.ld the 128 x 256 fp32 accumulator (thread = row), apply
//            bias / GELU / residual / pos-embed, store.  Two accumulators (2 x 256 TMEM columns)
//            so the epilogue of tile i overlaps the main loop of tile i+1.
// Tiles are walked n-fastest so the CTAs in flight share A k-blocks through L2.
// Roofline: tensor pipe.  Algorithmic FLOPs = 2*M*N*K per launch.
#include "gemm_tc_common.cuh"

#include <cstdlib>
#include <mutex>
#include <unordered_map>

namespace tpat {

// ---------------- tensor-map encoding (driver entry point, cached) ----------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      cudaGetLastError();
  });
  return fn;
}

struct TmapKey {
  const void* ptr; uint64_t rows, cols, pitch; uint32_t box_rows, box_cols; int elem_bytes; bool sw;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && pitch == o.pitch && box_rows == o.box_rows &&
           box_cols == o.box_cols && elem_bytes == o.elem_bytes && sw == o.sw;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    auto mix = [&](uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.rows); mix(k.cols); mix(k.pitch); mix(k.box_rows); mix(k.box_cols); mix((uint64_t)k.elem_bytes * 2 + k.sw);
    return h;
  }
};

int encode_tmap_2d(CUtensorMap* out, const void* gptr, int elem_bytes, uint64_t rows, uint64_t cols, uint64_t pitch_bytes,
                   uint32_t box_rows, uint32_t box_cols, bool swizzle128) {
  static std::mutex mu;
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  TmapKey key{gptr, rows, cols, pitch_bytes, box_rows, box_cols, elem_bytes, swizzle128};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return 0; }
  }
  EncodeTiledFn fn = get_encode_fn();
  TPAT_CHECK(fn != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  TPAT_CHECK(aligned16(gptr) && pitch_bytes % 16 == 0, "TMA needs a 16-byte aligned base and row pitch (pitch=%llu)", (unsigned long long)pitch_bytes);
  TPAT_CHECK(box_rows <= 256 && box_cols <= 256, "TMA box dims must be <= 256");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = fn(out, dt, 2, const_cast<void*>(gptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TPAT_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu pitch=%llu box=%ux%u)", (int)r,
             (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)pitch_bytes, box_rows, box_cols);
  {
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() > 8192) cache.clear();
    cache.emplace(key, *out);
  }
  return 0;
}

// 2-D bf16 map with a 32-column x 32-row box and 64-byte swizzle: the output blocks of the TMA-store epilogue (gemm_tc2.cu)
int encode_tmap_2d_c32(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols, uint64_t pitch_bytes) {
  static std::mutex mu;
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  TmapKey key{gptr, rows, cols, pitch_bytes, 32, 32, 2, true};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return 0; }
  }
  EncodeTiledFn fn = get_encode_fn();
  TPAT_CHECK(fn != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  TPAT_CHECK(aligned16(gptr) && pitch_bytes % 16 == 0, "TMA needs a 16-byte aligned base and row pitch (pitch=%llu)", (unsigned long long)pitch_bytes);
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {pitch_bytes};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(gptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TPAT_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (32 x 32 bf16, 64B swizzle) failed with CUresult %d (rows=%llu cols=%llu pitch=%llu)", (int)r,
             (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)pitch_bytes);
  {
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() > 8192) cache.clear();
    cache.emplace(key, *out);
  }
  return 0;
}

// 3-D map over a token-major activation [B][N][ld] (bf16): box = 64 columns x box_rows rows x 1 clip, 128B swizzle.
// Rows >= N of a clip are out of bounds of dim 1 and are zero-filled (never the next clip's rows).
int encode_tmap_3d_qkv(CUtensorMap* out, const void* gptr, int B, int N, int ld, int box_rows) {
  static std::mutex mu;
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  TmapKey key{gptr, (uint64_t)N, (uint64_t)ld, (uint64_t)B, (uint32_t)box_rows, 64, 2, true};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return 0; }
  }
  EncodeTiledFn fn = get_encode_fn();
  TPAT_CHECK(fn != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  TPAT_CHECK(aligned16(gptr) && (ld * 2) % 16 == 0, "TMA needs a 16-byte aligned base and row pitch");
  cuuint64_t gdim[3] = {(cuuint64_t)ld, (cuuint64_t)N, (cuuint64_t)B};
  cuuint64_t gstride[2] = {(cuuint64_t)ld * 2, (cuuint64_t)N * ld * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(gptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TPAT_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(3d) failed with CUresult %d (B=%d N=%d ld=%d)", (int)r, B, N, ld);
  {
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() > 8192) cache.clear();
    cache.emplace(key, *out);
  }
  return 0;
}

// Generic 3-D map over a [B][N][ld] tensor (clip, token, channel): box = {box_cols, box_rows, 1}, 128-byte swizzle.
// Rows >= N of a box are zero-filled on load and dropped on store / reduce: a tile never touches the next clip.
int encode_tmap_3d(CUtensorMap* out, const void* gptr, int elem_bytes, int B, int N, int ld, int box_rows, int box_cols) {
  static std::mutex mu;
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  TmapKey key{gptr, (uint64_t)N, (uint64_t)ld, (uint64_t)B, (uint32_t)box_rows, (uint32_t)box_cols, elem_bytes + 16, true};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return 0; }
  }
  EncodeTiledFn fn = get_encode_fn();
  TPAT_CHECK(fn != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  TPAT_CHECK(aligned16(gptr) && ((size_t)ld * elem_bytes) % 16 == 0 && box_cols * elem_bytes == 128, "TMA (3d) needs a 16-byte aligned base / pitch and a 128-byte box row");
  cuuint64_t gdim[3] = {(cuuint64_t)ld, (cuuint64_t)N, (cuuint64_t)B};
  cuuint64_t gstride[2] = {(cuuint64_t)ld * elem_bytes, (cuuint64_t)N * ld * elem_bytes};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = fn(out, dt, 3, const_cast<void*>(gptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TPAT_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(3d) failed with CUresult %d (B=%d N=%d ld=%d)", (int)r, B, N, ld);
  {
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() > 8192) cache.clear();
    cache.emplace(key, *out);
  }
  return 0;
}

// ---------------- kernel ----------------
constexpr int TG_STAGES = 4;
constexpr int TG_EPI_WARPS = 8;
constexpr int TG_STAGING_BYTES = tg_staging_bytes(TG_EPI_WARPS);
constexpr int TG_THREADS = tg_threads(TG_EPI_WARPS);
constexpr int TG_A_BYTES = TG_BM * TG_BK * 2;   // 16 KB
constexpr int TG_B_BYTES = TG_BN * TG_BK * 2;   // 32 KB
constexpr int TG_STAGE_BYTES = TG_A_BYTES + TG_B_BYTES;
constexpr int TG_SMEM_BYTES = TG_STAGES * TG_STAGE_BYTES + TG_STAGING_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

template <int EPI, typename OutT>
__global__ void __launch_bounds__(TG_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, const TcGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment computed as an OFFSET from the __shared__ array so that the compiler keeps the shared
  // address space (a round trip through uintptr_t turns every staging access into a generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + TG_STAGES * TG_A_BYTES;
  uint8_t* staging = smem + TG_STAGES * TG_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + TG_STAGING_BYTES);
  uint64_t* full_bar = bars;                    // [STAGES]
  uint64_t* empty_bar = bars + TG_STAGES;       // [STAGES]
  uint64_t* acc_full = bars + 2 * TG_STAGES;    // [2]
  uint64_t* acc_empty = acc_full + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.tiles_m * p.tiles_n;
  const int nkb = p.K / TG_BK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_a);
    ptx::prefetch_tensormap(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < TG_STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&acc_full[a], 1); ptx::mbar_init(&acc_empty[a], TG_EPI_WARPS); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc<512>(tmem_slot);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything above touched only on-chip state; global memory from here on

  if (warp == 0) {
    // ===== TMA producer =====
    if (ptx::elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = tc_tile_m(p, tile) * TG_BM, n0 = (tile % p.tiles_n) * p.bn;
        for (int kb = 0; kb < nkb; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          ptx::mbar_arrive_expect_tx(&full_bar[stage], TG_A_BYTES + p.bn * TG_BK * 2);
          ptx::tma_load_2d(smem_a + stage * TG_A_BYTES, &tmap_a, &full_bar[stage], kb * TG_BK, m0);
          if (p.w_kn) {        // W[K][N]: bn / 64 boxes of 64 k-rows x 64 n-columns, 8 KB apart
            for (int j = 0; j < p.bn / 64; ++j)
              ptx::tma_load_2d(smem_b + stage * TG_B_BYTES + j * 8192, &tmap_w, &full_bar[stage], n0 + 64 * j, kb * TG_BK);
          } else {
            ptx::tma_load_2d(smem_b + stage * TG_B_BYTES, &tmap_w, &full_bar[stage], kb * TG_BK, n0);
          }
          if (++stage == TG_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::idesc_bf16_f32(TG_BM, p.bn, 0, p.w_kn ? 1 : 0);
      const uint32_t b_lbo = p.w_kn ? 8192u : 16u;        // MN-major B: 64-column chunks 8 KB apart, 16 k-rows = 2048 B per step
      const uint64_t b_kstep = p.w_kn ? 128u : 2u;
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        ptx::mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * TG_BN;
        for (int kb = 0; kb < nkb; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint64_t a_desc = ptx::smem_desc_sw128(ptx::smem_u32(smem_a + stage * TG_A_BYTES), 16, 1024);
          const uint64_t b_desc = ptx::smem_desc_sw128(ptx::smem_u32(smem_b + stage * TG_B_BYTES), b_lbo, 1024);
#pragma unroll
          for (int k = 0; k < TG_BK / TG_UMMA_K; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128 B swizzle atom: +2 in the (addr >> 4) field
            ptx::mma_f16_ss(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + b_kstep * (uint64_t)k, idesc, (kb | k) != 0);
          }
          ptx::tc_commit(&empty_bar[stage]);   // stage reusable once these MMAs retire
          if (++stage == TG_STAGES) { stage = 0; phase ^= 1; }
        }
        ptx::tc_commit(&acc_full[acc]);        // accumulator complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===== epilogue warps 2..9: TMEM lane quarter = warp % 4; the two warps of a quarter split the chunks =====
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;
    uint8_t* stg = staging + (warp - 2) * 4096;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = tc_tile_m(p, tile) * TG_BM + q * 32, n0 = (tile % p.tiles_n) * p.bn;
      TcEpiPrefetch<TG_EPI_WARPS> pf;
      tc_epilogue_prefetch<TG_EPI_WARPS>(p, n0, cg, lane, pf);
      ptx::mbar_wait(&acc_full[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * TG_BN;
      uint64_t* rel_bar = &acc_empty[acc];
      tc_epilogue_tile<EPI, OutT, TG_EPI_WARPS>(p, taddr_row, m0, n0, cg, stg, lane, pf, [&]() { if (lane == 0) ptx::mbar_arrive(rel_bar); });
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

template <int EPI, typename OutT>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tw, const TcGemmParams& p, cudaStream_t st) {
  static DeviceOnce once;
  auto kern = gemm_tc_kernel<EPI, OutT>;
  if (once.first()) { TPAT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM_BYTES)); once.mark(); }
  const int tiles = p.tiles_m * p.tiles_n;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  TPAT_CUDA(launch_kernel(kern, dim3(grid), dim3(TG_THREADS), TG_SMEM_BYTES, st, ta, tw, p));
  TPAT_LAUNCH_CHECK();
  return 0;
}

int gemm_tc(const void* A, int lda, const void* W, void* C, int c_dtype, int ldc, int M, int N, int K,
            const EpiParams& ep, cudaStream_t st) {
  TPAT_CHECK(K % TG_BK == 0 && K >= TG_BK, "tpat_gemm(tc): need K %% 64 == 0 (K=%d)", K);
  TPAT_CHECK(N % 32 == 0, "tpat_gemm(tc): need N %% 32 == 0 (N=%d)", N);
  TPAT_CHECK(aligned16(C) && (ldc * dtype_size(c_dtype)) % 16 == 0, "tpat_gemm(tc): C must be 16-byte aligned with a 16-byte multiple pitch");
  TPAT_CHECK(ep.bias == nullptr || aligned16(ep.bias), "tpat_gemm(tc): bias must be 16-byte aligned");
  if (ep.epilogue == TPAT_EPI_BIAS_RESIDUAL)
    TPAT_CHECK(c_dtype == TPAT_F32 && ep.residual && aligned16(ep.residual) && ep.ldr % 4 == 0, "tpat_gemm(tc): residual epilogue needs fp32 C and an aligned fp32 residual");
  if (ep.epilogue == TPAT_EPI_BIAS_POS)
    TPAT_CHECK(c_dtype == TPAT_F32 && ep.pos && aligned16(ep.pos) && ep.P > 0, "tpat_gemm(tc): pos epilogue needs fp32 C, pos and P");
  {
    // CTA-pair kernel (cta_group::2) unless TPAT_GEMM_2CTA=0; read per call so tests can A/B both kernels
    const char* e = getenv("TPAT_GEMM_2CTA");
    const bool two_cta = e == nullptr || e[0] != '0';
    // Small batches (M <= 3072 rows, i.e. up to ~6 clips of 513 tokens): a 256 x 256 tile per CTA pair leaves most SMs
    // idle and its K loop alone is 3 us; the 1-CTA kernel with 128 x 128 / 128 x 64 tiles spreads the work over the whole
    // chip (measured r01g, CUDA-graph replay: B = 1 0.95 -> 0.72 ms, B = 4 1.01 -> 0.85 ms; TPAT_GEMM_SMALL_M moves the
    // threshold, TPAT_GEMM_NO_SMALL=1 disables it).
    static const bool small_off = getenv("TPAT_GEMM_NO_SMALL") != nullptr;
    static const int small_m = [] { const char* t = getenv("TPAT_GEMM_SMALL_M"); return t ? atoi(t) : 3072; }();
    const bool small = !small_off && M <= small_m && ep.xb == nullptr && ep.ln_part == nullptr;
    if (two_cta && sm_count() >= 2 && !small) return gemm_tc2(A, lda, W, C, c_dtype, ldc, M, N, K, ep, st);
  }
  int bn = TG_BN;
  if (M <= 8192) {
    const int tm = (M + TG_BM - 1) / TG_BM;
    if (tm * ((N + 255) / 256) < 100 && N % 128 == 0) bn = 128;
    if (tm * ((N + 127) / 128) < 100 && N % 64 == 0) bn = 64;
  }
  CUtensorMap ta, tw;
  if (int rc = encode_tmap_2d(&ta, A, 2, (uint64_t)M, (uint64_t)K, (uint64_t)lda * 2, TG_BM, TG_BK, true)) return rc;
  if (ep.w_kn) { if (int rc = encode_tmap_2d(&tw, W, 2, (uint64_t)K, (uint64_t)N, (uint64_t)N * 2, 64, 64, true)) return rc; }
  else if (int rc = encode_tmap_2d(&tw, W, 2, (uint64_t)N, (uint64_t)K, (uint64_t)K * 2, bn, TG_BK, true)) return rc;
  TPAT_CHECK(ep.xb == nullptr && ep.ln_part == nullptr, "tpat_gemm_ln: the LayerNorm fold needs the CTA-pair kernel (TPAT_GEMM_2CTA=0 is set)");
  TcGemmParams p{};
  p.M = M; p.N = N; p.K = K; p.C = C; p.ldc = ldc; p.bias = ep.bias; p.residual = ep.residual; p.ldr = ep.ldr;
  p.pos = ep.pos; p.P = ep.P; p.num_extra = ep.num_extra;
  p.dact_out = ep.dact_out; p.ld_dact = ep.ld_dact; p.aux = ep.aux; p.ld_aux = ep.ld_aux; p.row_scale = ep.row_scale; p.rows_per_clip = ep.rows_per_clip;
  p.bn = bn;
  p.w_kn = ep.w_kn;
  p.cs_part = ep.cs_part;
  p.tiles_m = (M + TG_BM - 1) / TG_BM; p.tiles_n = (N + bn - 1) / bn;
  p.desc = g_walk_desc;
  p.debug_skip = 0;
  switch (ep.epilogue) {
    case TPAT_EPI_BIAS:
      return c_dtype == TPAT_BF16 ? launch_tc<TPAT_EPI_BIAS, __nv_bfloat16>(ta, tw, p, st) : launch_tc<TPAT_EPI_BIAS, float>(ta, tw, p, st);
    case TPAT_EPI_BIAS_GELU:
      return c_dtype == TPAT_BF16 ? launch_tc<TPAT_EPI_BIAS_GELU, __nv_bfloat16>(ta, tw, p, st) : launch_tc<TPAT_EPI_BIAS_GELU, float>(ta, tw, p, st);
    case TPAT_EPI_DGELU:
      return c_dtype == TPAT_BF16 ? launch_tc<TPAT_EPI_DGELU, __nv_bfloat16>(ta, tw, p, st) : launch_tc<TPAT_EPI_DGELU, float>(ta, tw, p, st);
    case TPAT_EPI_BIAS_RESIDUAL:
      return launch_tc<TPAT_EPI_BIAS_RESIDUAL, float>(ta, tw, p, st);
    case TPAT_EPI_BIAS_POS:
      return launch_tc<TPAT_EPI_BIAS_POS, float>(ta, tw, p, st);
  }
  set_error("tpat_gemm(tc): bad epilogue %d", ep.epilogue);
  return 1;
}

}  // namespace tpat
