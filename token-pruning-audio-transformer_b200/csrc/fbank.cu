// Kaldi-compatible log-mel filterbank front end on the GPU: waveform -> normalised spectrogram, the tensor the
// patch-embed kernel consumes.
//
// Replaces the eval-time input pipeline of the reference data loaders (audiomae/dataset.py:175-230,298;
// ast/src/dataloader.py:98-149,204): `waveform - waveform.mean()`, torchaudio.compliance.kaldi.fbank(htk_compat=True,
// use_energy=False, window_type='hanning', num_mel_bins=128, dither=0.0, frame_shift=10) -- i.e. per 25 ms frame:
// remove the DC offset, pre-emphasis 0.97 (replicate padding), Hann window, zero-pad to 512, |rfft|^2, triangular
// mel filters, log(max(., eps)) -- then pad with the clip's minimum / crop to target_length frames and
// (x - norm_mean) / (2 norm_std).  The window and the mel filter table are built by the host mirror
// (tpat/frontend.py) with the same fp32 formulae torchaudio uses and passed in as tables.
//
// One warp per frame: samples -> shared memory, frame mean by shuffle, pre-emphasis + window, the 512-point real
// transform as a 256-point complex radix-4 Stockham FFT in shared memory (4 stages, 2 butterflies per lane and stage,
// __syncwarp between stages) + even/odd recombination, power spectrum, then every lane accumulates 4 mel bins over the
// non-zero range of their filters.  HBM-bound in principle (64 clips of 10.24 s: 42 MB in, 33.5 MB out, 22 us);
// in practice instruction / latency bound (~1.3 k warp instructions per frame).
#include "common.cuh"

#include <math.h>

namespace tpat {

constexpr int FB_WARPS = 8;
constexpr int FB_MEAN_CHUNKS = 8;   // partial sums per clip of the clip-mean pass (summed in a fixed order)
constexpr int FB_WS_FLOATS = 32;    // scratch per clip: 8 doubles (partial sums) + 1 uint (minimum log-mel, ordered key)

__device__ __forceinline__ unsigned fkey(float v) { const unsigned u = __float_as_uint(v); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float funkey(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// wave [B, L]; lengths[b] samples are valid (NULL: L); clip_sum: FB_MEAN_CHUNKS partial sums per clip (NULL: no
// clip-mean subtraction).  out [B, out_frames, n_mel]: rows f < n_frames(b) receive the RAW log-mel energies; the rest
// is left to fbank_finalize_kernel.
// The NFFT-point real transform is done as an NFFT/2-point complex one (z[n] = x[2n] + i x[2n+1]) with a radix-4
// Stockham autosort FFT (natural order in and out, ping-pong between two shared-memory buffers, HALF = 4^STAGES or
// 2 * 4^STAGES with one radix-2 stage in front) followed by the usual even/odd recombination.
template <int NFFT>
__global__ void __launch_bounds__(32 * FB_WARPS)
fbank_kernel(const float* __restrict__ wave, int L, const int* __restrict__ lengths, float* __restrict__ ws, int use_clip_mean,
             const float* __restrict__ window, int win, int shift, float preemph,
             const float* __restrict__ mel, const int* __restrict__ mel_start, const int* __restrict__ mel_len, int n_mel,
             float* __restrict__ out, int out_frames) {
  pdl_trigger();
  pdl_wait();
  constexpr int HALF = NFFT / 2;                               // complex FFT length
  constexpr bool ODD = (HALF == 128 || HALF == 512);           // odd power of two: one radix-2 stage first
  extern __shared__ float2 fb_smem[];
  float2* tw = fb_smem;                                        // [HALF]   e^{-2 pi i k / HALF}
  float2* tw2 = tw + HALF;                                     // [HALF]   e^{-2 pi i k / NFFT}
  float2* bufa = tw2 + HALF + (threadIdx.x >> 5) * (2 * HALF); // this warp's two work buffers
  float2* bufb = bufa + HALF;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  for (int k = threadIdx.x; k < HALF; k += blockDim.x) {
    float s, c;
    sincospif(-2.0f * (float)k / (float)HALF, &s, &c);
    tw[k] = make_float2(c, s);
    sincospif(-2.0f * (float)k / (float)NFFT, &s, &c);
    tw2[k] = make_float2(c, s);
  }
  __syncthreads();
  const int len = lengths ? min(lengths[b], L) : L;
  const int n_frames = len >= win ? min(1 + (len - win) / shift, out_frames) : 0;
  float cm = 0.f;
  if (use_clip_mean && len > 0) {
    const double* clip_sum = reinterpret_cast<const double*>(ws + (size_t)b * FB_WS_FLOATS);
    double t = 0.0;
    for (int i = 0; i < FB_MEAN_CHUNKS; ++i) t += clip_sum[i];
    cm = (float)(t / (double)len);
  }
  unsigned* min_key = reinterpret_cast<unsigned*>(ws + (size_t)b * FB_WS_FLOATS + 2 * FB_MEAN_CHUNKS);
  const float* wb = wave + (size_t)b * L;
  float* stage = reinterpret_cast<float*>(bufb);               // the raw frame (win <= NFFT floats = HALF float2)
  float* power = reinterpret_cast<float*>(bufb);               // later: HALF + 1 power values

  for (int f = blockIdx.x * FB_WARPS + warp; f < n_frames; f += gridDim.x * FB_WARPS) {
    const float* src = wb + (size_t)f * shift;
    float s = 0.f;
    for (int j = lane; j < win; j += 32) {
      const float v = __ldg(src + j) - cm;
      stage[j] = v;
      s += v;
    }
    s = warp_sum(s);
    const float fmean = s / (float)win;                        // remove_dc_offset (kaldi.py:_get_window)
    __syncwarp();
    // pre-emphasis with replicate padding, window, zero padding; packed as z[n] = y[2n] + i y[2n+1]
    auto sample = [&](int j) {
      if (j >= win) return 0.f;
      const float cur = stage[j] - fmean;
      const float prev = stage[j > 0 ? j - 1 : 0] - fmean;
      return (cur - preemph * prev) * __ldg(window + j);
    };
    for (int n = lane; n < HALF; n += 32) bufa[n] = make_float2(sample(2 * n), sample(2 * n + 1));
    __syncwarp();
    float2* in = bufa;
    float2* outb = bufb;
    int Ns = 1;
    if (ODD) {                                                 // radix-2 Stockham stage (Ns = 1: no twiddles)
      for (int j = lane; j < HALF / 2; j += 32) {
        const float2 a = in[j], c = in[j + HALF / 2];
        outb[2 * j] = make_float2(a.x + c.x, a.y + c.y);
        outb[2 * j + 1] = make_float2(a.x - c.x, a.y - c.y);
      }
      __syncwarp();
      float2* t = in; in = outb; outb = t;
      Ns = 2;
    }
#pragma unroll 1
    for (; Ns < HALF; Ns *= 4) {                               // radix-4 Stockham stages
      const int tstep = HALF / (Ns * 4);
      for (int j = lane; j < HALF / 4; j += 32) {
        const int k = j & (Ns - 1);
        float2 v0 = in[j], v1 = in[j + HALF / 4], v2 = in[j + HALF / 2], v3 = in[j + 3 * HALF / 4];
        if (Ns > 1) {
          v1 = cmul(v1, tw[k * tstep]);
          v2 = cmul(v2, tw[2 * k * tstep]);
          v3 = cmul(v3, tw[3 * k * tstep]);
        }
        const float2 a0 = make_float2(v0.x + v2.x, v0.y + v2.y), a1 = make_float2(v0.x - v2.x, v0.y - v2.y);
        const float2 a2 = make_float2(v1.x + v3.x, v1.y + v3.y), a3 = make_float2(v1.x - v3.x, v1.y - v3.y);
        const int d = ((j - k) << 2) + k;                      // (j / Ns) * Ns * 4 + j % Ns
        outb[d] = make_float2(a0.x + a2.x, a0.y + a2.y);
        outb[d + Ns] = make_float2(a1.x + a3.y, a1.y - a3.x);  // a1 - i a3
        outb[d + 2 * Ns] = make_float2(a0.x - a2.x, a0.y - a2.y);
        outb[d + 3 * Ns] = make_float2(a1.x - a3.y, a1.y + a3.x);  // a1 + i a3
      }
      __syncwarp();
      float2* t = in; in = outb; outb = t;
    }
    // `in` holds Z = FFT_HALF(z).  X[k] = (Z[k] + conj Z[HALF-k]) / 2 - i e^{-2 pi i k / NFFT} (Z[k] - conj Z[HALF-k]) / 2
    float pw[(HALF + 1 + 31) / 32];
#pragma unroll
    for (int i = 0; i < (HALF + 1 + 31) / 32; ++i) {
      const int k = lane + 32 * i;
      float v = 0.f;
      if (k <= HALF) {
        const float2 zk = in[k & (HALF - 1)], zc = in[(HALF - k) & (HALF - 1)];
        const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y));       // even part
        const float2 o = make_float2(0.5f * (zk.x - zc.x), 0.5f * (zk.y + zc.y));       // (Z[k] - conj Z[HALF-k]) / 2
        const float2 w = k < HALF ? tw2[k] : make_float2(-1.f, 0.f);
        const float2 t = cmul(w, o);                                                     // times -i: (t.y, -t.x)
        const float re = e.x + t.y, im = e.y - t.x;
        v = re * re + im * im;
      }
      pw[i] = v;
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < (HALF + 1 + 31) / 32; ++i) {
      const int k = lane + 32 * i;
      if (k <= HALF) power[k] = pw[i];
    }
    __syncwarp();
    // mel filters: lane m, m + 32, ... ; only the non-zero span of each triangle
    float* dst = out + ((size_t)b * out_frames + f) * n_mel;
    float fmin_lane = INFINITY;
    for (int m = lane; m < n_mel; m += 32) {
      const int k0 = mel_start[m], n = mel_len[m];
      const float* wrow = mel + (size_t)m * (HALF + 1) + k0;
      float acc = 0.f;
      for (int k = 0; k < n; ++k) acc = fmaf(power[k0 + k], __ldg(wrow + k), acc);
      const float lm = logf(fmaxf(acc, 1.1920928955078125e-07f));   // max(., FLT_EPSILON).log()
      dst[m] = lm;
      fmin_lane = fminf(fmin_lane, lm);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) fmin_lane = fminf(fmin_lane, __shfl_xor_sync(0xffffffffu, fmin_lane, o));
    if (lane == 0) atomicMin(min_key, fkey(fmin_lane));        // the clip's minimum (order independent: deterministic)
    __syncwarp();
  }
}

// partial sums of each clip (fixed chunking, double accumulation): grid (FB_MEAN_CHUNKS, B)
__global__ void __launch_bounds__(1024)
wave_sum_kernel(const float* __restrict__ wave, int L, const int* __restrict__ lengths, float* __restrict__ ws, int do_sum) {
  __shared__ double red[32];
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  if (blockIdx.x == 0 && threadIdx.x == 0)
    *reinterpret_cast<unsigned*>(ws + (size_t)b * FB_WS_FLOATS + 2 * FB_MEAN_CHUNKS) = 0xffffffffu;   // +max key
  if (!do_sum) return;
  double* clip_sum = reinterpret_cast<double*>(ws + (size_t)b * FB_WS_FLOATS);
  const int len = lengths ? min(lengths[b], L) : L;
  const int per = (len + FB_MEAN_CHUNKS - 1) / FB_MEAN_CHUNKS;
  const int lo = blockIdx.x * per, hi = min(len, lo + per);
  const float* wb = wave + (size_t)b * L;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;               // short fp32 chains (<= 20 terms each), then double
  int i = lo + threadIdx.x;
  for (; i + 3 * 1024 < hi; i += 4 * 1024) {
    s0 += __ldg(wb + i); s1 += __ldg(wb + i + 1024); s2 += __ldg(wb + i + 2048); s3 += __ldg(wb + i + 3072);
  }
  for (; i < hi; i += 1024) s0 += __ldg(wb + i);
  double s = ((double)s0 + (double)s1) + ((double)s2 + (double)s3);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 32; ++w) t += red[w];
    clip_sum[blockIdx.x] = t;
  }
}

// Frames >= n_frames(b) are filled with the clip's minimum log-mel value (dataset.py:214-221), then every value
// becomes (v - norm_mean) / (2 norm_std) (dataset.py:298).  In place on spec [B, T, n_mel]; grid (chunks, B).
__global__ void __launch_bounds__(256)
fbank_finalize_kernel(float* __restrict__ spec, int L, const int* __restrict__ lengths, const float* __restrict__ ws, int win,
                      int shift, int T, int n_mel, float norm_mean, float inv_2std) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const int len = lengths ? min(lengths[b], L) : L;
  const int n_frames = len >= win ? min(1 + (len - win) / shift, T) : 0;
  float4* sb = reinterpret_cast<float4*>(spec + (size_t)b * T * n_mel);
  const int n_valid = n_frames * n_mel / 4, n_all = T * n_mel / 4;          // n_mel % 4 == 0 (checked by the caller)
  const float mn = n_frames > 0 ? funkey(*reinterpret_cast<const unsigned*>(ws + (size_t)b * FB_WS_FLOATS + 2 * FB_MEAN_CHUNKS)) : 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_all; i += gridDim.x * blockDim.x) {
    float4 v = i < n_valid ? sb[i] : make_float4(mn, mn, mn, mn);
    v.x = (v.x - norm_mean) * inv_2std; v.y = (v.y - norm_mean) * inv_2std;
    v.z = (v.z - norm_mean) * inv_2std; v.w = (v.w - norm_mean) * inv_2std;
    sb[i] = v;
  }
}

}  // namespace tpat

extern "C" int tpat_fbank(const float* wave, const int* lengths, int B, int L, int subtract_clip_mean, float* clip_mean_ws,
                          const float* window, int win, int shift, int nfft, float preemph, const float* mel,
                          const int* mel_start, const int* mel_len, int n_mel, float* spec, int T, float norm_mean,
                          float norm_std, tpat_stream_t stream) {
  using namespace tpat;
  TPAT_CHECK(wave && window && mel && mel_start && mel_len && spec, "tpat_fbank: null pointer");
  TPAT_CHECK(B > 0 && L > 0 && T > 0 && n_mel > 0 && n_mel <= 1024, "tpat_fbank: bad sizes B=%d L=%d T=%d n_mel=%d", B, L, T, n_mel);
  TPAT_CHECK(win >= 2 && win <= nfft && shift > 0, "tpat_fbank: need 2 <= win <= nfft and shift > 0 (win=%d shift=%d nfft=%d)", win, shift, nfft);
  TPAT_CHECK(nfft == 256 || nfft == 512 || nfft == 1024, "tpat_fbank: nfft must be 256, 512 or 1024 (got %d)", nfft);
  TPAT_CHECK(clip_mean_ws && (reinterpret_cast<uintptr_t>(clip_mean_ws) & 7u) == 0, "tpat_fbank: needs an 8-byte aligned workspace of 32 * B floats");
  TPAT_CHECK(n_mel % 4 == 0 && aligned16(spec), "tpat_fbank: n_mel must be a multiple of 4 and spec 16-byte aligned");
  TPAT_CHECK(norm_std != 0.f, "tpat_fbank: zero norm_std");
  cudaStream_t st = as_stream(stream);
  float* ws = clip_mean_ws;
  const int use_mean = subtract_clip_mean ? 1 : 0;
  TPAT_CUDA(launch_kernel(wave_sum_kernel, dim3(use_mean ? FB_MEAN_CHUNKS : 1, B), dim3(1024), 0, st, wave, L, lengths, ws, use_mean));
  const int max_frames = L >= win ? 1 + (L - win) / shift : 0;
  const int frames = max_frames < T ? max_frames : T;
  if (frames > 0) {
    const dim3 grid((frames + FB_WARPS - 1) / FB_WARPS, B);
    const size_t smem = ((size_t)nfft + (size_t)FB_WARPS * nfft) * sizeof(float2);   // 2 twiddle tables + 2 buffers per warp
#define TPAT_FBANK_CASE(N)                                                                                         \
  case N: {                                                                                                            \
    static DeviceOnce once;                                                                                            \
    if (once.first()) { TPAT_CUDA(cudaFuncSetAttribute(fbank_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); once.mark(); } \
    TPAT_CUDA(launch_kernel(fbank_kernel<N>, dim3(grid), dim3(32 * FB_WARPS), smem, st, wave, L, lengths, ws, use_mean, window, win, shift, \
                            preemph, mel, mel_start, mel_len, n_mel, spec, T));                                        \
  } break;
    switch (nfft) {
      TPAT_FBANK_CASE(256)
      TPAT_FBANK_CASE(512)
      TPAT_FBANK_CASE(1024)
    }
#undef TPAT_FBANK_CASE
  }
  const int fin_chunks = (T * n_mel / 4 + 256 * 8 - 1) / (256 * 8);
  TPAT_CUDA(launch_kernel(fbank_finalize_kernel, dim3(fin_chunks, B), dim3(256), 0, st, spec, L, lengths, (const float*)ws, win, shift, T,
                          n_mel, norm_mean, 0.5f / norm_std));
  TPAT_LAUNCH_CHECK();
  return 0;
}
