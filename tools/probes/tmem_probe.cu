// Micro-benchmark: TMEM -> register bandwidth of tcgen05.ld on sm_100a (B200).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tmem_probe.cu && ./tmem_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}

// nwarps warps, each issues `iters` x (UNROLL loads of 32 columns) from its own lane quarter
template <int UNROLL>
__global__ void probe(int iters, int nwarps, long long* out_cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t base = slot;
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  if (warp < nwarps) {
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    __syncwarp();
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      uint32_t r[UNROLL][32];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) ld_x32(base + lane_off + ((i * UNROLL + u) & 15) * 32, r[u]);
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
#pragma unroll
        for (int k = 0; k < 32; ++k) acc ^= r[u][k];
    }
    t1 = clock64();
  }
  if ((threadIdx.x & 31) == 0 && warp < nwarps) out_cycles[blockIdx.x * 32 + warp] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(base) : "memory");
}

int main() {
  long long* d_cyc; uint32_t* d_sink;
  cudaMalloc(&d_cyc, 148 * 32 * sizeof(long long));
  cudaMalloc(&d_sink, 148 * 512 * sizeof(uint32_t));
  const int iters = 2000;
  for (int unroll = 1; unroll <= 4; unroll *= 2) {
    for (int nw : {1, 2, 4, 8, 16}) {
      cudaMemset(d_cyc, 0, 148 * 32 * sizeof(long long));
      const int threads = nw * 32 < 128 ? 128 : nw * 32;
      if (unroll == 1) probe<1><<<148, threads>>>(iters, nw, d_cyc, d_sink);
      else if (unroll == 2) probe<2><<<148, threads>>>(iters, nw, d_cyc, d_sink);
      else probe<4><<<148, threads>>>(iters, nw, d_cyc, d_sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[32];
      cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int w = 0; w < nw; ++w) mx = h[w] > mx ? h[w] : mx;
      const double bytes = (double)iters * unroll * 4096.0 * nw;   // per SM
      printf("unroll %d warps %2d: %8lld cycles  -> %.1f cycles per x32 load per warp, %.1f B/clk/SM\n", unroll, nw, mx,
             (double)mx / (iters * unroll), bytes / mx);
    }
  }
  return 0;
}
