// TMEM read bandwidth per SM (tcgen05.ld.32x32b.x32), the quantity that bounds the zone encoder's two accumulator
// drains: W warps of one CTA (warp w reads lane quadrant w % 4) each issue `iters` pairs of 4-KB loads back to back.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/tmem_ld_bench tools/micro/tmem_ld_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}

__global__ void __launch_bounds__(512, 1) bench(int warps, int iters, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot;
  const int warp = threadIdx.x >> 5;
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  if (warp < warps) {
    const uint32_t my = base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64 % 512);
    __syncwarp();
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      uint32_t v0[32], v1[32];
      tmem_ld32(my, v0);
      tmem_ld32(my + 32, v1);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int k = 0; k < 32; ++k) acc ^= v0[k] ^ v1[k];
    }
    t1 = clock64();
  }
  if ((threadIdx.x & 31) == 0 && warp < warps) cycles[blockIdx.x * 16 + warp] = t1 - t0;
  if (acc == 0x12345u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(base) : "memory");
}

int main() {
  long long* d; uint32_t* sink;
  cudaMalloc(&d, 148 * 16 * sizeof(long long)); cudaMalloc(&sink, 4);
  const int iters = 2000;
  for (int warps : {1, 2, 4, 8, 12, 16}) {
    cudaMemset(d, 0, 148 * 16 * sizeof(long long));
    bench<<<148, 512>>>(warps, iters, d, sink);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    long long h[16];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
    const double bytes = (double)warps * iters * 2 * 4096;
    printf("warps %2d: %lld cycles for %d x 2 x 4 KB per warp -> %.1f B/clk per SM, %.1f B/clk per warp, %.0f cycles per pair\n", warps, mx, iters,
           bytes / mx, bytes / mx / warps, (double)mx / iters);
  }
  return 0;
}
