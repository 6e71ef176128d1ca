// Micro-benchmark: tcgen05.ld (32x32b.x32) read throughput per SM as a function of how many warps
// issue it and which TMEM lane quarters they address.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32]) {
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
template <int kDepth>
__global__ void __launch_bounds__(512, 1) k(int nwarps, int iters, unsigned long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < nwarps) {
    const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t v[32], w[32];
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int c = 0; c < 8; c += 2) {
        ld32(base + ((c * 32 + (warp >> 2) * 64) & 511 & ~31), v);
        if (kDepth == 1) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        ld32(base + ((c * 32 + 32 + (warp >> 2) * 64) & 511 & ~31), w);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc ^= v[0] ^ v[31] ^ w[0] ^ w[31];
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (acc == 0x12345u) sink[threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}
int main() {
  unsigned long long* out; uint32_t* sink;
  cudaMalloc(&out, 148 * 8); cudaMalloc(&sink, 4096);
  const int iters = 2000;
  for (int depth = 1; depth <= 2; ++depth)
  for (int nw : {1, 2, 4, 8, 16}) {
    if (depth == 1) { k<1><<<148, 512>>>(nw, iters, out, sink); k<1><<<148, 512>>>(nw, iters, out, sink); }
    else { k<2><<<148, 512>>>(nw, iters, out, sink); k<2><<<148, 512>>>(nw, iters, out, sink); }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    unsigned long long h[148]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    double cyc = (double)h[0];
    double bytes = (double)nw * iters * 8 * 32 * 32 * 4;
    printf("{\"loads_in_flight\": %d, \"warps\": %d, \"cycles\": %.0f, \"bytes_per_clk_per_sm\": %.1f, \"clk_per_ld_x32_per_warp\": %.1f}\n", depth, nw, cyc, bytes / cyc, cyc / (iters * 8));
  }
  return 0;
}
