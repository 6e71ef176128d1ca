// Micro-benchmark: issue-to-completion rate of tcgen05.mma (kind::f16, bf16 operands, cta_group::1, M = 128)
// for the shapes the attention kernels use, one CTA per SM, nothing else running on the SM:
//   SS  N = 256 / 128 / 64, K-major A and B (S = Q K^T)
//   SS  N = 64, MN-major B                   (what P V would be with P in shared memory)
//   TS  N = 64, MN-major B, A from TMEM      (O = P V as the kernels do it)
// and the TS form again while 8 other warps hammer TMEM with tcgen05.ld (what the softmax warps do).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu ; run: ./umma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_k(uint32_t a) {  // K-major SWIZZLE_128B, SBO 1024
  return (uint64_t)((a & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint64_t desc_mn(uint32_t a) {  // MN-major SWIZZLE_128B (64 contiguous N), SBO 1024
  return (uint64_t)((a & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t idesc(int M, int N, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
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

// mode: 0 SS N=256 K-major | 1 SS N=128 | 2 SS N=64 | 3 SS N=64 MN-major B | 4 TS N=64 MN-major B
//       5 = mode 4 with 8 warps doing tcgen05.ld of OTHER columns meanwhile | 6 = mode 0 with the same TMEM load traffic
__global__ void __launch_bounds__(320, 1) k(int mode, int n_mma, int group, unsigned long long* out, uint32_t* sink) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint32_t slot;
  __shared__ uint64_t bar;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u + i;  // finite bf16 values
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    stop = 0;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  uint32_t accx = 0;
  if (warp == 0) {
    if (lane == 0) {
      const uint32_t a = smem_u32(smem), b = a + 16384;
      const int m = (mode == 5) ? 4 : (mode == 6 ? 0 : mode);
      const int N = m == 0 ? 256 : (m == 1 ? 128 : 64);
      const uint32_t id = idesc(128, N, m >= 3 ? 1 : 0);
      uint32_t phase = 0;
      const long long t0 = clock64();
      for (int i = 0; i < n_mma; i += group) {
        for (int j = 0; j < group; ++j) {
          const int ks = (i + j) & 3;
          if (m == 4) mma_ts(tmem + 256, tmem + ks * 8, desc_mn(b + ks * 2048), id, j != 0);
          else if (m == 3) mma_ss(tmem + 256, desc_k(a + ks * 32), desc_mn(b + ks * 2048), id, j != 0);
          else mma_ss(tmem + 256, desc_k(a + ks * 32), desc_k(b + ks * 32), id, j != 0);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok)
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                       : "=r"(ok) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
        phase ^= 1;
      }
      const long long t1 = clock64();
      out[blockIdx.x] = (unsigned long long)(t1 - t0);
      stop = 1;
    }
  } else if ((mode == 5 || mode == 6) && warp >= 2) {
    // TMEM load traffic on columns 0..255 of this warp's lane quarter (the MMA accumulates in 256..)
    const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t v[32];
    int c = 0;
    while (!stop) {
      ld32(base + ((c * 32) & 255), v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      accx ^= v[0] ^ v[31];
      ++c;
    }
  }
  if (accx == 0x12345u) sink[threadIdx.x] = accx;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
  unsigned long long* out; uint32_t* sink;
  cudaMalloc(&out, 148 * 8); cudaMalloc(&sink, 4096);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const char* names[] = {"SS N=256 K-major", "SS N=128 K-major", "SS N=64 K-major", "SS N=64 MN-major B", "TS N=64 MN-major B",
                         "TS N=64 + 8 warps of tcgen05.ld", "SS N=256 + 8 warps of tcgen05.ld"};
  const int n_mma = 4096;
  for (int mode = 0; mode < 7; ++mode)
    for (int group : {4, 16, 64}) {
      k<<<148, 320, 64 * 1024>>>(mode, n_mma, group, out, sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
      unsigned long long h[148];
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      double s = 0; for (int i = 0; i < 148; ++i) s += h[i];
      printf("%-36s commit every %2d MMAs: %7.1f clk per MMA (incl. commit + wait per group)\n", names[mode], group, s / 148 / n_mma);
    }
  return 0;
}
