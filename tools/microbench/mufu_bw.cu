// Micro-benchmark: MUFU.EX2 (ex2.approx.ftz.f32) throughput per SM vs. number of resident warps.
#include <cuda_runtime.h>
#include <stdio.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int kMode>
__global__ void k(int iters, unsigned long long* out, float* sink) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = -0.001f * (threadIdx.x + i);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (kMode == 0) a[i] = ex2(a[i]);                       // MUFU only
      else a[i] = ex2(fmaf(a[i], 0.18f, -0.5f)) + a[i] * 0.5f;  // FFMA + MUFU + FMUL + FADD (softmax-like mix)
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 123.456f) sink[threadIdx.x] = s;
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
}
int main() {
  unsigned long long* out; float* sink; cudaMalloc(&out, 148 * 8); cudaMalloc(&sink, 4096 * 4);
  const int iters = 4000;
  for (int mode = 0; mode < 2; ++mode)
  for (int warps : {4, 8, 16, 32}) {
    if (mode == 0) { k<0><<<148, warps * 32>>>(iters, out, sink); k<0><<<148, warps * 32>>>(iters, out, sink); }
    else { k<1><<<148, warps * 32>>>(iters, out, sink); k<1><<<148, warps * 32>>>(iters, out, sink); }
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
    unsigned long long h[148]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    double ops = (double)warps * 32 * iters * 8;
    printf("{\"mode\": \"%s\", \"warps_per_sm\": %d, \"cycles\": %llu, \"ex2_lanes_per_clk_per_sm\": %.2f}\n",
           mode == 0 ? "mufu_only" : "ffma+mufu+fmul+fadd", warps, h[0], ops / (double)h[0]);
  }
  return 0;
}
