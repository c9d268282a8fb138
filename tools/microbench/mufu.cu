// Microbenchmark: MUFU.EX2 / FFMA issue throughput per SM (one warp per SMSP and 2 warps per SMSP).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3f + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) a[i] = ex2(a[i]);
      if (MODE == 1) a[i] = fmaf(a[i], 1.0001f, 0.5f);
      if (MODE == 2) { a[i] = ex2(a[i]); a[i] = fmaf(a[i], 1.0001f, 0.5f); a[i] = a[i] + 1.0f; a[i] = fmaf(a[i], 0.999f, 0.25f);}  // 1 MUFU + 3 FP
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* o; long long* c; cudaMalloc(&o, 1 << 22); cudaMalloc(&c, 8 * 1024);
  const int iters = 2000;
  for (int threads : {128, 256, 512}) {
    for (int mode = 0; mode < 3; ++mode) {
      if (mode == 0) k<0><<<148, threads>>>(o, c, iters);
      if (mode == 1) k<1><<<148, threads>>>(o, c, iters);
      if (mode == 2) k<2><<<148, threads>>>(o, c, iters);
      cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
      double per = double(h[0]) / iters / 16;  // cycles per unrolled element per warp-set
      const char* nm[] = {"MUFU.EX2", "FFMA", "MUFU+3FP"};
      printf("threads/SM %d  %-9s: %.2f cycles per instr-group per warp slot -> %.1f lane-ops/clk/SM\n", threads, nm[mode], per, threads / per);
    }
  }
  return 0;
}
