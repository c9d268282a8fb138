// Microbenchmark: issue throughput per SM of MUFU.EX2, FFMA, F2FP (cvt.rn.f16x2.f32) and the attention exp mix.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float a, float b) { uint32_t r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; }
__device__ __forceinline__ uint32_t ex2h2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float a[16];
  uint32_t u[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3f + i;
#pragma unroll
  for (int i = 0; i < 8; ++i) u[i] = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) a[i] = ex2(a[i]);
      if (MODE == 1) a[i] = fmaf(a[i], 1.0001f, 0.5f);
      if (MODE == 2) { a[i] = ex2(a[i]); a[i] = fmaf(a[i], 1.0001f, 0.5f); a[i] = a[i] + 1.0f; }
    }
    if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { u[i] ^= pack(a[2 * i], a[2 * i + 1]); a[2 * i] += 1.0f; }
    }
    if (MODE == 4) {  // attention mix per pair of elements: 2 FFMA + 2 MUFU + 2 FADD + 1 F2FP
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float e0 = ex2(fmaf(a[2 * i], 0.5f, -1.0f)), e1 = ex2(fmaf(a[2 * i + 1], 0.5f, -1.0f));
        a[2 * i] += e0; a[2 * i + 1] += e1;
        u[i] ^= pack(e0, e1);
      }
    }
    if (MODE == 5) {  // packed-half exp2: 8 MUFU.EX2.F16x2 = 16 elements
#pragma unroll
      for (int i = 0; i < 8; ++i) u[i] = ex2h2(u[i]);
    }
    if (MODE == 6) {  // attention mix with the packed-half exp2: per pair 2 FFMA + 1 F2FP + 1 MUFU.F16x2 + 1 HADD2
      uint32_t acc = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t xh = pack(fmaf(a[2 * i], 0.5f, -1.0f), fmaf(a[2 * i + 1], 0.5f, -1.0f));
        const uint32_t eh = ex2h2(xh);
        asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(acc) : "r"(eh));
        u[i] ^= eh;
      }
      a[0] += __uint_as_float(acc);
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 16; ++i) s += a[i];
  for (int i = 0; i < 8; ++i) s += __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* o; long long* c; cudaMalloc(&o, 1 << 22); cudaMalloc(&c, 8 * 1024);
  const int iters = 2000;
  const char* nm[] = {"MUFU.EX2 x16", "FFMA x16", "(MUFU+FFMA+FADD) x16", "F2FP x8 (+8 FADD)", "attn mix: 16 elem", "MUFU.EX2.F16x2 x8 (16 el)", "attn mix f16x2: 16 elem"};
  for (int threads : {128, 256}) {
    for (int mode = 0; mode < 7; ++mode) {
      if (mode == 0) k<0><<<148, threads>>>(o, c, iters);
      if (mode == 1) k<1><<<148, threads>>>(o, c, iters);
      if (mode == 2) k<2><<<148, threads>>>(o, c, iters);
      if (mode == 3) k<3><<<148, threads>>>(o, c, iters);
      if (mode == 4) k<4><<<148, threads>>>(o, c, iters);
      if (mode == 5) k<5><<<148, threads>>>(o, c, iters);
      if (mode == 6) k<6><<<148, threads>>>(o, c, iters);
      cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
      printf("%d warps/SMSP  %-22s: %.1f cycles per iteration (per warp)\n", threads / 128, nm[mode], double(h[0]) / iters);
    }
  }
  return 0;
}
