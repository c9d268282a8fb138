// Microbenchmark: tcgen05.ld (TMEM -> registers) throughput per SM with 4 / 8 warps issuing 32x32b.x32 loads.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(float* out, long long* cyc, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t r[4][32];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[c][0]), "=r"(r[c][1]), "=r"(r[c][2]), "=r"(r[c][3]), "=r"(r[c][4]), "=r"(r[c][5]), "=r"(r[c][6]), "=r"(r[c][7]), "=r"(r[c][8]),
            "=r"(r[c][9]), "=r"(r[c][10]), "=r"(r[c][11]), "=r"(r[c][12]), "=r"(r[c][13]), "=r"(r[c][14]), "=r"(r[c][15]), "=r"(r[c][16]),
            "=r"(r[c][17]), "=r"(r[c][18]), "=r"(r[c][19]), "=r"(r[c][20]), "=r"(r[c][21]), "=r"(r[c][22]), "=r"(r[c][23]), "=r"(r[c][24]),
            "=r"(r[c][25]), "=r"(r[c][26]), "=r"(r[c][27]), "=r"(r[c][28]), "=r"(r[c][29]), "=r"(r[c][30]), "=r"(r[c][31])
          : "r"(base + ((warp >> 2) * 128 + c * 32)));
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int c = 0; c < 4; ++c) acc += __uint_as_float(r[c][0] ^ r[c][13] ^ r[c][31]);  // static indices: values stay in registers
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512));
}
int main() {
  float* o; long long* c; cudaMalloc(&o, 1 << 22); cudaMalloc(&c, 8 * 1024);
  const int iters = 500;
  for (int threads : {128, 256}) {
    k<<<148, threads>>>(o, c, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
    double cyc_per_ld = double(h[0]) / iters / 4;
    double bytes = threads / 32 * 4096.0;  // per SM per (ld by every warp)
    printf("%d threads/SM (%s): %.1f cycles per x32 load per warp (4 loads in flight per wait) -> %.1f B/clk/SM  [%s]\n", threads, cudaGetErrorString(e),
           cyc_per_ld, bytes / cyc_per_ld, "tcgen05.ld.32x32b.x32");
  }
  return 0;
}
