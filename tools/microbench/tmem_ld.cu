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
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(base + ((warp >> 2) * 128 + c * 32)));
      asm volatile("tcgen05.wait::ld.sync.aligned;");
      acc += __uint_as_float(r[it & 31]);
    }
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
    printf("%d threads/SM (%s): %.1f cycles per x32 load per warp (serialised with wait) -> %.1f B/clk/SM  [%s]\n", threads, cudaGetErrorString(e),
           cyc_per_ld, bytes / cyc_per_ld, "tcgen05.ld.32x32b.x32");
  }
  return 0;
}
