// Sustained issue-to-retire rate of the tcgen05.mma forms the attention kernel uses, one CTA per SM:
//   S  : SS form, M=128 N=128 K=16, A and B K-major in 128B-swizzled shared memory          (S = Q K^T)
//   PVt: TS form, M=128 N=64  K=16, A fp16 in tensor memory, B MN-major in shared memory      (O += P V, round 2)
//   PVs: SS form, M=128 N=64  K=16, A K-major smem, B MN-major smem                           (O += P V, round 1)
//   mix: per "key tile" 2 x (4 S + 8 PVt), the order the attention MMA warp issues them
// Prints clocks per MMA and the implied fp16 FLOP/clk/SM.  Operand contents are irrelevant for timing.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/microbench/mma_rate tools/microbench/mma_rate.cu
#include <cstdio>
#include "../../f5_tts_b200/csrc/common.cuh"
using namespace f5;

template <int MODE>
__global__ void __launch_bounds__(128, 1) k(long long* cyc, int reps) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));  // 4 tiles of 16 KB
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int t = threadIdx.x, warp = t >> 5;
  for (int i = t; i < 4 * 16384 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;  // 1.0h everywhere
  if (t == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (t == 0) {
    constexpr uint32_t idesc_s = make_idesc_f16(128, 128, 0, 0);
    constexpr uint32_t idesc_o = make_idesc_f16(128, 64, 0, 1);
    const uint64_t qd = make_smem_desc_sw128(smem_u32(sm));
    const uint64_t kd = make_smem_desc_sw128(smem_u32(sm + 16384));
    const uint64_t vd = make_smem_desc_sw128(smem_u32(sm + 32768));
    const long long t0 = clock64();
    long long n = 0;
    for (int r = 0; r < reps; ++r) {
      if (MODE == 0) {
        for (int i = 0; i < 8; ++i) tc_mma_ss(tm + (i & 1) * 128, qd + uint64_t(2 * (i & 3)), kd + uint64_t(2 * (i & 3)), idesc_s, 1);
        n += 8;
      } else if (MODE == 1) {
        for (int i = 0; i < 8; ++i)
          tc_mma_ts(tm + 256 + (i & 1) * 64, tm + 384 + i * 8, make_smem_desc_sw128(smem_u32(sm + 32768 + i * 2048)), idesc_o, 1);
        n += 8;
      } else if (MODE == 2) {
        for (int i = 0; i < 8; ++i)
          tc_mma_ss(tm + 256 + (i & 1) * 64, qd + uint64_t(2 * (i & 3)), make_smem_desc_sw128(smem_u32(sm + 32768 + i * 2048)), idesc_o, 1);
        n += 8;
      } else {
        for (int w = 0; w < 2; ++w)
          for (int i = 0; i < 4; ++i) tc_mma_ss(tm + w * 128, qd + uint64_t(2 * i), kd + uint64_t(2 * i), idesc_s, i != 0);
        for (int w = 0; w < 2; ++w)
          for (int i = 0; i < 8; ++i)
            tc_mma_ts(tm + 256 + w * 64, tm + 384 + w * 64 + i * 8, make_smem_desc_sw128(smem_u32(sm + 32768 + i * 2048)), idesc_o, 1);
        n += 24;
      }
      (void)vd;
    }
    tc_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    cyc[blockIdx.x * 2] = t1 - t0;
    cyc[blockIdx.x * 2 + 1] = n;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

template <int MODE>
void run(const char* name, double flop_per_mma) {
  long long* cyc;
  cudaMalloc(&cyc, 148 * 16);
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 16384 + 2048);
  for (int grid : {1, 148}) {
    k<MODE><<<grid, 128, 4 * 16384 + 2048>>>(cyc, 64);
    k<MODE><<<grid, 128, 4 * 16384 + 2048>>>(cyc, 64);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[296];
    cudaMemcpy(h, cyc, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost);
    double c = 0, n = 0;
    for (int i = 0; i < grid; ++i) { c += h[2 * i]; n += h[2 * i + 1]; }
    printf("%-34s grid %3d: %7.1f clk per MMA  (%6.0f fp16 FLOP/clk/SM)  %s\n", name, grid, c / n, flop_per_mma * n / c,
           cudaGetErrorString(e));
  }
}

int main() {
  run<0>("S   SS 128x128x16", 2.0 * 128 * 128 * 16);
  run<1>("PVt TS 128x64x16 (A in TMEM)", 2.0 * 128 * 64 * 16);
  run<2>("PVs SS 128x64x16 (B MN-major)", 2.0 * 128 * 64 * 16);
  run<3>("mix 8 S + 16 PVt per key tile", (8 * 2.0 * 128 * 128 * 16 + 16 * 2.0 * 128 * 64 * 16) / 24);
  return 0;
}
