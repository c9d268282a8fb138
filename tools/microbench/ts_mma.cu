// Layout probe for tcgen05.mma with the A operand in TENSOR MEMORY (TS form), as the attention kernel uses it for
// O += P V:  A[128 x 16] fp16 written with tcgen05.st.32x32b (lane = row, 32-bit column c = elements 2c | 2c+1 << 16),
// B = 16 x 64 fp16 MN-major in 128B-swizzled shared memory (the layout a TMA box [64 cols, rows] lands in).
// A[t][k] = 2^k, B[k][n] = (n == k)  =>  D[t][n] must be 2^n for n < 16 and 0 for n >= 16 if the packing hypothesis holds;
// otherwise the printed row shows which k each output column saw.  Second pass: K = 32 (two MMAs, A advanced by 8 columns,
// B by 16 rows) with A[t][k] = t + 3k (mod 97), B random small ints, checked against a CPU product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/ts_mma tools/microbench/ts_mma.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../f5_tts_b200/csrc/common.cuh"
using namespace f5;

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

__global__ void __launch_bounds__(128, 1) probe(const __half* A, const __half* B, float* D, int K) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sV = smem_raw + (base - smem_u32(smem_raw));  // K rows x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int t = threadIdx.x, warp = t >> 5;
  if (t == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&slot, 128);
  // B[k][n] -> row k, 16-byte chunk (n / 8) ^ (k & 7)
  for (int i = t; i < K * 64; i += 128) {
    const int k = i / 64, n = i % 64;
    *reinterpret_cast<__half*>(sV + k * 128 + (((n >> 3) ^ (k & 7)) << 4) + (n & 7) * 2) = B[k * 64 + n];
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  const uint32_t lane_off = uint32_t(warp * 32) << 16;
  uint32_t pk[16];
  for (int c = 0; c < 16; ++c) {
    pk[c] = 0;
    if (2 * c < K) {
      __half2 h = __halves2half2(A[t * K + 2 * c], A[t * K + 2 * c + 1]);
      pk[c] = *reinterpret_cast<uint32_t*>(&h);
    }
  }
  tmem_st16(tm + 64 + lane_off, pk);  // A at columns [64, 80), D at [0, 64)
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (t == 0) {
    constexpr uint32_t idesc = make_idesc_f16(128, 64, 0, 1);
    for (int kk = 0; kk < K / 16; ++kk) {
      const uint64_t vdesc = make_smem_desc_sw128(smem_u32(sV + kk * 16 * 128));
      tc_mma_ts(tm, tm + 64 + kk * 8, vdesc, idesc, kk != 0);
    }
    tc_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c = 0; c < 2; ++c) {
    uint32_t r[32];
    tmem_ld32(tm + lane_off + c * 32, r);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) D[t * 64 + c * 32 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 128);
}

int main() {
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  for (int pass = 0; pass < 2; ++pass) {
    const int K = pass == 0 ? 16 : 32;
    std::vector<__half> hA(128 * K), hB(K * 64);
    std::vector<float> ref(128 * 64, 0.f), fa(128 * K), fb(K * 64);
    for (int t = 0; t < 128; ++t)
      for (int k = 0; k < K; ++k) fa[t * K + k] = pass == 0 ? float(1 << k) : float((t + 3 * k) % 97);
    for (int k = 0; k < K; ++k)
      for (int n = 0; n < 64; ++n) fb[k * 64 + n] = pass == 0 ? float(n == k) : float((rand() % 7) - 3);
    for (size_t i = 0; i < fa.size(); ++i) hA[i] = __float2half(fa[i]);
    for (size_t i = 0; i < fb.size(); ++i) hB[i] = __float2half(fb[i]);
    for (int t = 0; t < 128; ++t)
      for (int n = 0; n < 64; ++n)
        for (int k = 0; k < K; ++k) ref[t * 64 + n] += fa[t * K + k] * fb[k * 64 + n];
    __half *dA, *dB;
    float* dD;
    cudaMalloc(&dA, hA.size() * 2);
    cudaMalloc(&dB, hB.size() * 2);
    cudaMalloc(&dD, 128 * 64 * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, 128 * 64 * 4);
    probe<<<1, 128, 16384>>>(dA, dB, dD, K);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> out(128 * 64);
    cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0;
    for (size_t i = 0; i < out.size(); ++i) worst = fmax(worst, fabs(out[i] - ref[i]));
    printf("pass %d (K = %d): %s, max |D - ref| = %g  -> packing hypothesis %s\n", pass, K, cudaGetErrorString(e), worst,
           worst == 0 ? "HOLDS" : "FAILS");
    if (worst != 0 || pass == 0) {
      for (int t : {0, 37, 127}) {
        printf("  row %3d:", t);
        for (int n = 0; n < 20; ++n) printf(" %g", out[t * 64 + n]);
        printf("\n");
      }
    }
  }
  return 0;
}
