// Microbenchmark of the attention softmax exp section in isolation (no TMEM / MMA): cycles per 128-score row per warp
// for the shipped path of csrc/attn.cuh (row max + FFMA2 + exponentials) and an experimental fast path (exponentials
// straight from the accumulator, partial-sum checks), with 1 and 2 warps per SM sub-partition.  Scores are re-read from shared memory
// every round (stands in for tcgen05.ld), so nothing can be hoisted out of the loop.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I f5_tts_b200/csrc -o tools/microbench/softmax_mix tools/microbench/softmax_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "attn.cuh"

using namespace f5;

// Experimental fast path (round-2 study, see profiles/README.md): the softmax reference and scale are folded into the
// Q K^T accumulator, so the exponentials read the accumulator registers directly (no FFMA2, no row-max pass).
template <int POLY>
__device__ __forceinline__ void exp_fast32(const uint32_t (&r)[32], int col0, int kv_rem, uint64_t& sum2, uint32_t* pk) {
  if (col0 >= kv_rem) {
#pragma unroll
    for (int i = 0; i < 16; ++i) pk[i] = 0u;
    return;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float e0, e1;
    const int pi = i & 7;
    if (((pi + 1) * POLY) / 8 != (pi * POLY) / 8) {
      ex2_poly2(f2_pack(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), e0, e1);
    } else {
      e0 = ex2_approx(__uint_as_float(r[2 * i]));
      e1 = ex2_approx(__uint_as_float(r[2 * i + 1]));
    }
    sum2 = f2_add(sum2, f2_pack(e0, e1));
    pk[i] = pack_half2(e0, e1);
  }
}

template <int POLY, int MODE>
__global__ void __launch_bounds__(256, 1) k(const float* in, uint32_t* out, long long* cyc, int rounds, float sc) {
  extern __shared__ float srow[];  // [threads][128], thread-major with a 4-float rotation (no bank conflicts for v4)
  float* mine = srow + threadIdx.x * 132;
  for (int i = 0; i < 128; ++i) mine[i] = in[(threadIdx.x * 128 + i) & 4095];
  __syncthreads();
  uint32_t acc = 0;
  float m_run = 3.0f, l_run = 0.f;
  const long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    uint32_t r0[32], r1[32], r2[32], r3[32];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint4 a = *reinterpret_cast<const uint4*>(mine + 4 * i);
      const uint4 b = *reinterpret_cast<const uint4*>(mine + 32 + 4 * i);
      const uint4 c = *reinterpret_cast<const uint4*>(mine + 64 + 4 * i);
      const uint4 d = *reinterpret_cast<const uint4*>(mine + 96 + 4 * i);
      r0[4 * i] = a.x; r0[4 * i + 1] = a.y; r0[4 * i + 2] = a.z; r0[4 * i + 3] = a.w;
      r1[4 * i] = b.x; r1[4 * i + 1] = b.y; r1[4 * i + 2] = b.z; r1[4 * i + 3] = b.w;
      r2[4 * i] = c.x; r2[4 * i + 1] = c.y; r2[4 * i + 2] = c.z; r2[4 * i + 3] = c.w;
      r3[4 * i] = d.x; r3[4 * i + 1] = d.y; r3[4 * i + 2] = d.z; r3[4 * i + 3] = d.w;
    }
    uint64_t sum2 = f2_pack(0.0f, 0.0f);
    uint32_t pa[32], pb[32];
    if (MODE == 0) {  // exact path of the shipped kernel before the reference was folded into the MMA
      const float mx = row_max128(r0, r1, r2, r3);
      m_run = fmaxf(m_run, mx * sc);
      const uint64_t sc2 = f2_pack(sc, sc);
      const uint64_t nms2 = f2_pack(-m_run, -m_run);
      exp_pack32<POLY>(r0, 0, 128, sc2, nms2, sum2, pa);
      exp_pack32<POLY>(r1, 32, 128, sc2, nms2, sum2, pa + 16);
      exp_pack32<POLY>(r2, 64, 128, sc2, nms2, sum2, pb);
      exp_pack32<POLY>(r3, 96, 128, sc2, nms2, sum2, pb + 16);
    } else {  // fast path: no max, no FFMA; MODE 2 adds the partial-sum checks
      auto over = [&](uint64_t part) {
        float a, c;
        f2_unpack(part, a, c);
        return __any_sync(0xffffffffu, !(a + c <= 32768.0f));
      };
      exp_fast32<POLY>(r0, 0, 128, sum2, pa);
      const uint64_t s1 = sum2;
      exp_fast32<POLY>(r1, 32, 128, sum2, pa + 16);
      const uint64_t s2 = sum2;
      if (MODE == 2 && over(s1)) acc += 1;
      exp_fast32<POLY>(r2, 64, 128, sum2, pb);
      const uint64_t s3 = sum2;
      if (MODE == 2 && over(s2)) acc += 2;
      exp_fast32<POLY>(r3, 96, 128, sum2, pb + 16);
      if (MODE == 2 && (over(s3) || over(sum2))) acc += 3;
    }
    float a, b;
    f2_unpack(sum2, a, b);
    l_run += a + b;
#pragma unroll
    for (int i = 0; i < 32; ++i) acc ^= pa[i] + pb[i];
    mine[r & 127] += 1e-6f * float(acc & 1u);  // the next round's scores depend on this round (one store)
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + __float_as_uint(l_run) + __float_as_uint(m_run);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int POLY, int MODE>
void run(const char* name) {
  float* in;
  uint32_t* out;
  long long* cyc;
  cudaMalloc(&in, 4096 * 4);
  cudaMalloc(&out, 148 * 256 * 4);
  cudaMalloc(&cyc, 148 * 8);
  float h[4096];
  for (int i = 0; i < 4096; ++i) h[i] = float((i * 7919) % 997) / 100.0f - 9.0f;  // [-9, 1): fast-path arguments
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  const int rounds = 256;
  cudaFuncSetAttribute(k<POLY, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 132 * 4);
  for (int threads : {128, 256}) {
    k<POLY, MODE><<<148, threads, 256 * 132 * 4>>>(in, out, cyc, rounds, 0.18f);
    k<POLY, MODE><<<148, threads, 256 * 132 * 4>>>(in, out, cyc, rounds, 0.18f);
    cudaDeviceSynchronize();
    long long c[148];
    cudaMemcpy(c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += c[i];
    avg /= 148;
    printf("%-44s %d warp(s)/SMSP: %7.1f clk per 128-score row (wall) = %7.1f clk per row-warp per SMSP\n", name,
           threads / 128, avg / rounds, avg / rounds / (threads / 128));
  }
}

int main() {
  run<3, 0>("exact: max + FFMA2 + exp, POLY 3/8");
  run<0, 0>("exact: max + FFMA2 + exp, POLY 0/8");
  run<3, 1>("fast: exp only, POLY 3/8");
  run<3, 2>("fast: exp + partial-sum checks, POLY 3/8");
  run<0, 1>("fast: exp only, POLY 0/8");
  run<1, 1>("fast: exp only, POLY 1/8");
  run<2, 1>("fast: exp only, POLY 2/8");
  run<4, 1>("fast: exp only, POLY 4/8");
  return 0;
}
