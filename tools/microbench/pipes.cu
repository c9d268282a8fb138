// Issue-rate microbenchmark for the instructions of the attention softmax loop on sm_100a: cycles per warp-instruction
// per SM sub-partition for MUFU.EX2, scalar FFMA / FADD, packed FFMA2 / FADD2 (fp32x2), F2FP (cvt.rn.f16x2.f32),
// FMNMX3 and IMAD, with 1, 2 and 4 warps per sub-partition (blockDim = 128 / 256 / 512 on ONE CTA per SM).
// Each thread runs 16 independent dependency chains of the instruction, 256 instructions per chain round.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/pipes tools/microbench/pipes.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int OP>
__global__ void k(float* out, long long* cyc, float a, float b, int rounds) {
  float x[16];
  uint64_t y[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = a + i + threadIdx.x * 1e-3f;
#pragma unroll
  for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(y[i]) : "f"(x[2 * i]), "f"(x[2 * i + 1]));
  uint64_t bb;
  asm("mov.b64 %0, {%1, %2};" : "=l"(bb) : "f"(b), "f"(b));
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        if (OP == 1) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x[i]) : "f"(b));
        if (OP == 2) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(b));
        if (OP == 3 && i < 8) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(y[i]) : "l"(bb));
        if (OP == 4 && i < 8) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(y[i]) : "l"(bb));
        if (OP == 5 && i < 8) {
          uint32_t h;
          asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x[2 * i]), "f"(x[2 * i + 1]));
          x[2 * i] = __uint_as_float(h);
        }
        if (OP == 6) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(b), "f"(a));
        if (OP == 7) {
          int v = __float_as_int(x[i]);
          asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(v) : "r"(3), "r"(7));
          x[i] = __int_as_float(v);
        }
        if (OP == 8 && i < 8) asm volatile("add.rm.f32x2 %0, %0, %1;" : "+l"(y[i]) : "l"(bb));
      }
    }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(y[i]));
    s += lo + hi;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int per_round) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4);
  cudaMalloc(&cyc, 148 * 8);
  const int rounds = 64;
  for (int threads : {128, 256, 512}) {
    k<OP><<<148, threads>>>(out, cyc, 0.5f, 1.0001f, rounds);
    k<OP><<<148, threads>>>(out, cyc, 0.5f, 1.0001f, rounds);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    const double inst_per_warp = double(rounds) * 16 * per_round;
    const int warps_per_smsp = threads / 128;
    printf("%-22s %d warp(s)/SMSP: %7.2f clk per warp-instruction per SMSP (%.0f clk total)\n", name, warps_per_smsp,
           avg / (inst_per_warp * warps_per_smsp), avg);
  }
}

int main() {
  run<0>("MUFU.EX2", 16);
  run<1>("FFMA (scalar)", 16);
  run<2>("FADD (scalar)", 16);
  run<3>("FFMA2 (f32x2)", 8);
  run<4>("FADD2 (f32x2)", 8);
  run<8>("FADD2.RM (f32x2)", 8);
  run<5>("F2FP f16x2 pack", 8);
  run<6>("FMNMX3", 16);
  run<7>("IMAD", 16);
  return 0;
}
