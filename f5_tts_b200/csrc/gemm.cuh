// tcgen05 GEMM / grouped-conv-as-GEMM for sm_100a.
//
//   C[m, n] = epilogue( sum_k A[m, k] * W[n, k] )        A, W fp16, K-major; fp32 accumulation in TMEM.
//
// One CTA per 128 x BN output tile.  Warp roles (192 threads): warp 0 = TMA producer (one elected lane),
// warp 1 = TMEM allocator + single-thread tcgen05.mma issuer, warps 2..5 = epilogue (TMEM -> registers -> global),
// each epilogue warp owning TMEM lane quarter (warp % 4).  Operand tiles are 64 fp16 wide (128 B rows) in
// SWIZZLE_128B layout written by TMA and read through UMMA shared-memory descriptors; a STAGES-deep
// full/empty mbarrier ring decouples TMA from the tensor core.  Tails in M, N and K are handled by TMA
// out-of-bounds zero fill plus guarded stores.
//
// CONV mode computes the reference's grouped Conv1d(k=31, groups=16, padding=15) (model/modules.py:175-201)
// as 31 accumulated 128x64x64 GEMMs: tap t multiplies the activation tile shifted by (t - 15) rows — the shift is
// just the TMA row coordinate, and rows outside [0, seq) of the SAME sample are zero-filled by the 3-D tensor map,
// which is exactly the conv's zero padding.
#pragma once
#include "common.cuh"
#include "kparams.h"

namespace f5 {

template <int BN, int STAGES>
constexpr size_t gemm_smem_bytes() {
  return size_t(STAGES) * (kBM * kBK * 2 + BN * kBK * 2) + 1024 /*align slack*/ + 256 /*barriers*/;
}

template <int BN, int STAGES, int EPI, int ACT, bool CONV>
__global__ void __launch_bounds__(kGemmThreads)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmParams p) {
  constexpr uint32_t A_BYTES = kBM * kBK * 2;
  constexpr uint32_t B_BYTES = BN * kBK * 2;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
  static_assert(BN == 64 || BN == 128 || BN == 256, "BN");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * (A_BYTES + B_BYTES));
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * kBM;
  const int bz = blockIdx.z;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      // ===== TMA producer =====
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], A_BYTES + B_BYTES);
        if (CONV) {
          // A: activation [batch][seq][channels]; group = blockIdx.x (BN == 64 == channels per group)
          tma_load_3d(sA + s * A_BYTES, &tmA, &full[s], n0, m0 + kb - p.conv_pad, bz);
          // W repacked [tap][out_channel][in 64]: rows = tap * n_out + out_channel
          tma_load_2d(sB + s * B_BYTES, &tmB, &full[s], 0, kb * p.n_out + n0);
        } else {
          tma_load_3d(sA + s * A_BYTES, &tmA, &full[s], kb * kBK, m0, bz);
          tma_load_2d(sB + s * B_BYTES, &tmB, &full[s], kb * kBK, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc = make_idesc_f16(kBM, BN, 0, 0);
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint64_t adesc = make_smem_desc_sw128(smem_u32(sA + s * A_BYTES));
        const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sB + s * B_BYTES));
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) {
          // +32 bytes (16 fp16) along K inside the 128B swizzle atom = +2 in the (addr >> 4) field
          tc_mma_ss(tmem_base, adesc + uint64_t(k * 2), bdesc + uint64_t(k * 2), idesc, (kb | k) != 0);
        }
        tc_commit(&empty[s]);  // smem slot reusable once these MMAs retire
      }
      tc_commit(acc_full);
    }
  } else {
    // ===== epilogue: warps 2..5 -> TMEM lane quarters 2,3,0,1 =====
    const int q = warp & 3;
    const int row_in_batch = m0 + q * 32 + int(lane_id());
    const bool row_ok = row_in_batch < p.rows;
    const long long grow = (long long)bz * p.rows + row_in_batch;
    bool valid = row_ok;
    int pos = 0;
    if (p.seq > 0) {
      pos = int(grow % p.seq);
      if (p.row_len != nullptr && row_ok) valid = pos < p.row_len[grow / p.seq];
    }
    const float* gate = nullptr;
    if (EPI == EPI_RESID && p.gate != nullptr) gate = p.gate + (p.step_ptr ? (long long)(*p.step_ptr) : 0) * p.gate_step_stride;

    mbar_wait(acc_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(c * 32), r);
      tmem_ld_wait();
      const int nc = n0 + c * 32;
      if (!row_ok || nc >= p.n_out) continue;
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float b = (p.bias != nullptr && nc + i < p.n_out) ? __ldg(p.bias + nc + i) : 0.0f;
        v[i] = __uint_as_float(r[i]) + b;
      }
      if (EPI == EPI_QKV_ROPE) {
        const int sec = nc / p.inner;
        const int head = (nc % p.inner) / 64;
        if (sec < 2 && head < p.pe_heads) {
          const int pair0 = (nc % 64) / 2;
          const float* cs = p.rope_cos + (long long)pos * 32 + pair0;
          const float* sn = p.rope_sin + (long long)pos * 32 + pair0;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float c_ = __ldg(cs + i), s_ = __ldg(sn + i);
            const float x0 = v[2 * i], x1 = v[2 * i + 1];
            v[2 * i] = x0 * c_ - x1 * s_;
            v[2 * i + 1] = x1 * c_ + x0 * s_;
          }
        }
      }
      if (ACT != ACT_NONE) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (ACT == ACT_GELU_TANH) v[i] = gelu_tanh(v[i]);
          if (ACT == ACT_GELU_ERF) v[i] = gelu_erf(v[i]);
          if (ACT == ACT_MISH) v[i] = mish(v[i]);
        }
      }
      const bool full_chunk = (nc + 32 <= p.n_out);
      if (EPI == EPI_F16 || EPI == EPI_QKV_ROPE) {
        __half* o = reinterpret_cast<__half*>(p.out) + grow * p.ldo + nc;
        if (full_chunk && (p.ldo % 8 == 0)) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 w;
            w.x = valid ? pack_half2(v[8 * i + 0], v[8 * i + 1]) : 0u;
            w.y = valid ? pack_half2(v[8 * i + 2], v[8 * i + 3]) : 0u;
            w.z = valid ? pack_half2(v[8 * i + 4], v[8 * i + 5]) : 0u;
            w.w = valid ? pack_half2(v[8 * i + 6], v[8 * i + 7]) : 0u;
            reinterpret_cast<uint4*>(o)[i] = w;
          }
        } else {
          for (int i = 0; i < 32 && nc + i < p.n_out; ++i) o[i] = __float2half_rn(valid ? v[i] : 0.0f);
        }
      } else if (EPI == EPI_F32) {
        float* o = reinterpret_cast<float*>(p.out) + grow * p.ldo + nc;
        if (full_chunk && (p.ldo % 4 == 0)) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            reinterpret_cast<float4*>(o)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        } else {
          for (int i = 0; i < 32 && nc + i < p.n_out; ++i) o[i] = v[i];
        }
        if (p.out16b != nullptr) {
          __half* o2 = p.out16b + grow * p.ldo + nc;
          for (int i = 0; i < 32 && nc + i < p.n_out; i += 2)
            *reinterpret_cast<uint32_t*>(o2 + i) = valid ? pack_half2(v[i], v[i + 1]) : 0u;
        }
      } else if (EPI == EPI_RESID) {
        float* o = p.resid + grow * p.ldo + nc;
        if (full_chunk && (p.ldo % 4 == 0)) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 x = reinterpret_cast<float4*>(o)[i];
            float g0 = 1.f, g1 = 1.f, g2 = 1.f, g3 = 1.f;
            if (gate != nullptr) {
              const float4 g = __ldg(reinterpret_cast<const float4*>(gate + nc) + i);
              g0 = g.x; g1 = g.y; g2 = g.z; g3 = g.w;
            }
            if (valid) {
              x.x += g0 * v[4 * i];
              x.y += g1 * v[4 * i + 1];
              x.z += g2 * v[4 * i + 2];
              x.w += g3 * v[4 * i + 3];
              reinterpret_cast<float4*>(o)[i] = x;
            }
          }
        } else {
          for (int i = 0; i < 32 && nc + i < p.n_out; ++i)
            if (valid) o[i] += (gate ? gate[nc + i] : 1.0f) * v[i];
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace f5
