// Non-causal flash-attention forward on tcgen05 for dim_head = 64 (reference: model/modules.py:471-556,
// F.scaled_dot_product_attention(q, k, v, attn_mask=None | key mask), scale 1/sqrt(64)).
//
// One CTA = one (sample, head, 128-query tile).  q/k/v are read in place from the fused QKV projection output
// [Be*seq, 3*inner] (fp16) through ONE 3-D TMA tensor map (cols, seq rows, sample): no head-split transpose ever
// touches HBM, and rows past the end of a sample are zero-filled by TMA instead of leaking the next sample.
//   S = Q K^T     : tcgen05.mma 128x128x16 x4, A = Q tile (K-major), B = K tile (K-major), fp32 in TMEM cols [0,128)
//   softmax       : 128 threads, one query row each (tcgen05.ld 32x32b: a thread owns a full row -> no shuffles),
//                   online max / sum in fp32, exp2 with the scale folded in
//   O += P V      : P written fp16 to shared memory in the 128B-swizzled K-major layout, V tile consumed MN-major
//                   straight from its TMA layout; O accumulates in TMEM cols [128,192) and is rescaled in TMEM.
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer / TMEM allocator, warps 2..5 softmax + epilogue.
// Shared memory is sized so two CTAs co-reside per SM: one CTA's softmax overlaps the other's MMAs.
#pragma once
#include "common.cuh"
#include "kparams.h"

namespace f5 {

__global__ void __launch_bounds__(kAttnThreads)
attn_fwd_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t pad = base - smem_u32(smem_raw);
  uint8_t* smem = smem_raw + pad;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kAttnTile;
  uint8_t* sV = sK + kAttnStages * kAttnTile;
  uint8_t* sP = sV + kAttnStages * kAttnTile;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kAttnTile);  // lives in the alignment slack (pad <= 896)
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;               // [stages]
  uint64_t* kv_empty = kv_full + kAttnStages; // [stages]
  uint64_t* s_full = kv_empty + kAttnStages;
  uint64_t* p_full = s_full + 1;
  uint64_t* o_full = p_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int kv_len = p.kv_len ? min(p.kv_len[b], p.seq) : p.seq;
  const int n_kv = (kv_len + kAttnBKV - 1) / kAttnBKV;
  const int q0 = qt * kAttnBQ;
  const int col_q = h * 64, col_k = p.inner + h * 64, col_v = 2 * p.inner + h * 64;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(q_full, 1);
    for (int s = 0; s < kAttnStages; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + 128;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(q_full, kAttnTile);
      tma_load_3d(sQ, &tmQKV, q_full, col_q, q0, b);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j % kAttnStages;
        const uint32_t ph = (j / kAttnStages) & 1;
        mbar_wait(&kv_empty[s], ph ^ 1);
        mbar_expect_tx(&kv_full[s], 2 * kAttnTile);
        tma_load_3d(sK + s * kAttnTile, &tmQKV, &kv_full[s], col_k, j * kAttnBKV, b);
        tma_load_3d(sV + s * kAttnTile, &tmQKV, &kv_full[s], col_v, j * kAttnBKV, b);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_f16(128, 128, 0, 0);  // S = Q K^T : both K-major
      constexpr uint32_t idesc_o = make_idesc_f16(128, 64, 0, 1);   // O = P V   : V is MN-major
      const uint64_t qdesc = make_smem_desc_sw128(smem_u32(sQ));
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      {
        const uint64_t kdesc = make_smem_desc_sw128(smem_u32(sK));
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_ss(tmem_S, qdesc + uint64_t(2 * k), kdesc + uint64_t(2 * k), idesc_s, k != 0);
        tc_commit(s_full);
      }
      for (int j = 0; j < n_kv; ++j) {
        const int s = j % kAttnStages;
        mbar_wait(p_full, j & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t pdesc = make_smem_desc_sw128(smem_u32(sP + (kk >> 2) * kAttnTile)) + uint64_t(2 * (kk & 3));
          const uint64_t vdesc = make_smem_desc_sw128(smem_u32(sV + s * kAttnTile + kk * 16 * 128));
          tc_mma_ss(tmem_O, pdesc, vdesc, idesc_o, (j | kk) != 0);
        }
        tc_commit(&kv_empty[s]);
        tc_commit(o_full);
        if (j + 1 < n_kv) {
          const int s1 = (j + 1) % kAttnStages;
          mbar_wait(&kv_full[s1], ((j + 1) / kAttnStages) & 1);
          tc_fence_after();
          const uint64_t kdesc = make_smem_desc_sw128(smem_u32(sK + s1 * kAttnTile));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_ss(tmem_S, qdesc + uint64_t(2 * k), kdesc + uint64_t(2 * k), idesc_s, k != 0);
          tc_commit(s_full);
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + int(lane_id());
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    float m_run = -INFINITY, l_run = 0.0f;
    for (int j = 0; j < n_kv; ++j) {
      const int kv_rem = kv_len - j * kAttnBKV;  // valid keys in this tile (>= 1)
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      // pass 1: row max
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_S + lane_off + uint32_t(c * 32), r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c * 32 + i < kv_rem) mx = fmaxf(mx, __uint_as_float(r[i]));
      }
      const float m_new = fmaxf(m_run, mx * p.scale_log2);
      const float alpha = exp2f(m_run - m_new);  // first tile: exp2(-inf) = 0
      if (j > 0) mbar_wait(o_full, (j - 1) & 1);  // PV_{j-1} retired: P buffer and O are ours
      tc_fence_after();
      // pass 2: P = exp2(s*scale - m_new) -> smem (fp16, swizzled), running sum
      float lsum = 0.0f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_S + lane_off + uint32_t(c * 32), r);
        tmem_ld_wait();
        float pv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float e = (c * 32 + i < kv_rem) ? exp2f(__uint_as_float(r[i]) * p.scale_log2 - m_new) : 0.0f;
          pv[i] = e;
          lsum += e;
        }
        uint8_t* prow = sP + (c >> 1) * kAttnTile + row * 128;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int chunk16 = (c & 1) * 4 + g;
          uint4 w;
          w.x = pack_half2(pv[8 * g + 0], pv[8 * g + 1]);
          w.y = pack_half2(pv[8 * g + 2], pv[8 * g + 3]);
          w.z = pack_half2(pv[8 * g + 4], pv[8 * g + 5]);
          w.w = pack_half2(pv[8 * g + 6], pv[8 * g + 7]);
          *reinterpret_cast<uint4*>(prow + ((chunk16 ^ (row & 7)) << 4)) = w;
        }
      }
      l_run = l_run * alpha + lsum;
      m_run = m_new;
      if (j > 0) {  // rescale the running output in TMEM
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tmem_ld32(tmem_O + lane_off + uint32_t(c * 32), r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
          tmem_st32(tmem_O + lane_off + uint32_t(c * 32), r);
        }
        tmem_st_wait();
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // epilogue: O / l -> fp16
    mbar_wait(o_full, (n_kv - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
    const int qrow = q0 + row;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      tmem_ld32(tmem_O + lane_off + uint32_t(c * 32), r);
      tmem_ld_wait();
      if (qrow < p.seq) {
        __half* o = p.out + ((long long)b * p.seq + qrow) * p.inner + h * 64 + c * 32;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 w;
          w.x = pack_half2(__uint_as_float(r[8 * g + 0]) * inv_l, __uint_as_float(r[8 * g + 1]) * inv_l);
          w.y = pack_half2(__uint_as_float(r[8 * g + 2]) * inv_l, __uint_as_float(r[8 * g + 3]) * inv_l);
          w.z = pack_half2(__uint_as_float(r[8 * g + 4]) * inv_l, __uint_as_float(r[8 * g + 5]) * inv_l);
          w.w = pack_half2(__uint_as_float(r[8 * g + 6]) * inv_l, __uint_as_float(r[8 * g + 7]) * inv_l);
          reinterpret_cast<uint4*>(o)[g] = w;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

}  // namespace f5
