// Non-causal flash-attention forward on tcgen05 for dim_head = 64 (reference: model/modules.py:471-556,
// F.scaled_dot_product_attention(q, k, v, attn_mask=None | key mask), scale 1/sqrt(64)).
//
// One CTA = one (sample, head, 256-query block) = TWO 128-row query tiles that ping-pong on the tensor core:
// while softmax warpgroup 0 works on S0 = Q0 K_j^T, the tensor core already computes S1 = Q1 K_j^T and the
// pending P V products, so MMA and the exp-heavy softmax overlap inside a CTA.
// q/k/v are read in place from the fused QKV projection output [Be*seq, 3*inner] (fp16) through ONE 3-D TMA tensor
// map (cols, seq rows, sample): no head-split transpose ever touches HBM, and rows past the end of a sample are
// zero-filled by TMA instead of leaking the next sample.
//   S_w = Q_w K^T  : tcgen05.mma 128x128x16 x4 (both operands K-major), fp32 in TMEM
//   softmax        : 2 x 128 threads, one query row each (tcgen05.ld 32x32b: a thread owns a full row -> no shuffles);
//                    the whole 128-key row lives in registers (single pass); online max / sum in fp32, exp2 with the
//                    scale folded in; lazy rescaling: O is only rescaled when some row's max grew by > 2^8
//   O_w += P_w V   : P written fp16 into shared memory in the 128B-swizzled K-major layout, V tile consumed MN-major
//                    straight from its TMA layout; O accumulates in TMEM.
// Warp roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer / TMEM allocator, warps 2..5 softmax WG0,
// warps 6..9 softmax WG1 (warp w touches TMEM lane quarter w % 4).
#pragma once
#include "common.cuh"
#include "kparams.h"

namespace f5 {

__device__ __forceinline__ float ex2_approx(float x) {  // one MUFU.EX2; inputs are <= ~8, -inf -> 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// 32 scores -> 16 packed fp16 pairs of exp2(s * sc - ms); FULL = every key of the chunk is valid (no masking code)
template <bool FULL>
__device__ __forceinline__ void exp_pack32(const uint32_t (&r)[32], int col0, int kv_rem, float sc, float ms,
                                           float& sum_a, float& sum_b, uint32_t* pk) {
  if (!FULL && col0 >= kv_rem) {  // warp-uniform: every key of this chunk is past the end of the sample -> P = 0, no MUFU
#pragma unroll
    for (int i = 0; i < 16; ++i) pk[i] = 0u;
    return;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float e0 = ex2_approx(__uint_as_float(r[2 * i]) * sc - ms);
    float e1 = ex2_approx(__uint_as_float(r[2 * i + 1]) * sc - ms);
    if (!FULL) {
      if (col0 + 2 * i >= kv_rem) e0 = 0.f;
      if (col0 + 2 * i + 1 >= kv_rem) e1 = 0.f;
    }
    sum_a += e0;  // two independent accumulation chains
    sum_b += e1;
    pk[i] = pack_half2(e0, e1);
  }
}

template <bool FULL>
__device__ __forceinline__ float row_max128(const uint32_t (&r0)[32], const uint32_t (&r1)[32], const uint32_t (&r2)[32],
                                            const uint32_t (&r3)[32], int kv_rem) {
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    if (FULL) {
      mx = fmaxf(mx, fmaxf(fmaxf(__uint_as_float(r0[i]), __uint_as_float(r1[i])),
                           fmaxf(__uint_as_float(r2[i]), __uint_as_float(r3[i]))));
    } else {
      if (i < kv_rem) mx = fmaxf(mx, __uint_as_float(r0[i]));
      if (32 + i < kv_rem) mx = fmaxf(mx, __uint_as_float(r1[i]));
      if (64 + i < kv_rem) mx = fmaxf(mx, __uint_as_float(r2[i]));
      if (96 + i < kv_rem) mx = fmaxf(mx, __uint_as_float(r3[i]));
    }
  }
  return mx;
}

__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fwd_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* sQ = smem;                              // 2 tiles
  uint8_t* sK = sQ + 2 * kAttnTile;                // kAttnStages tiles
  uint8_t* sV = sK + kAttnStages * kAttnTile;      // kAttnStages tiles
  uint8_t* sP = sV + kAttnStages * kAttnTile;      // 2 warpgroups x 2 sub-tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * kAttnTile);
  uint64_t* q_full = bars;                         // [1]
  uint64_t* k_full = q_full + 1;                   // [stages]
  uint64_t* k_empty = k_full + kAttnStages;        // [stages]
  uint64_t* v_full = k_empty + kAttnStages;        // [stages]
  uint64_t* v_empty = v_full + kAttnStages;        // [stages]
  uint64_t* s_full = v_empty + kAttnStages;        // [2]
  uint64_t* p_full = s_full + 2;                   // [2]
  uint64_t* o_full = p_full + 2;                   // [2]
  uint64_t* s_free = o_full + 2;                   // [2] softmax has pulled S into registers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + 2);

  const int warp = threadIdx.x >> 5;
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qb * 2 * kAttnBQ;
  const int col_q = h * 64, col_k = p.inner + h * 64, col_v = 2 * p.inner + h * 64;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(q_full, 1);
    for (int s = 0; s < kAttnStages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int w = 0; w < 2; ++w) {
      mbar_init(&s_full[w], 1);
      mbar_init(&p_full[w], 128);
      mbar_init(&o_full[w], 1);
      mbar_init(&s_free[w], 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();
  const int kv_len = p.kv_len ? min(p.kv_len[b], p.seq) : p.seq;
  const int n_kv = (kv_len + kAttnBKV - 1) / kAttnBKV;
  // columns: S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384)

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(q_full, 2 * kAttnTile);
      tma_load_3d(sQ, &tmQKV, q_full, col_q, q0, b);
      tma_load_3d(sQ + kAttnTile, &tmQKV, q_full, col_q, q0 + kAttnBQ, b);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j % kAttnStages;
        const uint32_t ph = (j / kAttnStages) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_expect_tx(&k_full[s], kAttnTile);
        tma_load_3d(sK + s * kAttnTile, &tmQKV, &k_full[s], col_k, j * kAttnBKV, b);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_expect_tx(&v_full[s], kAttnTile);
        tma_load_3d(sV + s * kAttnTile, &tmQKV, &v_full[s], col_v, j * kAttnBKV, b);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_f16(128, 128, 0, 0);  // S = Q K^T : both K-major
      constexpr uint32_t idesc_o = make_idesc_f16(128, 64, 0, 1);   // O = P V   : V is MN-major
      auto issue_s = [&](int w, int ks) {
        const uint64_t qdesc = make_smem_desc_sw128(smem_u32(sQ + w * kAttnTile));
        const uint64_t kdesc = make_smem_desc_sw128(smem_u32(sK + ks * kAttnTile));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_ss(tmem_base + w * 128, qdesc + uint64_t(2 * k), kdesc + uint64_t(2 * k), idesc_s, k != 0);
        tc_commit(&s_full[w]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_s(0, 0);
      issue_s(1, 0);
      tc_commit(&k_empty[0]);
      for (int j = 0; j < n_kv; ++j) {
        // S of the NEXT tile is issued as soon as the softmax warps have pulled the current S into registers
        // (s_free), i.e. long before their exponentials / P are done: the Q K^T latency leaves the softmax chain.
        if (j + 1 < n_kv) {
          const int sk = (j + 1) % kAttnStages;
          mbar_wait(&k_full[sk], ((j + 1) / kAttnStages) & 1);
          for (int w = 0; w < 2; ++w) {
            mbar_wait(&s_free[w], j & 1);
            tc_fence_after();
            issue_s(w, sk);
          }
          tc_commit(&k_empty[sk]);
        }
        const int sv = j % kAttnStages;
        mbar_wait(&v_full[sv], (j / kAttnStages) & 1);
        for (int w = 0; w < 2; ++w) {
          mbar_wait(&p_full[w], j & 1);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint64_t pdesc =
                make_smem_desc_sw128(smem_u32(sP + (2 * w + (kk >> 2)) * kAttnTile)) + uint64_t(2 * (kk & 3));
            const uint64_t vdesc = make_smem_desc_sw128(smem_u32(sV + sv * kAttnTile + kk * 16 * 128));
            tc_mma_ss(tmem_base + 256 + w * 64, pdesc, vdesc, idesc_o, (j | kk) != 0);
          }
          tc_commit(&o_full[w]);
        }
        tc_commit(&v_empty[sv]);
      }
    }
  } else {
    const int w = (warp - 2) >> 2;  // softmax warpgroup 0 / 1
    const int q = warp & 3;         // TMEM lane quarter
    const int row = q * 32 + int(lane_id());
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    const uint32_t tmem_S = tmem_base + w * 128 + lane_off;
    const uint32_t tmem_O = tmem_base + 256 + w * 64 + lane_off;
    uint8_t* sPw = sP + 2 * w * kAttnTile;
    float m_run = -INFINITY, l_run = 0.0f;
#ifdef F5_TRACE
    long long* ts = p.dbg_ts ? p.dbg_ts + ((long long)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16 + w * 8 : nullptr;
#endif
#ifdef F5_TRACE
    long long c_s = 0, c_turn = 0, c_exp = 0, c_o = 0, c_p = 0, c0 = 0, c_begin = 0, c_ld = 0, c_max = 0;
#endif
#ifdef F5_TRACE
    if (ts) c_begin = clock64();
#endif
    // Ping-pong turnstile (named barriers 3 + w, 256 threads): the two warpgroups take turns in the exp2-heavy
    // section, so one warpgroup's MUFU work overlaps the other's tensor-core work instead of both running in lockstep.
    if (p.turnstile && w == 1) named_bar_arrive(3, 256);  // WG0 goes first
    for (int j = 0; j < n_kv; ++j) {
      const int kv_rem = kv_len - j * kAttnBKV;  // valid keys in this tile (>= 1)
#ifdef F5_TRACE
      if (ts) c0 = clock64();
#endif
      mbar_wait(&s_full[w], j & 1);
      tc_fence_after();
#ifdef F5_TRACE
      if (ts) { const long long c1 = clock64(); c_s += c1 - c0; c0 = c1; }
#endif
      uint32_t r0[32], r1[32], r2[32], r3[32];
      tmem_ld32(tmem_S + 0, r0);
      tmem_ld32(tmem_S + 32, r1);
      tmem_ld32(tmem_S + 64, r2);
      tmem_ld32(tmem_S + 96, r3);
      tmem_ld_wait();
#ifdef F5_TRACE
      if (ts) { const long long c1 = clock64(); c_ld += c1 - c0; c0 = c1; }
#endif
      tc_fence_before();
      mbar_arrive(&s_free[w]);  // S_w is in registers: the tensor core may overwrite it with the next tile's scores
      const bool full_tile = kv_rem >= kAttnBKV;  // warp-uniform
      const float mx = full_tile ? row_max128<true>(r0, r1, r2, r3, kv_rem) : row_max128<false>(r0, r1, r2, r3, kv_rem);
      const float m_new = fmaxf(m_run, mx * p.scale_log2);
      // lazy rescale (warp-uniform decision because tcgen05.ld/st are warp-collective)
      const bool grow = (m_new - m_run) > 8.0f;  // also true on the first tile (m_run = -inf)
      const bool do_rescale = __any_sync(0xffffffffu, grow);
      float alpha = 1.0f;
      if (do_rescale) {
        alpha = ex2_approx(m_run - m_new);  // first tile: exp2(-inf) = 0
        m_run = m_new;
      }
#ifdef F5_TRACE
      if (ts) { const long long c1 = clock64(); c_max += c1 - c0; c0 = c1; }
#endif
      if (p.turnstile) named_bar_sync(3 + w, 256);  // only the MUFU-bound exp2 loop is serialised between the warpgroups
#ifdef F5_TRACE
      if (ts) { const long long c1 = clock64(); c_turn += c1 - c0; c0 = c1; }
#endif
      // exponentials -> packed fp16 (kept in registers until the P buffer is free)
      const float ms = m_run;
      const float sc = p.scale_log2;
      float lsum = 0.0f, lsum2 = 0.0f;
      uint32_t pk[64];
      if (full_tile) {
        exp_pack32<true>(r0, 0, kv_rem, sc, ms, lsum, lsum2, pk);
        exp_pack32<true>(r1, 32, kv_rem, sc, ms, lsum, lsum2, pk + 16);
        exp_pack32<true>(r2, 64, kv_rem, sc, ms, lsum, lsum2, pk + 32);
        exp_pack32<true>(r3, 96, kv_rem, sc, ms, lsum, lsum2, pk + 48);
      } else {
        exp_pack32<false>(r0, 0, kv_rem, sc, ms, lsum, lsum2, pk);
        exp_pack32<false>(r1, 32, kv_rem, sc, ms, lsum, lsum2, pk + 16);
        exp_pack32<false>(r2, 64, kv_rem, sc, ms, lsum, lsum2, pk + 32);
        exp_pack32<false>(r3, 96, kv_rem, sc, ms, lsum, lsum2, pk + 48);
      }
      lsum += lsum2;
      l_run = l_run * alpha + lsum;
      if (p.turnstile) named_bar_arrive(3 + (w ^ 1), 256);  // hand the MUFU-heavy section to the other warpgroup
#ifdef F5_TRACE
      if (ts) { const long long c1 = clock64(); c_exp += c1 - c0; c0 = c1; }
#endif
      if (j > 0) {
        mbar_wait(&o_full[w], (j - 1) & 1);  // P V of the previous tile retired: P buffer and O are ours
        tc_fence_after();
      }
#ifdef F5_TRACE
      if (ts) { const long long c1 = clock64(); c_o += c1 - c0; c0 = c1; }
#endif
      // P -> shared memory, 128B-swizzled K-major: key k lives in sub-tile k/64, 16-byte chunk (k%64)/8
      uint8_t* prow = sPw + row * 128;
#pragma unroll
      for (int g = 0; g < 16; ++g) {
        const int sub = g >> 3, chunk16 = g & 7;
        const uint4 wv = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
        *reinterpret_cast<uint4*>(prow + sub * kAttnTile + ((chunk16 ^ (row & 7)) << 4)) = wv;
      }
      if (do_rescale && j > 0) {
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tmem_ld32(tmem_O + uint32_t(c * 32), r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
          tmem_st32(tmem_O + uint32_t(c * 32), r);
        }
        tmem_st_wait();
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&p_full[w]);
#ifdef F5_TRACE
      if (ts) c_p += clock64() - c0;
#endif
    }
#ifdef F5_TRACE
    if (ts && row == 0) {
      ts[0] = c_s; ts[1] = c_turn; ts[2] = c_exp; ts[3] = c_o; ts[4] = c_p; ts[5] = clock64() - c_begin; ts[6] = c_max; ts[7] = c_ld;
    }
#endif
    // epilogue: O / l -> fp16
    mbar_wait(&o_full[w], (n_kv - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
    const int qrow = q0 + w * kAttnBQ + row;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      tmem_ld32(tmem_O + uint32_t(c * 32), r);
      tmem_ld_wait();
      if (qrow < p.seq) {
        __half* o = p.out + ((long long)b * p.seq + qrow) * p.inner + h * 64 + c * 32;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 wv;
          wv.x = pack_half2(__uint_as_float(r[8 * g + 0]) * inv_l, __uint_as_float(r[8 * g + 1]) * inv_l);
          wv.y = pack_half2(__uint_as_float(r[8 * g + 2]) * inv_l, __uint_as_float(r[8 * g + 3]) * inv_l);
          wv.z = pack_half2(__uint_as_float(r[8 * g + 4]) * inv_l, __uint_as_float(r[8 * g + 5]) * inv_l);
          wv.w = pack_half2(__uint_as_float(r[8 * g + 6]) * inv_l, __uint_as_float(r[8 * g + 7]) * inv_l);
          reinterpret_cast<uint4*>(o)[g] = wv;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}


}  // namespace f5
