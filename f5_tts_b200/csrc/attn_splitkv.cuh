// EXPERIMENTAL (selected only by F5_ATTN_VARIANT=6, never by default).  Written at the end of round 1: it passes the
// attention parity tests on B200 (`F5_ATTN_VARIANT=6 pytest tests/test_gpu_kernels.py -m gpu -k attention`: 7 / 7), but the
// round's GPU budget ran out before it could be TIMED — run tools/attn_bench.py with and without F5_ATTN_VARIANT=6 and the
// end-to-end tests before making it the default.
//
// Split-KV variant of the dim_head-64 flash-attention forward (same contract as attn_fwd_tcgen05_kernel, attn.cuh).
// Why: at cfg2 an SM owns only ~1.6 query tiles, i.e. two softmax warps per scheduler.  Measurements (profiles/README.md,
// tools/microbench/mufu.cu): a lone warp pays 13.9 clk per exponentiated element (its own MUFU issue blocks it), two
// warps that are both inside their exp loop reach 9.1 clk / element / scheduler, the MUFU floor is 8 — but with two
// warps per scheduler one of them is always in the ~1500 clk of per-tile latency (wait S, TMEM load, max, P store,
// proxy fence, wait P V), so the MUFU idles ~45 % of the time.  Here every 128-row query tile is worked on by TWO
// warpgroups that split the KEY range in halves (64-key tiles), each with its own running (m, l, O); that puts four
// independent softmax warps on every scheduler without touching the grid.  The two partial results of a query tile are
// merged at the end:  O = (2^(m0-m) O0 + 2^(m1-m) O1) / (2^(m0-m) l0 + 2^(m1-m) l1),  m = max(m0, m1).
//
// One CTA = (sample, head, 256 queries); 576 threads: warp 0 TMA producer, warp 1 MMA issuer, warps 2..17 softmax.
// Stream st = 2 * t + h (t = query tile 0/1, h = key half 0/1) owns warps 2 + 4 st .. 5 + 4 st; warp w touches TMEM lane
// quarter w % 4.  TMEM columns: S[st] at 64 st, O[st] at 256 + 64 st (512 in total).
// Shared memory: Q 2 x 16 KB | per half a ring of kSkvStages x {K 8 KB, V 8 KB} | P[st] 16 KB (128 rows x 64 keys fp16,
// one 128B-swizzle atom wide) | barriers | row statistics.  The P buffer of a stream doubles as its exchange buffer in
// the merge (32 of its 64 output columns, fp32).
#pragma once
#include "attn.cuh"

namespace f5 {

constexpr int kSkvThreads = 64 + 16 * 32;
constexpr int kSkvBK = 64;                        // keys per tile
constexpr int kSkvStages = 3;                     // K / V ring depth per key half
constexpr uint32_t kSkvKvTile = kSkvBK * 64 * 2;  // 8 KB
constexpr size_t kSkvSmem = size_t(kAttnTile) * 2 /*Q*/ + size_t(kSkvKvTile) * 2 * 2 * kSkvStages /*K,V rings*/ +
                            size_t(kAttnTile) * 4 /*P*/ + 1024 /*align*/ + 512 /*barriers*/ + 4 * 128 * 2 * 4 /*row stats*/;

__global__ void __launch_bounds__(kSkvThreads, 1)
attn_fwd_splitkv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* sQ = smem;                                                // 2 x 16 KB
  uint8_t* sK = sQ + 2 * kAttnTile;                                  // [half][stage] 8 KB
  uint8_t* sV = sK + 2 * kSkvStages * kSkvKvTile;                    // [half][stage] 8 KB
  uint8_t* sP = sV + 2 * kSkvStages * kSkvKvTile;                    // [stream] 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * kAttnTile);
  uint64_t* q_full = bars;                                           // [1]
  uint64_t* k_full = q_full + 1;                                     // [2][stages]
  uint64_t* k_empty = k_full + 2 * kSkvStages;
  uint64_t* v_full = k_empty + 2 * kSkvStages;
  uint64_t* v_empty = v_full + 2 * kSkvStages;
  uint64_t* s_full = v_empty + 2 * kSkvStages;                       // [4]
  uint64_t* s_free = s_full + 4;
  uint64_t* p_full = s_free + 4;
  uint64_t* o_full = p_full + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 4);
  float* sStat = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);  // [stream][128][2] = (m, l)

  const int warp = threadIdx.x >> 5;
  const int qb = blockIdx.x, h_idx = blockIdx.y, b = blockIdx.z;
  const int q0 = qb * 2 * kAttnBQ;
  const int col_q = h_idx * 64, col_k = p.inner + h_idx * 64, col_v = 2 * p.inner + h_idx * 64;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2 * kSkvStages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int st = 0; st < 4; ++st) {
      mbar_init(&s_full[st], 1);
      mbar_init(&s_free[st], 128);
      mbar_init(&p_full[st], 128);
      mbar_init(&o_full[st], 1);
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
  const int n64 = (kv_len + kSkvBK - 1) / kSkvBK;  // >= 1
  const int n_half[2] = {(n64 + 1) / 2, n64 / 2};  // tiles of key half 0 / 1 (half 1 may be empty)
  const int tile0[2] = {0, (n64 + 1) / 2};         // first 64-key tile of each half

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(q_full, 2 * kAttnTile);
      tma_load_3d(sQ, &tmQ, q_full, col_q, q0, b);
      tma_load_3d(sQ + kAttnTile, &tmQ, q_full, col_q, q0 + kAttnBQ, b);
      for (int i = 0; i < n_half[0]; ++i) {  // n_half[0] >= n_half[1]
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          if (i >= n_half[hh]) continue;
          const int s = hh * kSkvStages + i % kSkvStages;
          const uint32_t ph = (i / kSkvStages) & 1;
          const int key0 = (tile0[hh] + i) * kSkvBK;
          mbar_wait(&k_empty[s], ph ^ 1);
          mbar_expect_tx(&k_full[s], kSkvKvTile);
          tma_load_3d(sK + s * kSkvKvTile, &tmKV, &k_full[s], col_k, key0, b);
          mbar_wait(&v_empty[s], ph ^ 1);
          mbar_expect_tx(&v_full[s], kSkvKvTile);
          tma_load_3d(sV + s * kSkvKvTile, &tmKV, &v_full[s], col_v, key0, b);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_f16(128, 64, 0, 0);  // S = Q K^T : both K-major
      constexpr uint32_t idesc_o = make_idesc_f16(128, 64, 0, 1);  // O = P V   : V is MN-major
      auto issue_s = [&](int st, int ks) {  // ks: ring slot index (half * stages + stage)
        const uint64_t qdesc = make_smem_desc_sw128(smem_u32(sQ + (st >> 1) * kAttnTile));
        const uint64_t kdesc = make_smem_desc_sw128(smem_u32(sK + ks * kSkvKvTile));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_ss(tmem_base + st * 64, qdesc + uint64_t(2 * k), kdesc + uint64_t(2 * k), idesc_s, k != 0);
        tc_commit(&s_full[st]);
      };
      mbar_wait(q_full, 0);
      // Event-driven issue over the four independent streams: S_st(j) needs K_j of its half and "S_st(j-1) is in
      // registers"; P_st(j) V_j needs V_j and P_st(j).  K / V slots go back to the producer once both query tiles of the
      // half have consumed them.
      int js[4] = {0, 0, 0, 0}, jp[4] = {0, 0, 0, 0};
      int k_rel[2] = {0, 0}, v_rel[2] = {0, 0};
      long long spin_t0 = clock64();
      auto all_done = [&]() {
        return jp[0] >= n_half[0] && jp[2] >= n_half[0] && jp[1] >= n_half[1] && jp[3] >= n_half[1];
      };
      while (!all_done()) {
        bool progress = false;
#pragma unroll
        for (int st = 0; st < 4; ++st) {
          const int hh = st & 1, nh = n_half[hh];
          if (js[st] < nh) {
            const int j = js[st], ks = hh * kSkvStages + j % kSkvStages;
            if (mbar_test(&k_full[ks], (j / kSkvStages) & 1) && (j == 0 || mbar_test(&s_free[st], (j - 1) & 1))) {
              tc_fence_after();
              issue_s(st, ks);
              ++js[st];
              progress = true;
              if (min(js[hh], js[hh + 2]) > k_rel[hh]) {  // both query tiles have consumed K tile k_rel of this half
                tc_commit(&k_empty[hh * kSkvStages + k_rel[hh] % kSkvStages]);
                ++k_rel[hh];
              }
            }
          }
          if (jp[st] < nh) {
            const int j = jp[st], vs = hh * kSkvStages + j % kSkvStages;
            if (mbar_test(&v_full[vs], (j / kSkvStages) & 1) && mbar_test(&p_full[st], j & 1)) {
              tc_fence_after();
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                const uint64_t pdesc = make_smem_desc_sw128(smem_u32(sP + st * kAttnTile)) + uint64_t(2 * kk);
                const uint64_t vdesc = make_smem_desc_sw128(smem_u32(sV + vs * kSkvKvTile + kk * 16 * 128));
                tc_mma_ss(tmem_base + 256 + st * 64, pdesc, vdesc, idesc_o, (j | kk) != 0);
              }
              tc_commit(&o_full[st]);
              ++jp[st];
              progress = true;
              if (min(jp[hh], jp[hh + 2]) > v_rel[hh]) {
                tc_commit(&v_empty[hh * kSkvStages + v_rel[hh] % kSkvStages]);
                ++v_rel[hh];
              }
            }
          }
        }
        if (progress) {
          spin_t0 = clock64();
        } else if (clock64() - spin_t0 > F5_SPIN_TIMEOUT_CYCLES) {
          printf("f5: split-KV attention MMA issuer stalled block(%d,%d,%d) js %d %d %d %d jp %d %d %d %d\n", blockIdx.x,
                 blockIdx.y, blockIdx.z, js[0], js[1], js[2], js[3], jp[0], jp[1], jp[2], jp[3]);
          __trap();
        }
      }
    }
  } else {
    const int st = (warp - 2) >> 2;  // stream 0..3
    const int t = st >> 1, hh = st & 1;
    const int q = warp & 3;          // TMEM lane quarter
    const int row = q * 32 + int(lane_id());
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    const uint32_t tmem_S = tmem_base + st * 64 + lane_off;
    const uint32_t tmem_O = tmem_base + 256 + st * 64 + lane_off;
    uint8_t* sPs = sP + st * kAttnTile;
    uint8_t* prow = sPs + row * 128;
    const int nh = n_half[hh];
    float m_run = -INFINITY, l_run = 0.0f;
    for (int j = 0; j < nh; ++j) {
      const int kv_rem = kv_len - (tile0[hh] + j) * kSkvBK;  // valid keys in this tile (>= 1)
      mbar_wait(&s_full[st], j & 1);
      tc_fence_after();
      uint32_t r0[32], r1[32];
      tmem_ld32(tmem_S + 0, r0);
      tmem_ld32(tmem_S + 32, r1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_free[st]);  // S is in registers: the tensor core may overwrite it with the next tile's scores
      const bool full_tile = kv_rem >= kSkvBK;  // warp-uniform
      float mx = -INFINITY;
      if (full_tile) {
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(r0[i]), __uint_as_float(r1[i])));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i < kv_rem) mx = fmaxf(mx, __uint_as_float(r0[i]));
          if (32 + i < kv_rem) mx = fmaxf(mx, __uint_as_float(r1[i]));
        }
      }
      const float m_new = fmaxf(m_run, mx * p.scale_log2);
      const bool grow = (m_new - m_run) > 8.0f;  // also true on the first tile (m_run = -inf)
      const bool do_rescale = __any_sync(0xffffffffu, grow);  // warp-uniform: tcgen05.ld/st are warp-collective
      float alpha = 1.0f;
      if (do_rescale) {
        alpha = ex2_approx(m_run - m_new);  // first tile: exp2(-inf) = 0
        m_run = m_new;
      }
      float lsum = 0.0f, lsum2 = 0.0f;
      uint32_t pk[32];
      if (full_tile) {
        exp_pack32<true>(r0, 0, kv_rem, p.scale_log2, m_run, lsum, lsum2, pk);
        exp_pack32<true>(r1, 32, kv_rem, p.scale_log2, m_run, lsum, lsum2, pk + 16);
      } else {
        exp_pack32<false>(r0, 0, kv_rem, p.scale_log2, m_run, lsum, lsum2, pk);
        exp_pack32<false>(r1, 32, kv_rem, p.scale_log2, m_run, lsum, lsum2, pk + 16);
      }
      l_run = l_run * alpha + (lsum + lsum2);
      if (j > 0) {
        mbar_wait(&o_full[st], (j - 1) & 1);  // P V of the previous tile retired: P buffer and O are ours
        tc_fence_after();
      }
      // P -> shared memory, 128B-swizzled K-major: key k lives in 16-byte chunk k / 8 of the row
#pragma unroll
      for (int g = 0; g < 8; ++g)
        *reinterpret_cast<uint4*>(prow + ((g ^ (row & 7)) << 4)) = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
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
      mbar_arrive(&p_full[st]);
    }
    // ---- merge the two key halves of query tile t: each stream finalises 32 of the 64 output columns and hands the
    //      other 32 (fp32, [c4][row] float4 layout: conflict free) to its partner through its own, now idle, P buffer
    if (nh > 0) {
      mbar_wait(&o_full[st], (nh - 1) & 1);
      tc_fence_after();
    }
    float* stat = sStat + (st * 128 + row) * 2;
    stat[0] = m_run;  // -inf when this half had no tile
    stat[1] = l_run;
    const int give = (hh ^ 1) * 32;  // columns the partner finalises
    const int keep = hh * 32;        // columns this stream finalises
    {
      uint32_t r[32];
      if (nh > 0) {
        tmem_ld32(tmem_O + uint32_t(give), r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = 0u;  // O was never written
      }
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4)
        reinterpret_cast<uint4*>(sPs)[c4 * 128 + row] = make_uint4(r[4 * c4], r[4 * c4 + 1], r[4 * c4 + 2], r[4 * c4 + 3]);
    }
    named_bar_sync(1 + t, 256);  // both streams of query tile t
    const int pst = st ^ 1;      // partner stream
    const float m_o = sStat[(pst * 128 + row) * 2], l_o = sStat[(pst * 128 + row) * 2 + 1];
    const float m = fmaxf(m_run, m_o);  // finite: half 0 always has a tile with at least one valid key
    const float a_me = ex2_approx(m_run - m), a_o = ex2_approx(m_o - m);  // exp2(-inf) = 0 for an empty half
    const float inv_l = 1.0f / (l_run * a_me + l_o * a_o);
    uint32_t mine[32];
    if (nh > 0) {
      tmem_ld32(tmem_O + uint32_t(keep), mine);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) mine[i] = 0u;
    }
    const int qrow = q0 + t * kAttnBQ + row;
    if (qrow < p.seq) {
      const uint4* theirs = reinterpret_cast<const uint4*>(sP + pst * kAttnTile);
      __half* o = p.out + ((long long)b * p.seq + qrow) * p.inner + h_idx * 64 + keep;
      const float wa = a_me * inv_l, wb = a_o * inv_l;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint4 u0 = theirs[(2 * g) * 128 + row], u1 = theirs[(2 * g + 1) * 128 + row];
        uint4 wv;
        wv.x = pack_half2(__uint_as_float(mine[8 * g + 0]) * wa + __uint_as_float(u0.x) * wb,
                          __uint_as_float(mine[8 * g + 1]) * wa + __uint_as_float(u0.y) * wb);
        wv.y = pack_half2(__uint_as_float(mine[8 * g + 2]) * wa + __uint_as_float(u0.z) * wb,
                          __uint_as_float(mine[8 * g + 3]) * wa + __uint_as_float(u0.w) * wb);
        wv.z = pack_half2(__uint_as_float(mine[8 * g + 4]) * wa + __uint_as_float(u1.x) * wb,
                          __uint_as_float(mine[8 * g + 5]) * wa + __uint_as_float(u1.y) * wb);
        wv.w = pack_half2(__uint_as_float(mine[8 * g + 6]) * wa + __uint_as_float(u1.z) * wb,
                          __uint_as_float(mine[8 * g + 7]) * wa + __uint_as_float(u1.w) * wb);
        reinterpret_cast<uint4*>(o)[g] = wv;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace f5
