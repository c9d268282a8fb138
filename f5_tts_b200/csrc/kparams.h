// Kernel parameter blocks and tile constants shared by device code and the host planners.
#pragma once
#include <cuda_fp16.h>
#include <stddef.h>
#include <stdint.h>

namespace f5 {

enum : int { ACT_NONE = 0, ACT_GELU_TANH = 1, ACT_GELU_ERF = 2, ACT_MISH = 3 };
enum : int {
  EPI_F16 = 0,       // out16[m,n] = mask(act(acc + bias))
  EPI_F32 = 1,       // out32[m,n] = acc + bias ; optional out16b[m,n] = mask(acc + bias) (fp16 copy)
  EPI_RESID = 2,     // resid32[m,n] += gate[n] * mask(act(acc + bias))
  EPI_QKV_ROPE = 3,  // out16[m,n] = rope(acc + bias) for the first pe_heads heads of q,k; plain for the rest
};

struct GemmParams {
  int rows;      // valid rows per batch entry
  int n_out;     // output columns
  int num_kb;    // 64-wide K blocks (plain) or taps (conv)
  int batches;   // gridDim.z
  const float* bias;
  void* out;     // fp16 (EPI_F16/QKV) or fp32 (EPI_F32)
  __half* out16b;  // optional fp16 copy for EPI_F32
  float* resid;  // EPI_RESID in/out
  int ldo;       // leading dimension of out / resid (elements)
  const float* gate;  // [n_out] at gate + (*step_ptr) * gate_step_stride, or null (gate = 1)
  const int* step_ptr;
  long long gate_step_stride;
  const int* row_len;  // per-sample valid length (sample = global_row / seq), or null
  int seq;             // rows per sample
  const float* rope_cos;  // [seq, dh/2]
  const float* rope_sin;
  int inner;     // heads * dim_head
  int pe_heads;  // heads that get rotary (q and k sections)
  int conv_pad;  // CONV: taps/2
  int skip_pad;    // 1: 128-row (256 for CTA pairs) tiles whose rows all lie past their sample's row_len are not computed
  int w_prefetch;  // W tiles may be loaded before griddepcontrol.wait (weights are not produced by the predecessor)
  long long* dbg_ts;  // optional [gridDim.x][8] clock64/globaltimer trace (diagnostics; NULL in production)
  int dbg_mode;  // 0 normal; 1 = skip TMA loads, 2 = skip MMAs, 3 = skip epilogue math/stores (perf decomposition only)
};

constexpr int kBM = 128;
constexpr int kBK = 64;


struct AttnParams {
  int seq;            // tokens per sample
  int heads;
  int batches;        // packed batch Be
  int inner;          // heads * 64
  const int* kv_len;  // [Be] valid keys per sample, or null (= seq)
  float scale_log2;   // softmax scale * log2(e)
  __half* out;        // [Be*seq, inner]
  long long* dbg_ts;  // optional [CTAs][16] phase-cycle trace (diagnostics; NULL in production)
  int turnstile;      // 1: serialise the exp2 loops of the two softmax warpgroups (ping-pong); 0: free-running
};

constexpr int kAttnThreads = 384;   // producer warpgroup (TMA warp, MMA warp, 2 idle) + 2 softmax warpgroups
constexpr int kAttnBQ = 128;        // rows per query tile (two tiles per CTA)
constexpr int kAttnBKV = 128;
constexpr int kAttnStages = 4;      // K and V rings
constexpr uint32_t kAttnTile = 128 * 64 * 2;  // 16 KB
// Q x2 + K,V rings + alignment slack + barriers (P lives in TMEM)
constexpr size_t kAttnSmem = size_t(kAttnTile) * (2 + 2 * kAttnStages) + 1024 + 512;

}  // namespace f5
