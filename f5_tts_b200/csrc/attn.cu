// Host side of the tcgen05 attention kernel + C entry point f5_attention.
#include "attn.cuh"
#include "attn_splitkv.cuh"
#include "internal.h"

#include <cstdlib>
namespace f5 {
static long long* g_attn_trace = nullptr;

int attn_plan(AttnPlan* pl, const void* qkv, void* out, int batches, int seq, int heads, const int* kv_len,
              float scale) {
  if (batches <= 0 || seq <= 0 || heads <= 0) {
    set_error("attention: empty problem");
    return -1;
  }
  const int inner = heads * 64;
  int rc = encode_tmap_f16(&pl->tm, qkv, (uint64_t)3 * inner, (uint64_t)seq, (uint64_t)batches, (uint64_t)3 * inner * 2,
                           (uint64_t)seq * 3 * inner * 2, 64, 128, 3);
  if (rc) return rc;
  pl->p.seq = seq;
  pl->p.heads = heads;
  pl->p.batches = batches;
  pl->p.inner = inner;
  pl->p.kv_len = kv_len;
  pl->p.scale_log2 = scale * 1.4426950408889634f;
  pl->p.out = reinterpret_cast<__half*>(out);
  {
    static int ts = -1;
    if (ts < 0) {
      const char* e = getenv("F5_ATTN_TURNSTILE");
      ts = e ? atoi(e) : 1;  // measured: 22.6 us (on) vs 24.4 us (off) at Be=2, seq=938
    }
    pl->p.turnstile = ts;
    static int var = -1;
    if (var < 0) {
      const char* e = getenv("F5_ATTN_VARIANT");
      var = (e && atoi(e) == 6) ? 6 : 3;  // 6: experimental split-KV kernel (parity-green on B200, not timed yet)
    }
    pl->p.variant = var;
    if (var == 6) {
      rc = encode_tmap_f16(&pl->tm_kv64, qkv, (uint64_t)3 * inner, (uint64_t)seq, (uint64_t)batches, (uint64_t)3 * inner * 2,
                           (uint64_t)seq * 3 * inner * 2, 64, 64, 3);
      if (rc) return rc;
    }
  }
  {
    static long long* trace = nullptr;
    static int want = -1;
    if (want < 0) {
      want = getenv("F5_ATTN_TRACE") ? 1 : 0;
      if (want) cudaMalloc(&trace, sizeof(long long) * 16 * 4096);
    }
    pl->p.dbg_ts = trace;
    g_attn_trace = trace;
  }
  pl->grid = dim3((seq + 2 * kAttnBQ - 1) / (2 * kAttnBQ), heads, batches);
  return 0;
}

int attn_configure() {
  if (int rc = check_cuda(cudaFuncSetAttribute(attn_fwd_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)kAttnSmem),
                          "cudaFuncSetAttribute(attn smem)"))
    return rc;
  cudaFuncSetAttribute(attn_fwd_tcgen05_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  return 0;
}

int attn_run(const AttnPlan& pl, cudaStream_t s) {
  if (int rc = configure_kernels()) return rc;
  if (pl.p.variant == 6) {
    static int configured = 0;  // the experimental kernel is configured only when it is asked for
    if (!configured) {
      if (int rc = check_cuda(cudaFuncSetAttribute(attn_fwd_splitkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   (int)kSkvSmem),
                              "cudaFuncSetAttribute(split-KV attn smem)"))
        return rc;
      cudaFuncSetAttribute(attn_fwd_splitkv_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
      configured = 1;
    }
    PdlLaunch L(pl.grid, dim3(kSkvThreads), kSkvSmem, s);
    if (int rc = check_cuda(cudaLaunchKernelEx(&L.cfg, attn_fwd_splitkv_kernel, pl.tm, pl.tm_kv64, pl.p), "split-KV attention launch"))
      return rc;
  } else {
    PdlLaunch L(pl.grid, dim3(kAttnThreads), kAttnSmem, s);
    if (int rc = check_cuda(cudaLaunchKernelEx(&L.cfg, attn_fwd_tcgen05_kernel, pl.tm, pl.p), "attention launch")) return rc;
  }
  count_launch();
  return check_launch("attn_fwd_tcgen05_kernel launch");
}

}  // namespace f5

extern "C" int f5_debug_attn_trace(long long* host_out, int n_ctas) {
  if (!f5::g_attn_trace) return -1;
  cudaDeviceSynchronize();
  return cudaMemcpy(host_out, f5::g_attn_trace, sizeof(long long) * 16 * (size_t)n_ctas, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -2;
}

extern "C" int f5_attention(const void* qkv, void* out, int batches, int seq, int heads, const int* kv_len, float scale,
                            f5_stream_t stream) {
  f5::AttnPlan pl;
  if (int rc = f5::attn_plan(&pl, qkv, out, batches, seq, heads, kv_len, scale)) return rc;
  return f5::attn_run(pl, reinterpret_cast<cudaStream_t>(stream));
}
