// Host side of the tcgen05 attention kernel + C entry point f5_attention.
#include "attn.cuh"
#include "internal.h"

#include <cstdlib>
namespace f5 {
static long long* g_attn_trace = nullptr;

// The production library carries ONE kernel.  The instrumented build (make TRACE=1) also instantiates other MUFU / FMA
// splits of the exponentials (F5_ATTN_POLY = pairs of every 8 on the FMA pipe) and reads F5_ATTN_TURNSTILE, for A/B
// timing with tools/attn_bench.py.
#ifdef F5_TRACE
static int trace_poly() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("F5_ATTN_POLY");
    v = e ? atoi(e) : kPolyOf8;
  }
  return v;
}
#endif

typedef void (*AttnKernel)(const CUtensorMap, const AttnParams);
static AttnKernel attn_kernel() {
#ifdef F5_TRACE
  switch (trace_poly()) {
    case 0: return attn_fwd_tcgen05_kernel<0>;
    case 2: return attn_fwd_tcgen05_kernel<2>;
    case 4: return attn_fwd_tcgen05_kernel<4>;
    case 5: return attn_fwd_tcgen05_kernel<5>;
    default: break;
  }
#endif
  return attn_fwd_tcgen05_kernel<kPolyOf8>;
}

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
  pl->p.turnstile = 1;
  pl->p.dbg_ts = nullptr;
#ifdef F5_TRACE
  {
    static int ts = -1;
    if (ts < 0) {
      const char* e = getenv("F5_ATTN_TURNSTILE");
      ts = e ? atoi(e) : 1;
    }
    pl->p.turnstile = ts;
    static long long* trace = nullptr;
    static int want = -1;
    if (want < 0) {
      want = getenv("F5_ATTN_TRACE") ? 1 : 0;
      if (want) cudaMalloc(&trace, sizeof(long long) * 16 * 4096);
    }
    pl->p.dbg_ts = trace;
    g_attn_trace = trace;
  }
#endif
  pl->grid = dim3((seq + 2 * kAttnBQ - 1) / (2 * kAttnBQ), heads, batches);
  return 0;
}

static int attn_configure_one(AttnKernel k) {
  if (int rc = check_cuda(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAttnSmem),
                          "cudaFuncSetAttribute(attn smem)"))
    return rc;
  cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  return 0;
}

int attn_configure() {
#ifdef F5_TRACE
  if (int rc = attn_configure_one(attn_fwd_tcgen05_kernel<0>)) return rc;
  if (int rc = attn_configure_one(attn_fwd_tcgen05_kernel<2>)) return rc;
  if (int rc = attn_configure_one(attn_fwd_tcgen05_kernel<4>)) return rc;
  if (int rc = attn_configure_one(attn_fwd_tcgen05_kernel<5>)) return rc;
#endif
  return attn_configure_one(attn_fwd_tcgen05_kernel<kPolyOf8>);
}

int attn_run(const AttnPlan& pl, cudaStream_t s) {
  if (int rc = configure_kernels()) return rc;
  PdlLaunch L(pl.grid, dim3(kAttnThreads), kAttnSmem, s);
  if (int rc = check_cuda(cudaLaunchKernelEx(&L.cfg, attn_kernel(), pl.tm, pl.p), "attention launch")) return rc;
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
