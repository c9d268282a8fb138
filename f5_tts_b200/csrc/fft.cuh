// Shared-memory radix-2 FFT kernels for the Vocos mel front-end (STFT -> |.| -> mel -> log) and the Vocos
// ISTFT head (exp/clip/cos/sin -> inverse real FFT -> window -> overlap-add -> envelope normalise).
// n_fft = 1024, hop = 256, hann(periodic) — the only configuration the reference ships
// (infer/utils_infer.py:52-57; configs/*.yaml mel_spec).  One CTA (256 threads) per frame; the whole transform
// lives in shared memory (8 KB data + 4 KB twiddles); global traffic is the algorithmic minimum.
#pragma once
#include "common.cuh"

namespace f5 {

constexpr int kNfft = 1024;
constexpr int kHop = 256;
constexpr int kBins = kNfft / 2 + 1;

__device__ __forceinline__ int bitrev10(int i) { return int(__brev(unsigned(i)) >> 22); }

// in-place forward DFT (e^{-i...}) of 1024 complex points held bit-reversed in (re, im); 256 threads.
__device__ __forceinline__ void fft1024_inplace(float* re, float* im, const float* twc, const float* tws) {
#pragma unroll 1
  for (int s = 1; s <= 10; ++s) {
    const int half = 1 << (s - 1);
    const int stride = kNfft >> s;  // twiddle stride
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int bf = threadIdx.x + u * 256;  // butterfly id 0..511
      const int j = bf & (half - 1);
      const int k = (bf >> (s - 1)) << s;
      const int i0 = k + j, i1 = i0 + half;
      const float wr = twc[j * stride], wi = -tws[j * stride];
      const float xr = re[i1], xi = im[i1];
      const float tr = xr * wr - xi * wi, ti = xr * wi + xi * wr;
      const float ar = re[i0], ai = im[i0];
      re[i0] = ar + tr;
      im[i0] = ai + ti;
      re[i1] = ar - tr;
      im[i1] = ai - ti;
    }
    __syncthreads();
  }
}

__device__ __forceinline__ void load_twiddles(float* twc, float* tws) {
  for (int i = threadIdx.x; i < kNfft / 2; i += blockDim.x) {
    float s, c;
    sincospif(2.0f * float(i) / float(kNfft), &s, &c);
    twc[i] = c;
    tws[i] = s;
  }
}

__device__ __forceinline__ float hann_periodic(int i) { return 0.5f - 0.5f * cospif(2.0f * float(i) / float(kNfft)); }

// ---------------------------------------------------------------------------------------------------------
// mel front-end (model/modules.py:80-109 -> torchaudio MelSpectrogram(power=1, center=True, norm=None, htk))
// wav [B, nw] -> mel; frame t covers reflect-padded samples [t*256 - 512, t*256 + 512).
// out_btc != 0 : out[b, t, m]  (layout CFM.sample wants, cfm.py:106-109), else out[b, m, t] (MelSpec.forward).
// fb: dense [513, n_mels] filterbank (constant, built on the host with torchaudio's formula).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mel_stft_kernel(const float* wav, int nw, int T, const float* fb, int n_mels,
                                                        float* out, int out_btc) {
  __shared__ float re[kNfft], im[kNfft], twc[kNfft / 2], tws[kNfft / 2];
  const int t = blockIdx.x, b = blockIdx.y;
  load_twiddles(twc, tws);
  const float* w = wav + (long long)b * nw;
  for (int i = threadIdx.x; i < kNfft; i += 256) {
    int s = t * kHop + i - kNfft / 2;
    if (s < 0) s = -s;
    if (s >= nw) s = 2 * (nw - 1) - s;
    const int d = bitrev10(i);
    re[d] = w[s] * hann_periodic(i);
    im[d] = 0.f;
  }
  __syncthreads();
  fft1024_inplace(re, im, twc, tws);
  // magnitudes -> im[0..512] (each thread touches only its own bin)
  for (int f = threadIdx.x; f < kBins; f += 256) {
    const float a = re[f], c = im[f];
    im[f] = sqrtf(a * a + c * c);
  }
  __syncthreads();
  for (int m = threadIdx.x; m < n_mels; m += 256) {
    float acc = 0.f;
    for (int f = 0; f < kBins; ++f) acc += im[f] * __ldg(fb + f * n_mels + m);
    const float v = logf(fmaxf(acc, 1e-5f));
    if (out_btc) out[((long long)b * T + t) * n_mels + m] = v;
    else out[((long long)b * n_mels + m) * T + t] = v;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Vocos ISTFT head (vocos ISTFTHead + torch.istft(center=True)).
// head: [B*T, 1026] fp32 = [log-mag (513) | phase (513)] per frame.
// frames out: [B*T, 1024] windowed time-domain frames; a second kernel overlap-adds (deterministic gather).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) istft_frames_kernel(const float* head, int ld, float* frames) {
  __shared__ float re[kNfft], im[kNfft], twc[kNfft / 2], tws[kNfft / 2];
  const long long fr = blockIdx.x;
  load_twiddles(twc, tws);
  const float* h = head + fr * ld;
  // X[k] = mag (cos p + i sin p), Hermitian extension; inverse via conj(FFT(conj(X))) / N.  Load conj(X).
  for (int k = threadIdx.x; k < kNfft; k += 256) {
    const int f = k <= 512 ? k : kNfft - k;
    const float mag = fminf(expf(h[f]), 100.0f);
    float s, c;
    sincosf(h[kBins + f], &s, &c);
    float xr = mag * c, xi = mag * s;
    if (k > 512) xi = -xi;           // Hermitian mirror
    if (f == 0 || f == 512) xi = 0;  // c2r ignores imag of DC / Nyquist
    const int d = bitrev10(k);
    re[d] = xr;
    im[d] = -xi;  // conj
  }
  __syncthreads();
  fft1024_inplace(re, im, twc, tws);
  for (int n = threadIdx.x; n < kNfft; n += 256) frames[fr * kNfft + n] = re[n] * (1.0f / kNfft) * hann_periodic(n);
}

// wav[b, i] = sum_t frames[b, t, i + 512 - 256 t] / sum_t hann^2[i + 512 - 256 t],  i in [0, 256 (T-1))
__global__ void istft_ola_kernel(const float* frames, int T, float* wav, int B) {
  const int L = kHop * (T - 1);
  const long long total = (long long)B * L;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int b = int(idx / L), i = int(idx % L);
    const int pos = i + kNfft / 2;
    const int t_hi = min(T - 1, pos / kHop);
    const int t_lo = pos >= kNfft - kHop ? (pos - (kNfft - kHop)) / kHop : 0;  // smallest t with pos - 256 t <= 1023
    float acc = 0.f, env = 0.f;
    for (int t = t_lo; t <= t_hi; ++t) {
      const int n = pos - t * kHop;
      if (n < 0 || n >= kNfft) continue;
      acc += frames[((long long)b * T + t) * kNfft + n];
      const float w = hann_periodic(n);
      env += w * w;
    }
    wav[idx] = acc / env;
  }
}

// Vocos embed Conv1d(100 -> 512, k=7, pad=3) as im2col: A[b*T + t, tap*C + c] = mel[b, c, t + tap - 3]  (fp16)
__global__ void vocos_im2col_kernel(const float* mel, int B, int C, int T, __half* A, int Kpad) {
  const long long row = blockIdx.x;
  const int b = int(row / T), t = int(row % T);
  for (int k = threadIdx.x; k < Kpad; k += blockDim.x) {
    float v = 0.f;
    if (k < 7 * C) {
      const int tap = k / C, c = k % C;
      const int tt = t + tap - 3;
      if (tt >= 0 && tt < T) v = mel[((long long)b * C + c) * T + tt];
    }
    A[row * Kpad + k] = __float2half_rn(v);
  }
}

}  // namespace f5
