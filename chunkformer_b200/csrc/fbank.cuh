// Kaldi-compatible log-mel filterbank on the GPU: the step right before the hot path.
// Replaces torchaudio.compliance.kaldi.fbank(waveform, num_mel_bins=80, frame_length=25, frame_shift=10, dither=0.0,
// energy_floor=0.0, sample_frequency=16000) as the reference calls it (chunkformer_model.py:307-315): snip_edges framing,
// DC-offset removal, pre-emphasis 0.97 with a replicated first sample, povey window, zero padding to 512, power spectrum,
// triangular mel filters (20 Hz .. Nyquist), log(max(e, FLT_EPSILON)).
// One warp per frame: the 400 samples go through shared memory once, a 256-point complex radix-2 FFT of the packed real
// frame (z[n] = x[2n] + i x[2n+1]) plus the real-FFT split gives the 257 power bins, and each lane sums the (sparse)
// triangular filters of up to three mel bins.  Everything fp32; bytes per frame: 640 B read (frame shift) + 320 B written.
#pragma once
#include "common.cuh"

namespace cf {

constexpr int FB_PAD = 512;            // padded window (round_to_power_of_two)
constexpr int FB_HALF = FB_PAD / 2;    // complex FFT length
constexpr int FB_WARPS = 8;
constexpr int FB_MAX_BINS = 96;        // mel bins (3 per lane)

struct FbankParams {
  const float* pcm;         // [n_samples]
  float* out;               // [n_frames, num_bins]
  const float* window;      // [frame_len]
  const float* mel_w;       // flat weights of the non-zero part of every mel filter
  const int2* mel_rng;      // [num_bins] (first power bin, offset into mel_w); count = next offset - offset
  const int* mel_cnt;       // [num_bins]
  long long n_frames;
  int frame_len, frame_shift, num_bins;
  float preemph;
};

__global__ void __launch_bounds__(FB_WARPS * 32) fbank_kernel(FbankParams p) {
  __shared__ float s_win[FB_PAD];
  __shared__ float2 s_tw[FB_HALF / 2];                 // exp(-2 pi i k / 256), k < 128
  __shared__ float2 s_tw512[FB_HALF + 1];              // exp(-2 pi i k / 512), k <= 256
  __shared__ float s_raw[FB_WARPS][FB_PAD];
  __shared__ float2 s_z[FB_WARPS][FB_HALF];
  __shared__ float s_pw[FB_WARPS][FB_HALF + 4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < FB_PAD; i += blockDim.x) s_win[i] = i < p.frame_len ? __ldg(p.window + i) : 0.f;
  for (int i = threadIdx.x; i < FB_HALF / 2; i += blockDim.x) {
    float sn, cs;
    sincospif(-2.0f * float(i) / float(FB_HALF), &sn, &cs);
    s_tw[i] = make_float2(cs, sn);
  }
  for (int i = threadIdx.x; i <= FB_HALF; i += blockDim.x) {
    float sn, cs;
    sincospif(-2.0f * float(i) / float(FB_PAD), &sn, &cs);
    s_tw512[i] = make_float2(cs, sn);
  }
  __syncthreads();
  float* raw = s_raw[warp];
  float2* z = s_z[warp];
  float* pw = s_pw[warp];
  // this lane's mel filters
  int m_first[3], m_cnt[3], m_off[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int b = lane + 32 * j;
    m_cnt[j] = 0; m_first[j] = 0; m_off[j] = 0;
    if (b < p.num_bins) { const int2 r = __ldg(p.mel_rng + b); m_first[j] = r.x; m_off[j] = r.y; m_cnt[j] = __ldg(p.mel_cnt + b); }
  }
  for (long long f = (long long)blockIdx.x * FB_WARPS + warp; f < p.n_frames; f += (long long)gridDim.x * FB_WARPS) {
    const float* src = p.pcm + f * p.frame_shift;
    // ---- load, DC offset
    float sum = 0.f;
    for (int j = lane; j < FB_PAD; j += 32) {
      const float v = j < p.frame_len ? __ldg(src + j) : 0.f;
      raw[j] = v;
      sum += v;
    }
    const float mean = warp_sum(sum) / float(p.frame_len);
    __syncwarp();
    // ---- pre-emphasis (first sample replicated), window, packed into the bit-reversed complex array
    for (int j = lane; j < FB_PAD; j += 32) {
      float v = 0.f;
      if (j < p.frame_len) {
        const float cur = raw[j] - mean;
        const float prev = raw[j > 0 ? j - 1 : 0] - mean;
        v = (cur - p.preemph * prev) * s_win[j];
      }
      const int n = j >> 1;
      const int br = int(__brev(unsigned(n)) >> 24);   // 8-bit reversal
      if (j & 1) z[br].y = v; else z[br].x = v;
    }
    __syncwarp();
    // ---- 256-point complex FFT, decimation in time
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const int half = 1 << s;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int t = lane + 32 * q;                   // butterfly 0..127
        const int pos = t & (half - 1);
        const int i0 = ((t >> s) << (s + 1)) + pos, i1 = i0 + half;
        const float2 w = s_tw[pos << (7 - s)];
        const float2 a = z[i0], b = z[i1];
        const float2 bw = make_float2(b.x * w.x - b.y * w.y, b.x * w.y + b.y * w.x);
        z[i0] = make_float2(a.x + bw.x, a.y + bw.y);
        z[i1] = make_float2(a.x - bw.x, a.y - bw.y);
      }
      __syncwarp();
    }
    // ---- real-FFT split: X[k] = (Z[k] + conj(Z[N-k]))/2 - i/2 * e^{-2 pi i k / 512} * (Z[k] - conj(Z[N-k])), power
    for (int k = lane; k <= FB_HALF; k += 32) {
      const float2 a = z[k & (FB_HALF - 1)], bq = z[(FB_HALF - k) & (FB_HALF - 1)];
      const float2 e = make_float2(0.5f * (a.x + bq.x), 0.5f * (a.y - bq.y));      // (Z[k] + conj(Z[N-k])) / 2
      const float2 o = make_float2(0.5f * (a.x - bq.x), 0.5f * (a.y + bq.y));      // (Z[k] - conj(Z[N-k])) / 2
      const float2 w = s_tw512[k];
      // -i * w * o
      const float2 wo = make_float2(w.x * o.x - w.y * o.y, w.x * o.y + w.y * o.x);
      const float xr = e.x + wo.y, xi = e.y - wo.x;
      pw[k] = xr * xr + xi * xi;
    }
    __syncwarp();
    // ---- mel filters + log
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int b = lane + 32 * j;
      if (b < p.num_bins) {
        float acc = 0.f;
        for (int i = 0; i < m_cnt[j]; ++i) acc = fmaf(__ldg(p.mel_w + m_off[j] + i), pw[m_first[j] + i], acc);
        p.out[f * p.num_bins + b] = logf(fmaxf(acc, 1.1920928955078125e-07f));
      }
    }
    __syncwarp();
  }
}

}  // namespace cf
