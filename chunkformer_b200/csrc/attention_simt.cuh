// Generic chunked relative-position attention (any c / l / r / d_k), CUDA-core version.
// Used for window shapes the tcgen05 kernel (attention_tc.cuh) is not specialised for (streaming presets such as
// chunk 16 / left 64 / right 0, d_k = 128 models). Same math as attention.py:420-505 + 104-150: per (chunk, head)
//   S[i,q] = ((Q_i+u).K_q + (Q_i+v).P[(c-1)-i+q]) / sqrt(d_k),  softmax over valid slots, context = A.V
// reading K/V windows by index from the flat buffer [cache l | frames | r zeros] (window of chunk g = rows
// [c*g, c*g+W)), so the 5x K/V unfold copy of the reference (attention.py:459-477) does not exist.
#pragma once
#include "common.cuh"

namespace cf {

struct AttnParams {
  const __nv_bfloat16* qkv;   // [l + frames + tail, 4d]: Qu | Qv | K | V ; buffer row = l + flat_frame
  const __nv_bfloat16* pos;   // [R(+pad), d] projected relative-position table of this layer
  const int2* range;          // [n_chunks] valid key slots [lo, hi)
  __nv_bfloat16* ctx;         // [frames, d]
  int n_chunks, c, l, r, d, heads;
  float scale;                // 1/sqrt(d_k)
  int prescaled;              // Q+u / Q+v already carry scale * log2(e) (folded into the QKV projection weights)
};

template <int DK>
__global__ void __launch_bounds__(128) attention_simt_kernel(AttnParams p) {
  extern __shared__ float at_smem[];
  const int W = p.l + p.c + p.r;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* s_q = at_smem + warp * (2 * DK + W);   // Qu[DK] | Qv[DK] | scores[W]
  float* s_s = s_q + 2 * DK;
  const int chunk = blockIdx.x, h = blockIdx.y;
  const int2 rg = p.range[chunk];
  const float k_log2 = p.prescaled ? 1.0f : p.scale * 1.4426950408889634f;   // scores live in the log2 domain
  const long long ld = 4LL * p.d;
  const __nv_bfloat16* kbase = p.qkv + (long long)chunk * p.c * ld + 2 * p.d + h * DK;
  const __nv_bfloat16* vbase = kbase + p.d;

  for (int i = warp; i < p.c; i += 4) {
    const __nv_bfloat16* qrow = p.qkv + ((long long)p.l + (long long)chunk * p.c + i) * ld + h * DK;
    for (int k = lane; k < DK; k += 32) {
      s_q[k] = __bfloat162float(qrow[k]);
      s_q[DK + k] = __bfloat162float(qrow[p.d + k]);
    }
    __syncwarp();
    float mx = -INFINITY;
    for (int q = rg.x + lane; q < rg.y; q += 32) {
      const uint4* kr = reinterpret_cast<const uint4*>(kbase + (long long)q * ld);
      const uint4* pr = reinterpret_cast<const uint4*>(p.pos + (long long)(p.c - 1 - i + q) * p.d + h * DK);
      float acc = 0.f;
#pragma unroll
      for (int k8 = 0; k8 < DK / 8; ++k8) {
        const uint4 kv = __ldg(kr + k8), pv = __ldg(pr + k8);
        const float* qu = s_q + k8 * 8;
        const float* qv = s_q + DK + k8 * 8;
        acc += qu[0] * bf16_lo(kv.x) + qu[1] * bf16_hi(kv.x) + qu[2] * bf16_lo(kv.y) + qu[3] * bf16_hi(kv.y) +
               qu[4] * bf16_lo(kv.z) + qu[5] * bf16_hi(kv.z) + qu[6] * bf16_lo(kv.w) + qu[7] * bf16_hi(kv.w);
        acc += qv[0] * bf16_lo(pv.x) + qv[1] * bf16_hi(pv.x) + qv[2] * bf16_lo(pv.y) + qv[3] * bf16_hi(pv.y) +
               qv[4] * bf16_lo(pv.z) + qv[5] * bf16_hi(pv.z) + qv[6] * bf16_lo(pv.w) + qv[7] * bf16_hi(pv.w);
      }
      acc *= k_log2;
      s_s[q] = acc;
      mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int q = rg.x + lane; q < rg.y; q += 32) {
      const float e = exp2f(s_s[q] - mx);
      s_s[q] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    const float inv = sum > 0.f ? 1.0f / sum : 0.f;   // no valid key: zero context (attention.py:133-136)
    float o[DK / 32];
#pragma unroll
    for (int j = 0; j < DK / 32; ++j) o[j] = 0.f;
    for (int q = rg.x; q < rg.y; ++q) {
      const float a = s_s[q];
      const __nv_bfloat16* vr = vbase + (long long)q * ld;
#pragma unroll
      for (int j = 0; j < DK / 32; ++j) o[j] = fmaf(a, __bfloat162float(vr[lane + 32 * j]), o[j]);
    }
    __nv_bfloat16* orow = p.ctx + ((long long)chunk * p.c + i) * p.d + h * DK;
#pragma unroll
    for (int j = 0; j < DK / 32; ++j) orow[lane + 32 * j] = __float2bfloat16(o[j] * inv);
    __syncwarp();
  }
}

}  // namespace cf
