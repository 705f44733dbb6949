// 8x depthwise-conv2d subsampling front-end (subsampling.py:70-112,128-175), stencil parts.
// The two pointwise 1x1 convs and the final Linear run on the tcgen05 GEMM (gemm.cuh); the kernels here produce
// their A operands directly in GEMM layout, reading the ragged feature buffer through the packer's chunk table, so
// the (n, 519, 80) chunk tensor of the reference (encoder.py:557-564, 606) is never materialised.
#pragma once
#include "common.cuh"

namespace cf {

struct ChunkSrc {      // one entry per chunk, built by the packer (plan.cpp)
  long long feat_row;  // row of the chunk's first input frame in the flat feature buffer
  int in_len;          // input frames present (rest is zero padding, encoder.py:557-561), may be <= 0
  int pad_;
};

// ---------------------------------------------------------------------------------------------
// conv0 (1->D, 3x3, stride 2) + ReLU + depthwise conv1 (3x3, stride 2, per channel), fused: each output position
// (t2, f2) of a chunk depends on a 7x7 input patch; the 3x3 intermediate is recomputed in registers and never stored.
//   out A1[(chunk, t2, f2), ch] bf16,  t2 in [0, T2), f2 in [0, F2)
// CTA: 128 output positions x all channels; input rows staged in smem (with CMVN applied after zero padding,
// encoder.py:615-616 / cmvn.py:40-43); weights staged in smem as [ch][20] = {w0[9], b0, w1[9], b1}.
// ---------------------------------------------------------------------------------------------
struct Fe1Params {
  const float* feats;        // [total_rows, feat_dim]
  const ChunkSrc* chunks;    // slab-local chunk table
  const float* wpack;        // [D][20]
  const float* cmvn_mean;    // nullable
  const float* cmvn_istd;
  __nv_bfloat16* out;        // [n_chunks * T2 * F2, D]
  int n_chunks, feat_dim, T2, F2, in_rows;  // in_rows = frames per chunk (8(c-1)+15)
};

template <int D>
__global__ void __launch_bounds__(128) frontend_conv0_dw1_kernel(Fe1Params p) {
  constexpr int POS = 128;
  constexpr int MAX_T2 = 9;                 // 128 positions span at most ceil(127/F2)+1 t2 rows (F2 >= 19 -> 8)
  constexpr int MAX_ROWS = 4 * MAX_T2 + 3;
  extern __shared__ float fe_smem[];
  float* s_w = fe_smem;                       // D*20
  float* s_in = s_w + D * 20;                 // MAX_ROWS * feat_dim
  uint32_t* s_out = reinterpret_cast<uint32_t*>(s_in + MAX_ROWS * p.feat_dim);  // [POS][33]

  const int tid = threadIdx.x;
  const int per_chunk = p.T2 * p.F2;
  const int blocks_per_chunk = (per_chunk + POS - 1) / POS;
  const int chunk = blockIdx.x / blocks_per_chunk;
  const int pos0 = (blockIdx.x - chunk * blocks_per_chunk) * POS;
  const int npos = min(POS, per_chunk - pos0);
  const ChunkSrc cs = p.chunks[chunk];

  for (int i = tid; i < D * 20; i += 128) s_w[i] = __ldg(p.wpack + i);
  const int t2_first = pos0 / p.F2;
  const int t2_last = (pos0 + npos - 1) / p.F2;
  const int row0 = 4 * t2_first;
  const int nrows = 4 * (t2_last - t2_first) + 7;
  for (int i = tid; i < nrows * p.feat_dim; i += 128) {
    const int rr = i / p.feat_dim, k = i - rr * p.feat_dim;
    const int t = row0 + rr;
    float v = 0.f;
    if (t < cs.in_len && t < p.in_rows) v = __ldg(p.feats + (cs.feat_row + t) * p.feat_dim + k);
    if (p.cmvn_mean) v = (v - __ldg(p.cmvn_mean + k)) * __ldg(p.cmvn_istd + k);
    s_in[i] = v;
  }
  __syncthreads();

  const int pos = pos0 + tid;
  const bool active = tid < npos;
  float patch[49];
  {
    const int t2 = active ? pos / p.F2 : t2_first;
    const int f2 = active ? pos - t2 * p.F2 : 0;
    const float* src = s_in + (4 * (t2 - t2_first)) * p.feat_dim + 4 * f2;
#pragma unroll
    for (int a = 0; a < 7; ++a)
#pragma unroll
      for (int b = 0; b < 7; ++b) patch[a * 7 + b] = src[a * p.feat_dim + b];
  }

  for (int cb = 0; cb < D; cb += 64) {
#pragma unroll 1
    for (int cc = 0; cc < 64; cc += 2) {
      float res[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const float4* wp = reinterpret_cast<const float4*>(s_w + (cb + cc + u) * 20);
        const float4 q0 = wp[0], q1 = wp[1], q2 = wp[2], q3 = wp[3], q4 = wp[4];
        const float w0[9] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x};
        const float b0 = q2.y;
        const float w1[9] = {q2.z, q2.w, q3.x, q3.y, q3.z, q3.w, q4.x, q4.y, q4.z};
        float acc = q4.w;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int b = 0; b < 3; ++b) {
            float s = b0;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
              for (int j = 0; j < 3; ++j) s = fmaf(w0[i * 3 + j], patch[(2 * a + i) * 7 + 2 * b + j], s);
            acc = fmaf(w1[a * 3 + b], fmaxf(s, 0.f), acc);
          }
        res[u] = acc;
      }
      s_out[tid * 33 + (cc >> 1)] = pack_bf16(res[0], res[1]);
    }
    __syncthreads();
    // coalesced copy-out: each warp writes rows of 64 channels (128 B)
    const int lane = tid & 31, warp = tid >> 5;
    for (int r = warp; r < npos; r += 4) {
      uint32_t* dst = reinterpret_cast<uint32_t*>(p.out + ((long long)chunk * per_chunk + pos0 + r) * D + cb);
      dst[lane] = s_out[r * 33 + lane];
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// depthwise conv2 (3x3, stride 2, per channel) on the pw1 output: [(chunk, t2, f2), ch] -> [(chunk, t3, f3), ch].
// Pure bandwidth: thread = 8 channels of one output position, 16-byte loads/stores. Weights as [9][D] + bias [D].
// ---------------------------------------------------------------------------------------------
struct Fe2Params {
  const __nv_bfloat16* in;   // [n_chunks * T2 * F2, D]
  __nv_bfloat16* out;        // [n_chunks * T3 * F3, D]
  const float* w;            // [9][D]
  const float* bias;         // [D]
  long long total;           // n_chunks * T3 * F3 * (D/8)
  int T2, F2, T3, F3;
};

template <int D>
__global__ void __launch_bounds__(256) frontend_dw2_kernel(Fe2Params p) {
  constexpr int G = D / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.total) return;
  const int cg = int(idx % G);
  long long o = idx / G;
  const int f3 = int(o % p.F3); o /= p.F3;
  const int t3 = int(o % p.T3);
  const long long chunk = o / p.T3;
  float acc[8];
  {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + cg * 8));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + cg * 8 + 4));
    acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w;
    acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
  }
  const long long in_base = chunk * p.T2 * p.F2;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const long long r = in_base + (long long)(2 * t3 + a) * p.F2 + (2 * f3 + b);
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.in + r * D + cg * 8));
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.w + (a * 3 + b) * D + cg * 8));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.w + (a * 3 + b) * D + cg * 8 + 4));
      acc[0] = fmaf(w0.x, bf16_lo(v.x), acc[0]); acc[1] = fmaf(w0.y, bf16_hi(v.x), acc[1]);
      acc[2] = fmaf(w0.z, bf16_lo(v.y), acc[2]); acc[3] = fmaf(w0.w, bf16_hi(v.y), acc[3]);
      acc[4] = fmaf(w1.x, bf16_lo(v.z), acc[4]); acc[5] = fmaf(w1.y, bf16_hi(v.z), acc[5]);
      acc[6] = fmaf(w1.z, bf16_lo(v.w), acc[6]); acc[7] = fmaf(w1.w, bf16_hi(v.w), acc[7]);
    }
  const long long orow = (chunk * p.T3 + t3) * p.F3 + f3;
  *reinterpret_cast<uint4*>(p.out + orow * D + cg * 8) =
      make_uint4(pack_bf16(acc[0], acc[1]), pack_bf16(acc[2], acc[3]), pack_bf16(acc[4], acc[5]), pack_bf16(acc[6], acc[7]));
}

}  // namespace cf
