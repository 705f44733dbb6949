// 8x depthwise-conv2d subsampling front-end (subsampling.py:70-112,128-175), stencil parts.
// The two pointwise 1x1 convs and the final Linear run on the tcgen05 GEMM (gemm.cuh); the kernels here produce
// their A operands directly in GEMM layout, reading the ragged feature buffer through the packer's chunk table, so
// the (n, 519, 80) chunk tensor of the reference (encoder.py:557-564, 606) is never materialised.
#pragma once
#include "common.cuh"
#include "gemm.cuh"

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
// Channel-major tensor-core version of conv0 + ReLU + depthwise conv1 (same output as the two kernels above).
// Work unit = one output time row (chunk, t2): the three conv0 rows t1 = 2*t2 + {0,1,2} it depends on are ONE UMMA per
// block of 128 channels, with the roles swapped relative to the kernel above:
//     G[ch, p] = W0'[ch, 0..15] . X[p, 0..15],   p = 3*f1 + r   (f1 < 39 conv0 frequency bins, r < 3 conv0 rows)
// so TMEM lanes are channels and TMEM columns are conv0 positions.  An epilogue thread owns one channel: its nine
// depthwise taps and bias live in registers for the whole kernel, every conv0 value is read from TMEM and rectified
// exactly once per unit (6.2 per output instead of 9), and the depthwise conv is 9 register FMAs per output.
// Outputs of neighbouring channels are paired with one shuffle so each lane stores a bf16x2.
//   warps [0, 4*NWG)   : epilogue, warpgroup g owns channels [128g, 128g+128) and TMEM columns [128g, 128g+128)
//   warp  4*NWG        : TMEM allocator + MMA issuer
//   warps 4*NWG + 1, 2 : producers: thread = conv0 frequency bin, 7x3 input patch -> three im2col rows (bf16)
// ---------------------------------------------------------------------------------------------
template <int D> constexpr int fecm_threads() { return (4 * (D / 128) + 3) * 32; }
constexpr int FECM_STAGES = 3;
constexpr int FECM_F1 = 39, FECM_F2 = 19;      // maxima (feat_dim <= 80)

template <int D>
__global__ void __launch_bounds__(fecm_threads<D>(), 1) frontend_conv0_dw1_cm_kernel(Fe1Params p, int total_units) {
  static_assert(D % 128 == 0, "D");
  constexpr int NWG = D / 128;
  constexpr int EPI_WARPS = 4 * NWG;
  extern __shared__ __align__(1024) uint8_t fec_smem[];
  uint8_t* sW = fec_smem;                                   // NWG x [128 ch x 16 k] bf16 (A operands)
  uint8_t* sB = sW + NWG * 4096;                            // FECM_STAGES x [128 pos x 16 k] bf16 (B operands)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + FECM_STAGES * 4096);
  uint64_t* b_full = bars;                                  // [STAGES] producers -> MMA (64 arrivals)
  uint64_t* b_empty = bars + FECM_STAGES;                   // [STAGES] MMA -> producers (commit)
  uint64_t* t_full = bars + 2 * FECM_STAGES;                // [NWG] MMA -> warpgroup (commit)
  uint64_t* t_free = bars + 2 * FECM_STAGES + NWG;          // [NWG] warpgroup -> MMA (4 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * FECM_STAGES + 2 * NWG);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int upc = (total_units + gridDim.x - 1) / gridDim.x;
  const int u_begin = blockIdx.x * upc;
  const int u_end = min(total_units, u_begin + upc);

  if (threadIdx.x == 0) {
    for (int s = 0; s < FECM_STAGES; ++s) { mbar_init(&b_full[s], 64); mbar_init(&b_empty[s], 1); }
    for (int g = 0; g < NWG; ++g) { mbar_init(&t_full[g], 1); mbar_init(&t_free[g], 4); }
    fence_barrier_init();
  }
  if (warp == EPI_WARPS) tmem_alloc(tmem_slot, NWG * 128);
  // conv0 weights -> A operand images (no-swizzle K-major core matrices: 8 rows x 16 B, K groups 128 B apart)
  for (int ch = threadIdx.x; ch < D; ch += blockDim.x) {
    const float* wp = p.wpack + ch * 20;
    const int blk = ch >> 7, row = ch & 127;
    __nv_bfloat16* k0 = reinterpret_cast<__nv_bfloat16*>(sW + blk * 4096 + (row >> 3) * 256 + (row & 7) * 16);
    __nv_bfloat16* k8 = k0 + 64;
#pragma unroll
    for (int k = 0; k < 8; ++k) k0[k] = __float2bfloat16(wp[k]);
    k8[0] = __float2bfloat16(wp[8]);
    k8[1] = __float2bfloat16(wp[9]);        // conv0 bias rides on the constant-1 input column
#pragma unroll
    for (int k = 2; k < 8; ++k) k8[k] = __float2bfloat16(0.f);
  }
  for (int i = threadIdx.x; i < FECM_STAGES * 4096 / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(sB)[i] = make_uint4(0u, 0u, 0u, 0u);   // positions past 3*F1 stay zero
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < EPI_WARPS) {
    // ------------------------------------------------------------------ epilogue: thread = channel
    const int g = warp >> 2, quad = warp & 3;
    const int ch = g * 128 + quad * 32 + lane;
    float w1[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) w1[t] = __ldg(p.wpack + ch * 20 + 10 + t);
    const float bias1 = __ldg(p.wpack + ch * 20 + 19);
    const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + g * 128;
    const int odd = lane & 1;
    __nv_bfloat16* out_ch = p.out + g * 128 + quad * 32 + (lane & ~1);
    uint32_t it = 0;
    for (int u = u_begin; u < u_end; ++u, ++it) {
      mbar_wait(&t_full[g], it & 1);
      tc_fence_after();
      __nv_bfloat16* orow = out_ch + (long long)u * FECM_F2 * D;
      uint32_t v[4][32];
      float acc = 0.f, acc_prev = 0.f, held = 0.f;
#pragma unroll
      for (int ck = 0; ck < 4; ++ck) {
        tmem_ld32(taddr + ck * 32, v[ck]);
        tmem_ld_wait();
        if (ck == 3) {                                    // every column of this unit is in registers
          tc_fence_before();
          if (lane == 0) mbar_arrive(&t_free[g]);
        }
#pragma unroll
        for (int f1 = 0; f1 < FECM_F1; ++f1) {
          if ((3 * f1 + 2) / 32 != ck) continue;          // handled when its last column arrives (compile-time)
          const float y0 = fmaxf(__uint_as_float(v[(3 * f1) / 32][(3 * f1) % 32]), 0.f);
          const float y1 = fmaxf(__uint_as_float(v[(3 * f1 + 1) / 32][(3 * f1 + 1) % 32]), 0.f);
          const float y2 = fmaxf(__uint_as_float(v[(3 * f1 + 2) / 32][(3 * f1 + 2) % 32]), 0.f);
          if ((f1 & 1) == 0) {
            const int k = f1 >> 1;                        // closes output k-1 (tap column 2), opens output k (column 0)
            if (k > 0) {
              acc_prev = fmaf(w1[2], y0, acc);
              acc_prev = fmaf(w1[5], y1, acc_prev);
              acc_prev = fmaf(w1[8], y2, acc_prev);
              const int ko = k - 1;
              if (ko & 1) {                               // pair (ko-1, ko) complete: exchange with the neighbour channel
                const float send = odd ? held : acc_prev;
                const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
                const uint32_t pk = odd ? pack_bf16(recv, acc_prev) : pack_bf16(held, recv);
                *reinterpret_cast<uint32_t*>(orow + (long long)(ko - 1 + odd) * D) = pk;
              } else {
                held = acc_prev;
              }
            }
            if (k < FECM_F2) {
              acc = fmaf(w1[0], y0, bias1);
              acc = fmaf(w1[3], y1, acc);
              acc = fmaf(w1[6], y2, acc);
            }
          } else {
            acc = fmaf(w1[1], y0, acc);
            acc = fmaf(w1[4], y1, acc);
            acc = fmaf(w1[7], y2, acc);
          }
        }
      }
      (orow + (long long)(FECM_F2 - 1) * D)[odd] = __float2bfloat16(held);   // F2 = 19 is odd: the last output has no partner
    }
  } else if (warp == EPI_WARPS) {
    // ------------------------------------------------------------------ MMA issuer (convergent warp, elected lane issues)
    {
      constexpr uint32_t idesc = make_idesc_bf16(128, 128);
      const uint64_t dw0 = make_nosw_desc(smem_u32(sW), 128, 256);
      const uint64_t db0 = make_nosw_desc(smem_u32(sB), 128, 256);
      uint32_t it = 0;
      for (int u = u_begin; u < u_end; ++u, ++it) {
        const int s = it % FECM_STAGES;
        mbar_wait(&b_full[s], (it / FECM_STAGES) & 1);
        tc_fence_after();
        const uint64_t db = db0 + uint64_t((s * 4096) >> 4);
#pragma unroll
        for (int g = 0; g < NWG; ++g) {
          mbar_wait(&t_free[g], (it & 1) ^ 1);
          tc_fence_after();
          if (elect_one()) {
            umma_bf16_ss(tmem_base + g * 128, dw0 + uint64_t((g * 4096) >> 4), db, idesc, 0);
            umma_commit(&t_full[g]);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(&b_empty[s]);
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ producers: thread = conv0 frequency bin
    const int f1 = threadIdx.x - (EPI_WARPS + 1) * 32;          // 0..63
    const int F1 = (p.feat_dim - 3) / 2 + 1;
    const bool live = f1 < F1 && f1 < FECM_F1;
    float mean[3] = {0.f, 0.f, 0.f}, istd[3] = {1.f, 1.f, 1.f};
    if (live && p.cmvn_mean) {
#pragma unroll
      for (int j = 0; j < 3; ++j) { mean[j] = __ldg(p.cmvn_mean + 2 * f1 + j); istd[j] = __ldg(p.cmvn_istd + 2 * f1 + j); }
    }
    auto load_patch = [&](int u, float (&x)[7][3]) {
      const int chunk = u / p.T2, t2 = u - chunk * p.T2;
      const ChunkSrc cs = p.chunks[chunk];
      const int lim = min(cs.in_len, p.in_rows);
      const float* src = p.feats + (cs.feat_row + 4 * t2) * p.feat_dim + 2 * f1;
#pragma unroll
      for (int i = 0; i < 7; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          float vv = 0.f;
          if (4 * t2 + i < lim) vv = __ldg(src + i * p.feat_dim + j);
          x[i][j] = (vv - mean[j]) * istd[j];               // CMVN after zero padding (encoder.py:615-616)
        }
    };
    float x[7][3], xn[7][3];
    if (live && u_begin < u_end) load_patch(u_begin, xn);
    uint32_t it = 0;
    for (int u = u_begin; u < u_end; ++u, ++it) {
#pragma unroll
      for (int i = 0; i < 7; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) x[i][j] = xn[i][j];
      if (live && u + 1 < u_end) load_patch(u + 1, xn);     // next unit's loads fly while this one is written
      const int s = it % FECM_STAGES;
      mbar_wait(&b_empty[s], ((it / FECM_STAGES) & 1) ^ 1);
      if (live) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int pos = 3 * f1 + r;
          uint8_t* dst = sB + s * 4096 + (pos >> 3) * 256 + (pos & 7) * 16;
          const uint4 lo = make_uint4(pack_bf16(x[2 * r][0], x[2 * r][1]), pack_bf16(x[2 * r][2], x[2 * r + 1][0]),
                                      pack_bf16(x[2 * r + 1][1], x[2 * r + 1][2]), pack_bf16(x[2 * r + 2][0], x[2 * r + 2][1]));
          const uint4 hi = make_uint4(pack_bf16(x[2 * r + 2][2], 1.0f), 0u, 0u, 0u);
          *reinterpret_cast<uint4*>(dst) = lo;
          *reinterpret_cast<uint4*>(dst + 128) = hi;
        }
      }
      fence_proxy_async();
      mbar_arrive(&b_full[s]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == EPI_WARPS) tmem_dealloc(tmem_base, NWG * 128);
}

template <int D>
constexpr size_t frontend_cm_smem_bytes() { return (D / 128) * 4096 + FECM_STAGES * 4096 + 256; }

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


// Row-walking version of depthwise conv2: thread = 8 channels of one output time row (chunk, t3); it keeps its 72 taps in
// registers and slides a 3x3 window of 16-byte vectors across the row's F2 input columns, so every input vector is
// loaded once per thread (6.3 loads per output instead of 9) and the taps are loaded once per row instead of once per
// output (the per-output tap loads were most of the L1 traffic of the kernel above).
template <int D>
__global__ void __launch_bounds__(192, 2) frontend_dw2_rows_kernel(Fe2Params p, long long total_rows) {
  constexpr int G = D / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total_rows * G) return;
  const int cg = int(idx % G);
  const long long orow = idx / G;                 // chunk * T3 + t3
  const int t3 = int(orow % p.T3);
  const long long chunk = orow / p.T3;
  float w[9][8], bias[8];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p.w + t * D + cg * 8));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p.w + t * D + cg * 8 + 4));
    w[t][0] = a.x; w[t][1] = a.y; w[t][2] = a.z; w[t][3] = a.w; w[t][4] = b.x; w[t][5] = b.y; w[t][6] = b.z; w[t][7] = b.w;
  }
  {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p.bias + cg * 8));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + cg * 8 + 4));
    bias[0] = a.x; bias[1] = a.y; bias[2] = a.z; bias[3] = a.w; bias[4] = b.x; bias[5] = b.y; bias[6] = b.z; bias[7] = b.w;
  }
  const __nv_bfloat16* in0 = p.in + ((chunk * p.T2 + 2 * t3) * p.F2) * D + cg * 8;   // input row 2*t3, column 0
  const long long rstride = (long long)p.F2 * D;                                      // one input time row
  __nv_bfloat16* out = p.out + (orow * p.F3) * D + cg * 8;
  uint4 win[3][3];                                                                   // [input row a][column slot]
  uint4 nxt[3][2];                                                                   // the two new columns of the next output
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    win[a][2] = __ldg(reinterpret_cast<const uint4*>(in0 + a * rstride));
    nxt[a][0] = __ldg(reinterpret_cast<const uint4*>(in0 + a * rstride + D));
    nxt[a][1] = __ldg(reinterpret_cast<const uint4*>(in0 + a * rstride + 2 * D));
  }
#pragma unroll 1
  for (int f3 = 0; f3 < p.F3; ++f3) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      win[a][0] = win[a][2];
      win[a][1] = nxt[a][0];
      win[a][2] = nxt[a][1];
    }
    if (f3 + 1 < p.F3) {                          // loads of the next output fly while this one is computed
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        nxt[a][0] = __ldg(reinterpret_cast<const uint4*>(in0 + a * rstride + (long long)(2 * f3 + 3) * D));
        nxt[a][1] = __ldg(reinterpret_cast<const uint4*>(in0 + a * rstride + (long long)(2 * f3 + 4) * D));
      }
    }
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = bias[e];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const uint4 v = win[a][b];
        const float* ww = w[a * 3 + b];
        acc[0] = fmaf(ww[0], bf16_lo(v.x), acc[0]); acc[1] = fmaf(ww[1], bf16_hi(v.x), acc[1]);
        acc[2] = fmaf(ww[2], bf16_lo(v.y), acc[2]); acc[3] = fmaf(ww[3], bf16_hi(v.y), acc[3]);
        acc[4] = fmaf(ww[4], bf16_lo(v.z), acc[4]); acc[5] = fmaf(ww[5], bf16_hi(v.z), acc[5]);
        acc[6] = fmaf(ww[6], bf16_lo(v.w), acc[6]); acc[7] = fmaf(ww[7], bf16_hi(v.w), acc[7]);
      }
    *reinterpret_cast<uint4*>(out + (long long)f3 * D) =
        make_uint4(pack_bf16(acc[0], acc[1]), pack_bf16(acc[2], acc[3]), pack_bf16(acc[4], acc[5]), pack_bf16(acc[6], acc[7]));
  }
}

}  // namespace cf
