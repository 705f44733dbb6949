// Bandwidth-bound kernels of the encoder layer: LayerNorm variants and the chunk-aware depthwise-conv core.
#pragma once
#include <type_traits>

#include "common.cuh"
#include "gemm.cuh"

namespace cf {

// ---------------------------------------------------------------------------------------------
// LayerNorm over d channels, one warp per row, fp32 statistics (eps 1e-5; encoder_layer.py:46-57).
//   MODE 0: y_bf16 = LN1(x)
//   MODE 1: x_f32 <- LN1(x) in place, y_bf16 = LN2(LN1(x))     (norm_final of layer i + norm_ff_macaron of layer i+1)
//   MODE 2: out = LN2(LN1(x))  (norm_final of the last layer + after_norm, encoder.py:670-671), fp32 and/or bf16 out
// Rows >= `zero_from` per batch element are written as zeros in MODE 0 when row_limit != nullptr (forward_encoder's
// x.masked_fill_ before pointwise_conv1, convolution.py:125-127).
// ---------------------------------------------------------------------------------------------
struct LnParams {
  const float* x_in;
  float* x_out;            // MODE 1: normalised residual stream (may alias x_in); MODE 2: fp32 output (nullable)
  __nv_bfloat16* y;        // bf16 output (MODE 2: nullable)
  const float* w1; const float* b1;
  const float* w2; const float* b2;
  long long rows;
  const int* row_limit;    // optional [rows / rows_per_seq]: rows with (row % rows_per_seq) >= limit are zeroed in y
  int rows_per_seq;
};

template <int D, int MODE>
__global__ void __launch_bounds__(256) layernorm_kernel(LnParams p) {
  constexpr int PER = D / 32;      // floats per lane
  constexpr int V4 = PER / 4;      // float4 per lane
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= p.rows) return;
  const float* xr = p.x_in + row * D;
  float v[PER];
#pragma unroll
  for (int k = 0; k < V4; ++k) {
    const float4 t = *reinterpret_cast<const float4*>(xr + k * 128 + lane * 4);
    v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
  }
  auto normalise = [&](const float* w, const float* b) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) s += v[i];
    const float mean = warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) { const float dlt = v[i] - mean; q += dlt * dlt; }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + 1e-5f);
#pragma unroll
    for (int k = 0; k < V4; ++k) {
      const float4 ww = __ldg(reinterpret_cast<const float4*>(w + k * 128 + lane * 4));
      const float4 bb = __ldg(reinterpret_cast<const float4*>(b + k * 128 + lane * 4));
      v[4 * k] = (v[4 * k] - mean) * rstd * ww.x + bb.x;
      v[4 * k + 1] = (v[4 * k + 1] - mean) * rstd * ww.y + bb.y;
      v[4 * k + 2] = (v[4 * k + 2] - mean) * rstd * ww.z + bb.z;
      v[4 * k + 3] = (v[4 * k + 3] - mean) * rstd * ww.w + bb.w;
    }
  };
  normalise(p.w1, p.b1);
  if (MODE == 1) {
    float* xo = p.x_out + row * D;
#pragma unroll
    for (int k = 0; k < V4; ++k)
      *reinterpret_cast<float4*>(xo + k * 128 + lane * 4) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
  }
  if (MODE >= 1) normalise(p.w2, p.b2);
  if (MODE == 2 && p.x_out) {
    float* xo = p.x_out + row * D;
#pragma unroll
    for (int k = 0; k < V4; ++k)
      *reinterpret_cast<float4*>(xo + k * 128 + lane * 4) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
  }
  if (p.y) {
    bool zero = false;
    if (MODE == 0 && p.row_limit) {
      const long long s = row / p.rows_per_seq;
      zero = (row - s * p.rows_per_seq) >= p.row_limit[s];
    }
    __nv_bfloat16* yr = p.y + row * D;
#pragma unroll
    for (int k = 0; k < V4; ++k) {
      uint2 o;
      o.x = zero ? 0u : pack_bf16(v[4 * k], v[4 * k + 1]);
      o.y = zero ? 0u : pack_bf16(v[4 * k + 2], v[4 * k + 3]);
      *reinterpret_cast<uint2*>(yr + k * 128 + lane * 4) = o;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Chunk-aware depthwise conv core (convolution.py:224-249): for every chunk, 15-tap depthwise conv over the GLU
// output with context halos read by index from the flat buffer (no unfold copy), + bias, LayerNorm over channels,
// SiLU.  One CTA = FG consecutive frames of one chunk x all D channels; thread = 2 adjacent channels.
//   g      : [halo + frames + halo, D] bf16, row (lorder + flat_frame)
//   window slot q of chunk (frame c*chunk - lorder + q) contributes iff range[chunk].x <= q < range[chunk].y
//   z      : [frames, D] bf16
// Bytes per frame: read 2*D (each row once from HBM; halo re-reads hit L2), write 2*D.
// ---------------------------------------------------------------------------------------------
struct DwConvParams {
  const __nv_bfloat16* g;
  __nv_bfloat16* z;
  const float* w;       // [D, KW] depthwise taps (depthwise_conv.weight (d,1,15))
  const float* bias;    // [D]
  const float* ln_w; const float* ln_b;
  const int2* range;    // [n_chunks] valid window-slot range
  int c;                // chunk size (frames)
  int n_chunks;
  int sub_chunk;        // > 0 (streaming with right context, generic kernel only): output frame f of a chunk only sees window
                        // slots left of the end of its sub-chunk of `sub_chunk` frames, q < (f / sub_chunk + 1) * sub_chunk + lorder
                        // (convolution.py:150-167 with chunk_size = sub_chunk on a sequence of chunk + right-context frames)
  int no_norm;          // 1: no LayerNorm (cnn_module_norm = batch_norm: the eval-mode affine is folded into w / bias), z = SiLU(conv)
  int chunk_stride = 1, chunk_first = 0;   // generic kernel: work item i is chunk chunk_first + i * chunk_stride (compact streaming
                                           // visits only the real chunk of every stream); n_chunks counts the visited chunks
};


// Sum v[f] over all threads of the CTA for every frame f; result (scaled, optionally rsqrt(x + eps)) lands in s_out[f].
// FG == 32 uses a lane-transposing butterfly (31 shuffles for 32 values) instead of 32 full warp reductions.
template <int FG, int NW>
CF_DEVINL void frame_reduce(float (&v)[FG], float (*s_part)[FG], float* s_out, float scale, bool to_rstd) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (FG == 32) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const bool up = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < off; ++i) {
        const float send = up ? v[i] : v[i + off];
        const float keep = up ? v[i + off] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
    }
    s_part[warp][lane] = v[0];
  } else {
#pragma unroll
    for (int f = 0; f < FG; ++f) {
      const float a = warp_sum(v[f]);
      if (lane == 0) s_part[warp][f] = a;
    }
  }
  __syncthreads();
  if (threadIdx.x < FG) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) a += s_part[w][threadIdx.x];
    a *= scale;
    s_out[threadIdx.x] = to_rstd ? rsqrtf(a + 1e-5f) : a;
  }
  __syncthreads();
}

template <int D, int KW, int FG>
__global__ void __launch_bounds__(D / 2) dwconv_ln_silu_kernel(DwConvParams p) {
  constexpr int NT = D / 2;
  constexpr int NW = NT / 32;
  constexpr int ROWS = FG + KW - 1;
  __shared__ float s_part[NW][FG];
  __shared__ float s_mean[FG];
  __shared__ float s_rstd[FG];

  const int tid = threadIdx.x;
  const int groups = p.c / FG;
  const int item = blockIdx.x / groups;
  const int chunk = p.chunk_first + item * p.chunk_stride;
  const int f0 = (blockIdx.x - item * groups) * FG;    // first frame of this group inside the chunk
  const int2 rg = p.range[chunk];
  const long long base_row = (long long)chunk * p.c + f0;  // buffer row of window slot (f0 + 0)

  // all loads up front: ROWS independent 4-byte loads per thread
  uint32_t in[ROWS];
  const uint32_t* gp = reinterpret_cast<const uint32_t*>(p.g) + base_row * (D / 2) + tid;
#pragma unroll
  for (int t = 0; t < ROWS; ++t) {
    const int q = f0 + t;
    in[t] = (q >= rg.x && q < rg.y) ? __ldg(gp + (long long)t * (D / 2)) : 0u;
  }
  float w0[KW], w1[KW];
#pragma unroll
  for (int t = 0; t < KW; ++t) {
    w0[t] = __ldg(p.w + (2 * tid) * KW + t);
    w1[t] = __ldg(p.w + (2 * tid + 1) * KW + t);
  }
  const float bias0 = __ldg(p.bias + 2 * tid), bias1 = __ldg(p.bias + 2 * tid + 1);

  float o0[FG], o1[FG];
#pragma unroll
  for (int f = 0; f < FG; ++f) {
    float a0 = bias0, a1 = bias1;
    // first window slot this output frame may NOT see (right edge of its sub-chunk + lorder); no limit without sub-chunks
    const int lim = p.sub_chunk > 0 ? ((f0 + f) / p.sub_chunk + 1) * p.sub_chunk + KW / 2 : 0x7fffffff;
#pragma unroll
    for (int t = 0; t < KW; ++t) {
      const bool ok = f0 + f + t < lim;
      a0 = fmaf(w0[t], ok ? bf16_lo(in[f + t]) : 0.f, a0);
      a1 = fmaf(w1[t], ok ? bf16_hi(in[f + t]) : 0.f, a1);
    }
    o0[f] = a0; o1[f] = a1;
  }

  // LayerNorm statistics per frame over the D channels, two-pass (mean, then squared deviations) in fp32
  float ps[FG];
#pragma unroll
  for (int f = 0; f < FG; ++f) ps[f] = o0[f] + o1[f];
  frame_reduce<FG, NW>(ps, s_part, s_mean, 1.0f / D, false);
#pragma unroll
  for (int f = 0; f < FG; ++f) {
    const float m = s_mean[f];
    const float d0 = o0[f] - m, d1 = o1[f] - m;
    ps[f] = d0 * d0 + d1 * d1;
  }
  frame_reduce<FG, NW>(ps, s_part, s_rstd, 1.0f / D, true);
  const float lw0 = __ldg(p.ln_w + 2 * tid), lw1 = __ldg(p.ln_w + 2 * tid + 1);
  const float lb0 = __ldg(p.ln_b + 2 * tid), lb1 = __ldg(p.ln_b + 2 * tid + 1);
  uint32_t* zp = reinterpret_cast<uint32_t*>(p.z) + ((long long)chunk * p.c + f0) * (D / 2) + tid;
#pragma unroll
  for (int f = 0; f < FG; ++f) {
    const float mean = p.no_norm ? 0.f : s_mean[f], rstd = p.no_norm ? 1.f : s_rstd[f];
    const float y0 = (o0[f] - mean) * rstd * lw0 + lb0;
    const float y1 = (o1[f] - mean) * rstd * lw1 + lb1;
    zp[(long long)f * (D / 2)] = pack_bf16(silu(y0), silu(y1));
  }
}


// ---------------------------------------------------------------------------------------------
// TMA-staged version for chunk sizes that are multiples of 32 (the benchmark's chunk 64): persistent CTAs, the (32+14)
// halo'd input rows of a 32-frame group arrive in shared memory by cp.async.bulk.tensor (two-stage ring, the load of group
// i+1 overlaps the math of group i), each thread streams its two channels through a 15-row sliding register window,
// packed fp32x2 FMAs (fma.rn.f32x2), LayerNorm statistics as above, SiLU through one tanh.approx.
// ---------------------------------------------------------------------------------------------
CF_DEVINL unsigned long long pack_f32x2(float lo, float hi) {
  return (unsigned long long)__float_as_uint(lo) | ((unsigned long long)__float_as_uint(hi) << 32);
}
CF_DEVINL unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
CF_DEVINL unsigned long long fadd2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
CF_DEVINL float f32x2_lo(unsigned long long v) { return __uint_as_float((unsigned)(v & 0xffffffffull)); }
CF_DEVINL float f32x2_hi(unsigned long long v) { return __uint_as_float((unsigned)(v >> 32)); }

template <int D, int FG, int STAGES>
constexpr size_t dwconv_tma_smem_bytes() { return size_t(STAGES) * (FG + 14) * D * 2 + 1024; }

// FG = frames per group: 32 (two CTAs per SM, 128 registers) or 16 (three CTAs per SM, 84 registers, 30-row tiles)
template <int D, int FG, int STAGES>
__global__ void __launch_bounds__(D / 2, FG == 16 ? 3 : 2)
dwconv_ln_silu_tma_kernel(const __grid_constant__ CUtensorMap tma_g, DwConvParams p, int total_groups) {
  constexpr int KW = 15, ROWS = FG + KW - 1;
  constexpr int NT = D / 2, NW = NT / 32;
  constexpr int HALVES = D / 256;                 // TMA boxes of 256 channels
  constexpr uint32_t STAGE_BYTES = ROWS * D * 2;
  extern __shared__ __align__(1024) uint8_t dw_smem[];
  uint8_t* s_in = dw_smem;                        // [STAGES][HALVES][ROWS][256 ch] bf16
  uint64_t* full = reinterpret_cast<uint64_t*>(s_in + STAGES * STAGE_BYTES);
  __shared__ unsigned long long s_part2[2][NW][FG];   // double-buffered by iteration parity: no barrier at the end of a group

  const int tid = threadIdx.x;
  const int groups_per_chunk = p.c / FG;
  if (tid == 0) {
    tma_prefetch_desc(&tma_g);
#pragma unroll
    for (int i = 0; i < STAGES; ++i) mbar_init(&full[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](int grp, int stage) {
    const int chunk = grp / groups_per_chunk;
    const int f0 = (grp - chunk * groups_per_chunk) * FG;
    const int row = chunk * p.c + f0;             // buffer row of window slot f0
    mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
#pragma unroll
    for (int hb = 0; hb < HALVES; ++hb)
      tma_load_2d(s_in + stage * STAGE_BYTES + hb * (ROWS * 512), &tma_g, &full[stage], hb * 256, row);
  };
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < STAGES - 1; ++i)
      if (int(blockIdx.x) + i * int(gridDim.x) < total_groups) issue(blockIdx.x + i * gridDim.x, i);
  }

  unsigned long long w[KW];
#pragma unroll
  for (int t = 0; t < KW; ++t) w[t] = pack_f32x2(__ldg(p.w + (2 * tid) * KW + t), __ldg(p.w + (2 * tid + 1) * KW + t));
  const unsigned long long bias2 = pack_f32x2(__ldg(p.bias + 2 * tid), __ldg(p.bias + 2 * tid + 1));
  // SiLU(x) = h + h tanh(h), h = x / 2: the halving is folded into the LayerNorm affine
  const unsigned long long lwh2 = pack_f32x2(0.5f * __ldg(p.ln_w + 2 * tid), 0.5f * __ldg(p.ln_w + 2 * tid + 1));
  const unsigned long long lbh2 = pack_f32x2(0.5f * __ldg(p.ln_b + 2 * tid), 0.5f * __ldg(p.ln_b + 2 * tid + 1));
  const int hb = tid >> 7;                        // which 256-channel box this thread's channels live in
  const uint32_t col_off = hb * (ROWS * 512) + (tid & 127) * 4;
  const int lane = tid & 31, warp = tid >> 5;

  int2 rg_next = (int(blockIdx.x) < total_groups) ? __ldg(&p.range[blockIdx.x / groups_per_chunk]) : make_int2(0, 0);
  int it = 0;
  for (int grp = blockIdx.x; grp < total_groups; grp += gridDim.x, ++it) {
    const int stage = it % STAGES;
    const int par = it & 1;                       // parity of the double-buffered statistics
    const int nxt = grp + gridDim.x;
    // the tile of group it + STAGES - 1 goes where group it - 1 was: every thread finished reading it before the LayerNorm
    // barrier of the previous iteration
    if (tid == 0 && grp + (STAGES - 1) * int(gridDim.x) < total_groups)
      issue(grp + (STAGES - 1) * gridDim.x, (it + STAGES - 1) % STAGES);
    const int chunk = grp / groups_per_chunk;
    const int f0 = (grp - chunk * groups_per_chunk) * FG;
    const int2 rg = rg_next;
    if (nxt < total_groups) rg_next = __ldg(&p.range[nxt / groups_per_chunk]);   // consumed one iteration later
    mbar_wait(&full[stage], (it / STAGES) & 1);
    const uint8_t* src = s_in + stage * STAGE_BYTES + col_off;

    // Row-outer order: input row s feeds the 15 frames s-14 .. s, so the 15 FMAs issued per loaded row are independent
    // (no sliding register window, no serial accumulator chain); frame f still sums bias, tap 0, ..., tap 14 in that order.
    unsigned long long o[FG];
    auto conv = [&](auto masked_tag) {
      constexpr bool MASKED = decltype(masked_tag)::value;
#pragma unroll
      for (int s = 0; s < ROWS; ++s) {
        const uint32_t v = *reinterpret_cast<const uint32_t*>(src + s * 512);
        const bool ok = !MASKED || ((f0 + s >= rg.x) && (f0 + s < rg.y));
        const unsigned long long x = ok ? pack_f32x2(bf16_lo(v), bf16_hi(v)) : 0ull;
#pragma unroll
        for (int t = 0; t < KW; ++t) {
          const int f = s - t;
          if (f >= 0 && f < FG) o[f] = ffma2(w[t], x, t == 0 ? bias2 : o[f]);
        }
      }
    };
    // groups whose 46 window slots are all valid (the bulk of a long utterance) skip the per-row mask tests
    if (f0 >= rg.x && f0 + ROWS <= rg.y) conv(std::false_type{}); else conv(std::true_type{});

    // LayerNorm over the D channels of every frame: (sum, sum of squares) travel as one f32x2 pair through a
    // lane-transposing butterfly (lane f ends with frame f's warp total), one barrier pair per group
    unsigned long long P[FG];
#pragma unroll
    for (int f = 0; f < FG; ++f) {
      const float lo = f32x2_lo(o[f]), hi = f32x2_hi(o[f]);
      P[f] = pack_f32x2(lo + hi, fmaf(lo, lo, hi * hi));
    }
#pragma unroll
    for (int off = FG / 2; off >= 1; off >>= 1) {
      const bool up = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < off; ++i) {
        const unsigned long long send = up ? P[i] : P[i + off];
        const unsigned long long keep = up ? P[i + off] : P[i];
        P[i] = fadd2(keep, __shfl_xor_sync(0xffffffffu, send, off));
      }
    }
#pragma unroll
    for (int off = FG; off < 32; off <<= 1) P[0] = fadd2(P[0], __shfl_xor_sync(0xffffffffu, P[0], off));
    if (lane < FG) s_part2[par][warp][lane] = P[0];
    __syncthreads();
    // every warp finishes the statistics itself (lane f owns frame f), the output loop fetches them by shuffle: one
    // barrier per group, no single-warp finalize step on the critical path
    float rs, nm;
    {
      float a = 0.f, q = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < NW; ++w2) { const unsigned long long v = s_part2[par][w2][lane & (FG - 1)]; a += f32x2_lo(v); q += f32x2_hi(v); }
      const float mean = a * (1.0f / D);
      const float var = fmaxf(q * (1.0f / D) - mean * mean, 0.f);
      rs = rsqrtf(var + 1e-5f);
      nm = -mean * rs;
      if (p.no_norm) { rs = 1.f; nm = 0.f; }
    }
    uint32_t* zp = reinterpret_cast<uint32_t*>(p.z) + ((long long)chunk * p.c + f0) * (D / 2) + tid;
#pragma unroll
    for (int f = 0; f < FG; ++f) {
      const float rf = __shfl_sync(0xffffffffu, rs, f), mf = __shfl_sync(0xffffffffu, nm, f);
      const unsigned long long h = ffma2(ffma2(o[f], pack_f32x2(rf, rf), pack_f32x2(mf, mf)), lwh2, lbh2);   // half the LayerNorm output
      const unsigned long long th = pack_f32x2(tanh_approx(f32x2_lo(h)), tanh_approx(f32x2_hi(h)));
      const unsigned long long z = ffma2(h, th, h);
      zp[(long long)f * (D / 2)] = pack_bf16(f32x2_lo(z), f32x2_hi(z));
    }
  }
}

}  // namespace cf
