// Small helper kernels: streaming-cache import/export, CTC partial reduction, row log-softmax.
#pragma once
#include "common.cuh"

namespace cf {

// att_cache (l, H, 2*d_k) fp32 of one layer  <->  K/V columns of the first rows of the flat QKV buffer
// (attention.py:459-467: cache rows precede all frames; K in [..., :d_k], V in [..., d_k:]).
__global__ void att_cache_import_kernel(const float* cache, __nv_bfloat16* qkv, int l, int H, int dk, int d) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = l * H * 2 * dk;
  if (idx >= total) return;
  const int e = idx % (2 * dk);
  const int h = (idx / (2 * dk)) % H;
  const int t = idx / (2 * dk * H);
  const int col = (e < dk) ? (2 * d + h * dk + e) : (3 * d + h * dk + (e - dk));
  qkv[(long long)t * 4 * d + col] = __float2bfloat16(cache[idx]);
}
// new cache = rows [trunc, trunc + l) of the buffer (attention.py:466-467)
__global__ void att_cache_export_kernel(float* cache, const __nv_bfloat16* qkv, int l, int H, int dk, int d, int trunc) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = l * H * 2 * dk;
  if (idx >= total) return;
  const int e = idx % (2 * dk);
  const int h = (idx / (2 * dk)) % H;
  const int t = idx / (2 * dk * H);
  const int col = (e < dk) ? (2 * d + h * dk + e) : (3 * d + h * dk + (e - dk));
  cache[idx] = __bfloat162float(qkv[(long long)(trunc + t) * 4 * d + col]);
}
// cnn_cache (d, lorder) fp32 of one layer <-> first rows of the GLU buffer (convolution.py:224-232)
__global__ void cnn_cache_import_kernel(const float* cache, __nv_bfloat16* g, int d, int lo) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= d * lo) return;
  const int t = idx % lo, ch = idx / lo;
  g[(long long)t * d + ch] = __float2bfloat16(cache[idx]);
}
__global__ void cnn_cache_export_kernel(float* cache, const __nv_bfloat16* g, int d, int lo, int trunc) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= d * lo) return;
  const int t = idx % lo, ch = idx / lo;
  cache[idx] = __bfloat162float(g[(long long)(trunc + t) * d + ch]);
}

// ---------------------------------------------------------------------------------------------
// Multi-stream caches (frame-synchronous streaming, encoder.py:310-390): B streams in one pass.  Stream s is an utterance of
// `ph` placeholder chunks + one real chunk in the flat buffers (S = (ph + 1) * c rows per stream); after the projection GEMM
// of a layer the last l (attention) / lo (conv) placeholder rows are overwritten with that stream's cache, so the real
// chunk's window reads [cache | new frames] exactly as forward_chunk concatenates them; the new cache is the last l / lo
// rows of cache + frames (encoder.py:376-388).  Layouts as the reference passes them: att (B, H, l, 2 d_k) and cnn (B, d, lo)
// per layer.  `first` = buffer row of flat frame 0 (l for the K/V buffer, lo for the GLU buffer).
// ---------------------------------------------------------------------------------------------
// Thread = 8 consecutive cache elements (32 bytes of fp32 <-> one 16-byte bf16 vector of a K or V row; 8 divides d_k, so a group
// never straddles the K / V halves): the element-per-thread version with four 64-bit divisions per element ran at 1.1 TB/s and
// was 44 % of a 256-stream step.
__global__ void att_cache_streams_kernel(float* cache, __nv_bfloat16* qkv, int B, int l, int H, int dk, int d, int S, int ph_rows,
                                         int c, int first, int do_export) {
  const long long idx8 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int per_row = (2 * dk) >> 3;
  const long long total8 = (long long)B * H * l * per_row;
  if (idx8 >= total8) return;
  const int e = int(idx8 % per_row) << 3;
  const long long rest = idx8 / per_row;          // (s * H + h) * l + t
  const int t = int(rest % l);
  const int sh = int(rest / l);
  const int h = sh % H, s = sh / H;
  const int col = (e < dk) ? (2 * d + h * dk + e) : (3 * d + h * dk + (e - dk));
  const long long row = (long long)first + (long long)s * S + ph_rows - l + t + (do_export ? c : 0);
  float4* cp = reinterpret_cast<float4*>(cache + idx8 * 8);
  uint4* qp = reinterpret_cast<uint4*>(qkv + row * 4 * d + col);
  if (do_export) {
    const uint4 v = *qp;
    cp[0] = make_float4(bf16_lo(v.x), bf16_hi(v.x), bf16_lo(v.y), bf16_hi(v.y));
    cp[1] = make_float4(bf16_lo(v.z), bf16_hi(v.z), bf16_lo(v.w), bf16_hi(v.w));
  } else {
    const float4 a = cp[0], b = cp[1];
    *qp = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
  }
}
__global__ void cnn_cache_streams_kernel(float* cache, __nv_bfloat16* g, int B, int d, int lo, int S, int ph_rows, int c, int first,
                                         int do_export) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * d * lo) return;
  const int t = int(idx % lo);
  const int ch = int((idx / lo) % d);
  const int s = int(idx / ((long long)lo * d));
  const long long row = (long long)first + (long long)s * S + ph_rows - lo + t + (do_export ? c : 0);
  if (do_export) cache[idx] = __bfloat162float(g[row * d + ch]);
  else g[row * d + ch] = __float2bfloat16(cache[idx]);
}

// Greedy CTC: combine per-tile (best, runner-up, index) partials written by the GEMM epilogue; ties resolve to the
// lowest index like torch.argmax (ctc.py:83-91).
__global__ void ctc_reduce_kernel(const float* best, const float* second, const int* index, int n_tiles, long long rows,
                                  long long* tokens, float* margin) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  float b = -INFINITY, s = -INFINITY;
  int bi = 0;
  for (int t = 0; t < n_tiles; ++t) {
    const float tb = best[row * n_tiles + t], ts = second[row * n_tiles + t];
    if (tb > b) { s = fmaxf(b, ts); b = tb; bi = index[row * n_tiles + t]; }
    else { s = fmaxf(s, tb); }
  }
  tokens[row] = bi;
  if (margin) margin[row] = b - s;
}

// In-place row log-softmax over V fp32 logits (ctc.py:81), one warp per row.
__global__ void log_softmax_rows_kernel(float* x, long long rows, int V) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float* r = x + row * V;
  float mx = -INFINITY;
  for (int i = lane; i < V; i += 32) mx = fmaxf(mx, r[i]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int i = lane; i < V; i += 32) s += __expf(r[i] - mx);
  s = warp_sum(s);
  const float lse = mx + __logf(s);
  for (int i = lane; i < V; i += 32) r[i] -= lse;
}

// ---------------------------------------------------------------------------------------------
// Device-side CTC token compaction (utils/model_utils.py:23-32 and the blank filter of :186-196).
// Greedy token ids stay on the device as a flat row buffer in which utterance s owns rows [seg_start[s], seg_start[s] +
// seg_len[s]) (chunk padding rows in between belong to nobody).  mode 0 keeps a row when its token is not blank and differs
// from the previous row of the same utterance (remove_duplicates_and_blank); mode 1 keeps every non-blank row (the frames
// get_output_with_timestamps collects before it collapses each segment).  Kept (token, frame-in-utterance) pairs are
// written compactly in row order; out_offsets[s] is the position of utterance s's first pair, out_offsets[n_seg] the total.
// Two launches, no spin-waits: per-block keep counts, then every block sums the counts of the blocks before it (at most a
// few hundred values), scans its own 2048 rows and scatters.
// ---------------------------------------------------------------------------------------------
constexpr int CTC_COMPACT_THREADS = 256;
constexpr int CTC_COMPACT_ITEMS = 8;
constexpr int CTC_COMPACT_TILE = CTC_COMPACT_THREADS * CTC_COMPACT_ITEMS;

struct CtcCompactParams {
  const long long* tokens;
  long long rows;
  const long long* seg_start;   // [n_seg] ascending
  const int* seg_len;           // [n_seg]
  int n_seg;
  int mode;
  long long blank;
  long long* out_tokens;
  int* out_frames;
  long long* out_offsets;       // [n_seg + 1]
  int* block_counts;            // [num_blocks]
};

// last segment whose first row is <= row (-1 if none)
CF_DEVINL int ctc_segment_of(const long long* seg_start, int n_seg, long long row) {
  int lo = 0, hi = n_seg;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(seg_start + mid) <= row) lo = mid + 1; else hi = mid;
  }
  return lo - 1;
}

CF_DEVINL bool ctc_keep(const CtcCompactParams& p, long long row, int* seg_out, long long* frame_out) {
  if (row >= p.rows) return false;
  const int s = ctc_segment_of(p.seg_start, p.n_seg, row);
  if (s < 0) return false;
  const long long f = row - __ldg(p.seg_start + s);
  *seg_out = s; *frame_out = f;
  if (f >= (long long)__ldg(p.seg_len + s)) return false;
  const long long tok = __ldg(p.tokens + row);
  if (tok == p.blank) return false;
  if (p.mode == 0 && f > 0 && __ldg(p.tokens + row - 1) == tok) return false;
  return true;
}

__global__ void __launch_bounds__(CTC_COMPACT_THREADS) ctc_compact_count_kernel(CtcCompactParams p) {
  __shared__ int s_warp[CTC_COMPACT_THREADS / 32];
  const long long base = (long long)blockIdx.x * CTC_COMPACT_TILE + (long long)threadIdx.x * CTC_COMPACT_ITEMS;
  int n = 0, seg; long long fr;
#pragma unroll
  for (int i = 0; i < CTC_COMPACT_ITEMS; ++i) n += ctc_keep(p, base + i, &seg, &fr) ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = n;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int w = 0; w < CTC_COMPACT_THREADS / 32; ++w) t += s_warp[w];
    p.block_counts[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(CTC_COMPACT_THREADS) ctc_compact_scatter_kernel(CtcCompactParams p) {
  __shared__ long long s_red[CTC_COMPACT_THREADS / 32];
  __shared__ int s_warp[CTC_COMPACT_THREADS / 32];
  __shared__ long long s_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // exclusive prefix of this block = sum of the counts of all earlier blocks
  long long acc = 0;
  for (int b = threadIdx.x; b < int(blockIdx.x); b += CTC_COMPACT_THREADS) acc += p.block_counts[b];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) s_red[warp] = acc;
  const long long base = (long long)blockIdx.x * CTC_COMPACT_TILE + (long long)threadIdx.x * CTC_COMPACT_ITEMS;
  bool keep[CTC_COMPACT_ITEMS]; int seg[CTC_COMPACT_ITEMS]; long long fr[CTC_COMPACT_ITEMS];
  int n = 0;
#pragma unroll
  for (int i = 0; i < CTC_COMPACT_ITEMS; ++i) {
    seg[i] = -1; fr[i] = -1;
    keep[i] = ctc_keep(p, base + i, &seg[i], &fr[i]);
    n += keep[i] ? 1 : 0;
  }
  int incl = n;                                   // inclusive scan of the per-thread counts inside the warp
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t = 0;
#pragma unroll
    for (int w = 0; w < CTC_COMPACT_THREADS / 32; ++w) t += s_red[w];
    s_base = t;
  }
  __syncthreads();
  long long pos = s_base + (incl - n);
  for (int w = 0; w < warp; ++w) pos += s_warp[w];
#pragma unroll
  for (int i = 0; i < CTC_COMPACT_ITEMS; ++i) {
    const long long row = base + i;
    // the row an utterance starts on publishes that utterance's output offset (and that of empty utterances sharing the row)
    if (row < p.rows && fr[i] == 0) {
      for (int s = seg[i]; s >= 0 && __ldg(p.seg_start + s) == row; --s) p.out_offsets[s] = pos;
    }
    if (keep[i]) {
      p.out_tokens[pos] = __ldg(p.tokens + row);
      p.out_frames[pos] = int(fr[i]);
      ++pos;
    }
  }
  // the thread that owns the last row closes the table: total, and utterances that start at or beyond the end of the buffer
  if (base <= p.rows - 1 && p.rows - 1 < base + CTC_COMPACT_ITEMS) {
    p.out_offsets[p.n_seg] = pos;
    for (int s = p.n_seg - 1; s >= 0 && __ldg(p.seg_start + s) >= p.rows; --s) p.out_offsets[s] = pos;
  }
}

// fp32 -> bf16 (round to nearest even), 4 elements per thread and iteration, grid-stride.  Used by cf_ctc_greedy when the caller
// hands over the fp32 encoder output (the reference's tensor type) instead of cf_encode's bf16 twin.
__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(in)[i];
    uint2 o;
    o.x = pack_bf16(v.x, v.y);
    o.y = pack_bf16(v.z, v.w);
    reinterpret_cast<uint2*>(out)[i] = o;
  }
}

// 16-byte-granular copy / zero fill by SMs.  cf_encode uses them instead of cudaMemcpyAsync / cudaMemsetAsync on the compute
// stream: a copy-engine operation queues behind every host-to-device feature copy issued earlier (one DMA FIFO per direction),
// which serialised the whole 461 MB upload in front of the encoder (measured: +6 ms per end-to-end step).  `src` may be mapped
// pinned host memory (the plan tables): the SMs read it over PCIe directly.
__global__ void copy16_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long long n16) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) dst[i] = src[i];
}
__global__ void zero16_kernel(uint4* __restrict__ dst, long long n16) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x)
    dst[i] = make_uint4(0u, 0u, 0u, 0u);
}

}  // namespace cf
