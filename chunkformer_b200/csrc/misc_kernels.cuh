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

}  // namespace cf
