// Transducer greedy search on the device: LSTM predictor + joint network + per-utterance greedy control,
// replacing the Python loop of transducer/search/greedy_search.py:6-80 (one predictor step, one joint call and a dozen
// indexing kernels per frame and symbol) by five small kernels per *iteration*, all utterances of the batch in lock-step.
//
// What makes the search cheap: the predictor output only changes when a non-blank symbol is emitted, so with the current
// predictor vector g_b fixed the joint can be evaluated speculatively for the next RNNT_FB frames of utterance b at once
// (greedy_search.py evaluates them one at a time with the same vector); the decide kernel walks those frames in order,
// stops at the first non-blank symbol, and only then is a predictor step needed.  An iteration is
//   predictor step for the utterances that just emitted (LSTM layers, fused projection o pred_ffn)
//   -> joint: tanh(enc_ffn(enc)[t_b .. t_b + FB) + g_b) . ffn_out^T, per-tile argmax
//   -> decide: first non-blank of the block, emit, flip the state double-buffer, advance (t_b, step_b).
// Everything is fp32 on CUDA cores: the search is a chain of argmax decisions, so a reduced-precision joint would change
// the hypothesis after its first near-tie; the work per iteration is a few MFLOP and latency-bound anyway.
#pragma once
#include "common.cuh"

namespace cf {

constexpr int RNNT_FB = 8;         // frames evaluated speculatively per iteration and utterance
constexpr int RNNT_BT = 8;         // utterances per register tile in the predictor kernels
constexpr int RNNT_JV = 128;       // vocabulary entries per joint CTA (one per thread)

struct RnntState {                 // device arrays, one entry per utterance unless noted
  int* t;                          // next frame to evaluate
  int* step;                       // 1-based symbol slot on that frame
  int* token;                      // predictor input (last emitted symbol, blank at start)
  int* cur;                        // which half of the state double-buffer is committed
  int* count;                      // symbols emitted so far
  int* active;                     // [B] utterances that need a predictor step this iteration
  int* n_active;                   // [1]
  int* remaining;                  // [1] utterances not finished
  int* overflow;                   // [1] set when an utterance ran out of output capacity
  float* h;                        // [2][layers][B][H]
  float* c;                        // [2][layers][B][H]
  float* g;                        // [B][J]   pred_ffn(projection(h_top)) of the candidate state
  float* part_val;                 // [B][FB][n_vtiles]
  int* part_idx;                   // [B][FB][n_vtiles]
};

// ---------------------------------------------------------------------------------------------
// C[M, N] = A[M, K] . W[N, K]^T + bias, fp32 (enc_ffn over all encoder rows, once per call)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rnnt_linear_f32_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                              const float* __restrict__ bias, float* __restrict__ C,
                                                              long long M, int N, int K) {
  __shared__ float sA[16][64 + 1];
  __shared__ float sW[16][64 + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const long long m0 = (long long)blockIdx.y * 64;
  const int n0 = blockIdx.x * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int r = i >> 4, kk = i & 15;
      sA[kk][r] = (m0 + r < M && k0 + kk < K) ? A[(m0 + r) * K + k0 + kk] : 0.f;
      sW[kk][r] = (n0 + r < N && k0 + kk < K) ? W[(long long)(n0 + r) * K + k0 + kk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sA[kk][ty * 4 + i]; w[i] = sW[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long m = m0 + ty * 4 + i;
      const int n = n0 + tx * 4 + j;
      if (m < M && n < N) C[m * N + n] = acc[i][j] + bias[n];
    }
}

// ---------------------------------------------------------------------------------------------
// One LSTM layer step for the active utterances (torch.nn.LSTM cell, gate order i, f, g, o; predictor.py:190-203).
// Warp = one hidden unit for every active utterance: its four gate rows are read once per tile of RNNT_BT utterances.
// x: layer 0 -> embedding row of the utterance's token; layer > 0 -> candidate h of the layer below.
// Reads the committed half of the state double-buffer, writes the candidate half.
// ---------------------------------------------------------------------------------------------
struct RnntLstmParams {
  const float* w_ih; const float* w_hh; const float* b_ih; const float* b_hh;   // [4H, In], [4H, H], [4H], [4H]
  const float* embed;       // layer 0: [V, In]; else null
  int layer, In, H, B, layers;
};

CF_DEVINL float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(128) rnnt_lstm_kernel(RnntLstmParams p, RnntState s) {
  const int n_act = *s.n_active;
  if (n_act == 0) return;
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (j >= p.H) return;
  const size_t half = size_t(p.layers) * p.B * p.H;
  for (int b0 = 0; b0 < n_act; b0 += RNNT_BT) {
    float acc[RNNT_BT][4];
#pragma unroll
    for (int u = 0; u < RNNT_BT; ++u)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[u][q] = 0.f;
    const float* xin[RNNT_BT]; const float* hin[RNNT_BT];
#pragma unroll
    for (int u = 0; u < RNNT_BT; ++u) {
      const int b = (b0 + u < n_act) ? s.active[b0 + u] : s.active[b0];
      const int cur = s.cur[b];
      hin[u] = s.h + cur * half + (size_t(p.layer) * p.B + b) * p.H;
      xin[u] = p.embed ? p.embed + size_t(s.token[b]) * p.In
                       : s.h + (cur ^ 1) * half + (size_t(p.layer - 1) * p.B + b) * p.H;
    }
    for (int k = lane * 4; k < p.In; k += 128) {
      float4 w[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) w[q] = __ldg(reinterpret_cast<const float4*>(p.w_ih + size_t(q * p.H + j) * p.In + k));
#pragma unroll
      for (int u = 0; u < RNNT_BT; ++u) {
        const float4 x = *reinterpret_cast<const float4*>(xin[u] + k);
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[u][q] += w[q].x * x.x + w[q].y * x.y + w[q].z * x.z + w[q].w * x.w;
      }
    }
    for (int k = lane * 4; k < p.H; k += 128) {
      float4 w[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) w[q] = __ldg(reinterpret_cast<const float4*>(p.w_hh + size_t(q * p.H + j) * p.H + k));
#pragma unroll
      for (int u = 0; u < RNNT_BT; ++u) {
        const float4 x = *reinterpret_cast<const float4*>(hin[u] + k);
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[u][q] += w[q].x * x.x + w[q].y * x.y + w[q].z * x.z + w[q].w * x.w;
      }
    }
#pragma unroll
    for (int u = 0; u < RNNT_BT; ++u)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[u][q] = warp_sum(acc[u][q]);
    // lane u finalises utterance u of the tile
#pragma unroll
    for (int u = 0; u < RNNT_BT; ++u) {
      if (lane == u && b0 + u < n_act) {
        const int b = s.active[b0 + u];
        const int cur = s.cur[b];
        float gate[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) gate[q] = acc[u][q] + p.b_ih[q * p.H + j] + p.b_hh[q * p.H + j];
        const size_t at = (size_t(p.layer) * p.B + b) * p.H + j;
        const float c_old = s.c[cur * half + at];
        const float c_new = sigmoidf_acc(gate[1]) * c_old + sigmoidf_acc(gate[0]) * tanhf(gate[2]);
        s.c[(cur ^ 1) * half + at] = c_new;
        s.h[(cur ^ 1) * half + at] = sigmoidf_acc(gate[3]) * tanhf(c_new);
      }
    }
  }
}

// g_b = Wc . h_top'(b) + bc  with  Wc = pred_ffn.W . projection.W,  bc = pred_ffn.W . projection.b + pred_ffn.b
// (predictor.py:205 followed by joint.py:88; composed once at load time in fp64).  Warp = one output row.
__global__ void __launch_bounds__(128) rnnt_predproj_kernel(const float* __restrict__ Wc, const float* __restrict__ bc, int J,
                                                            int H, int B, int layers, RnntState s) {
  const int n_act = *s.n_active;
  if (n_act == 0) return;
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (r >= J) return;
  const size_t half = size_t(layers) * B * H;
  for (int b0 = 0; b0 < n_act; b0 += RNNT_BT) {
    float acc[RNNT_BT];
    const float* hin[RNNT_BT];
#pragma unroll
    for (int u = 0; u < RNNT_BT; ++u) {
      acc[u] = 0.f;
      const int b = (b0 + u < n_act) ? s.active[b0 + u] : s.active[b0];
      hin[u] = s.h + (s.cur[b] ^ 1) * half + (size_t(layers - 1) * B + b) * H;
    }
    for (int k = lane * 4; k < H; k += 128) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(Wc + size_t(r) * H + k));
#pragma unroll
      for (int u = 0; u < RNNT_BT; ++u) {
        const float4 x = *reinterpret_cast<const float4*>(hin[u] + k);
        acc[u] += w.x * x.x + w.y * x.y + w.z * x.z + w.w * x.w;
      }
    }
#pragma unroll
    for (int u = 0; u < RNNT_BT; ++u) acc[u] = warp_sum(acc[u]);
#pragma unroll
    for (int u = 0; u < RNNT_BT; ++u)
      if (lane == u && b0 + u < n_act) s.g[size_t(s.active[b0 + u]) * J + r] = acc[u] + bc[r];
  }
}

// ---------------------------------------------------------------------------------------------
// Joint over the next RNNT_FB frames of every unfinished utterance with its current predictor vector
// (joint.py:94-101: tanh(enc_ffn(enc) + pred_ffn(pred)) -> ffn_out), per-tile argmax (lowest index on ties, torch.argmax).
// grid (ceil(V / 128), B); thread = one vocabulary entry, ffn_out stored transposed [J, V] for coalesced reads.
// ---------------------------------------------------------------------------------------------
struct RnntJointParams {
  const float* E;            // [rows, J] enc_ffn output
  const float* WoT;          // [J, V]
  const float* bo;           // [V]
  const long long* seg_start; const int* seg_len;   // device [B]
  int J, V, n_vtiles;
};

template <int JMAX>
__global__ void __launch_bounds__(RNNT_JV) rnnt_joint_kernel(RnntJointParams p, RnntState s) {
  __shared__ __align__(16) float a[RNNT_FB][JMAX];
  __shared__ float s_val[RNNT_FB][RNNT_JV / 32];
  __shared__ int s_idx[RNNT_FB][RNNT_JV / 32];
  const int b = blockIdx.y;
  const int t0 = s.t[b], len = p.seg_len[b];
  if (t0 >= len) return;
  const int nf = min(RNNT_FB, len - t0);
  const float* g = s.g + size_t(b) * p.J;
  const float* e = p.E + (p.seg_start[b] + t0) * (long long)p.J;
  for (int i = threadIdx.x; i < RNNT_FB * p.J; i += RNNT_JV) {
    const int f = i / p.J, k = i - f * p.J;
    a[f][k] = f < nf ? tanhf(e[(long long)f * p.J + k] + g[k]) : 0.f;
  }
  __syncthreads();
  const int v = blockIdx.x * RNNT_JV + threadIdx.x;
  const int vc = min(v, p.V - 1);
  float acc[RNNT_FB];
#pragma unroll
  for (int f = 0; f < RNNT_FB; ++f) acc[f] = 0.f;
  for (int k = 0; k < p.J; k += 4) {
    const float w0 = __ldg(p.WoT + size_t(k) * p.V + vc), w1 = __ldg(p.WoT + size_t(k + 1) * p.V + vc);
    const float w2 = __ldg(p.WoT + size_t(k + 2) * p.V + vc), w3 = __ldg(p.WoT + size_t(k + 3) * p.V + vc);
#pragma unroll
    for (int f = 0; f < RNNT_FB; ++f) {
      const float4 x = *reinterpret_cast<const float4*>(&a[f][k]);
      acc[f] = fmaf(x.x, w0, acc[f]); acc[f] = fmaf(x.y, w1, acc[f]);
      acc[f] = fmaf(x.z, w2, acc[f]); acc[f] = fmaf(x.w, w3, acc[f]);
    }
  }
  const float bias = __ldg(p.bo + vc);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int f = 0; f < RNNT_FB; ++f) {
    float val = v < p.V ? acc[f] + bias : -INFINITY;
    int idx = v;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, val, o);
      const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ov > val || (ov == val && oi < idx)) { val = ov; idx = oi; }
    }
    if (lane == 0) { s_val[f][warp] = val; s_idx[f][warp] = idx; }
  }
  __syncthreads();
  if (threadIdx.x < RNNT_FB) {
    const int f = threadIdx.x;
    float val = s_val[f][0]; int idx = s_idx[f][0];
#pragma unroll
    for (int w = 1; w < RNNT_JV / 32; ++w)
      if (s_val[f][w] > val || (s_val[f][w] == val && s_idx[f][w] < idx)) { val = s_val[f][w]; idx = s_idx[f][w]; }
    const size_t at = (size_t(b) * RNNT_FB + f) * p.n_vtiles + blockIdx.x;
    s.part_val[at] = val;
    s.part_idx[at] = idx;
  }
}

// ---------------------------------------------------------------------------------------------
// Greedy control (greedy_search.py:24-76), one thread per utterance: walk the evaluated frames in order; blank -> next
// frame, slot 1; non-blank -> record (token, frame), make the candidate predictor state the committed one, stay on the
// frame unless its n_steps slots are used up, and queue the utterance for a predictor step.
// ---------------------------------------------------------------------------------------------
struct RnntDecideParams {
  const int* seg_len;
  long long* out_tokens; int* out_frames; int* out_counts;     // [B][cap], [B][cap], [B]
  int B, n_vtiles, n_steps, cap, blank;
};

__global__ void __launch_bounds__(256) rnnt_decide_kernel(RnntDecideParams p, RnntState s) {
  __shared__ int s_nact, s_rem;
  if (threadIdx.x == 0) { s_nact = 0; s_rem = 0; }
  __syncthreads();
  for (int b = threadIdx.x; b < p.B; b += blockDim.x) {
    int t = s.t[b];
    const int len = p.seg_len[b];
    if (t >= len) continue;
    int step = s.step[b];
    bool emitted = false;
    for (int f = 0; f < RNNT_FB && t < len; ++f) {
      const size_t at = (size_t(b) * RNNT_FB + f) * p.n_vtiles;
      float val = s.part_val[at]; int idx = s.part_idx[at];
      for (int w = 1; w < p.n_vtiles; ++w) {
        const float ov = s.part_val[at + w]; const int oi = s.part_idx[at + w];
        if (ov > val || (ov == val && oi < idx)) { val = ov; idx = oi; }
      }
      if (idx == p.blank) { ++t; step = 1; continue; }
      const int n = s.count[b];
      if (n >= p.cap) { *s.overflow = 1; t = len; break; }
      p.out_tokens[size_t(b) * p.cap + n] = idx;
      p.out_frames[size_t(b) * p.cap + n] = t;
      s.count[b] = n + 1;
      s.token[b] = idx;
      s.cur[b] ^= 1;
      if (++step > p.n_steps) { ++t; step = 1; }
      emitted = true;
      break;
    }
    s.t[b] = t; s.step[b] = step;
    if (t < len) {
      atomicAdd(&s_rem, 1);
      if (emitted) s.active[atomicAdd(&s_nact, 1)] = b;
    }
    if (t >= len) p.out_counts[b] = s.count[b];
  }
  __syncthreads();
  if (threadIdx.x == 0) { *s.n_active = s_nact; *s.remaining = s_rem; }
}

// start of a search: zero state, blank token, every non-empty utterance queued for its first predictor step
__global__ void rnnt_init_kernel(RnntState s, const int* seg_len, int* out_counts, int B, int blank, size_t state_floats) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  for (size_t k = i; k < state_floats; k += size_t(gridDim.x) * blockDim.x) { s.h[k] = 0.f; s.c[k] = 0.f; }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int n = 0;
    for (int b = 0; b < B; ++b) {
      s.t[b] = 0; s.step[b] = 1; s.token[b] = blank; s.cur[b] = 0; s.count[b] = 0; out_counts[b] = 0;
      if (seg_len[b] > 0) s.active[n++] = b;
    }
    *s.n_active = n; *s.remaining = n; *s.overflow = 0;
  }
}

}  // namespace cf
