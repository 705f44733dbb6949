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
  // tiles of RNNT_BT utterances run side by side (blockIdx.y): the step is a chain of L2 round trips, not throughput
  for (int b0 = blockIdx.y * RNNT_BT; b0 < n_act; b0 += gridDim.y * RNNT_BT) {
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
    // both products in slabs of 512 columns: the slab's 16 weight vectors of this unit are all requested before the first
    // FMA (the kernel is latency-bound: a few hundred FMAs per warp behind L2 round trips)
    auto slab = [&](const float* wmat, int ld, const float* const* vec, int k0) {
      float4 w[4][4];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int k = k0 + m * 128 + lane * 4;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          w[q][m] = k < ld ? __ldg(reinterpret_cast<const float4*>(wmat + size_t(q * p.H + j) * ld + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < RNNT_BT; ++u) {
        float4 x[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int k = k0 + m * 128 + lane * 4;
          x[m] = k < ld ? *reinterpret_cast<const float4*>(vec[u] + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int m = 0; m < 4; ++m)
#pragma unroll
          for (int q = 0; q < 4; ++q)
            acc[u][q] += w[q][m].x * x[m].x + w[q][m].y * x[m].y + w[q][m].z * x[m].z + w[q][m].w * x[m].w;
      }
    };
    for (int k0 = 0; k0 < p.In; k0 += 512) slab(p.w_ih, p.In, xin, k0);
    for (int k0 = 0; k0 < p.H; k0 += 512) slab(p.w_hh, p.H, hin, k0);
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
  for (int b0 = blockIdx.y * RNNT_BT; b0 < n_act; b0 += gridDim.y * RNNT_BT) {
    float acc[RNNT_BT];
    const float* hin[RNNT_BT];
#pragma unroll
    for (int u = 0; u < RNNT_BT; ++u) {
      acc[u] = 0.f;
      const int b = (b0 + u < n_act) ? s.active[b0 + u] : s.active[b0];
      hin[u] = s.h + (s.cur[b] ^ 1) * half + (size_t(layers - 1) * B + b) * H;
    }
    for (int k0 = 0; k0 < H; k0 += 512) {
      float4 w[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int k = k0 + m * 128 + lane * 4;
        w[m] = k < H ? __ldg(reinterpret_cast<const float4*>(Wc + size_t(r) * H + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < RNNT_BT; ++u) {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int k = k0 + m * 128 + lane * 4;
          const float4 x = k < H ? *reinterpret_cast<const float4*>(hin[u] + k) : make_float4(0.f, 0.f, 0.f, 0.f);
          acc[u] += w[m].x * x.x + w[m].y * x.y + w[m].z * x.z + w[m].w * x.w;
        }
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

constexpr int RNNT_JK = 4;        // K slices per vocabulary entry (threads per CTA = RNNT_JV * RNNT_JK)
constexpr int RNNT_JU = 1;        // utterances per CTA (more share one read of the ffn_out tile, but leave SMs idle: measured slower)
constexpr int RNNT_JR = RNNT_JU * RNNT_FB;   // activation rows per CTA

inline size_t rnnt_joint_smem_bytes(int J) {
  return size_t(RNNT_JR) * J * 4 + size_t(RNNT_JK - 1) * RNNT_JR * RNNT_JV * 4 + size_t(RNNT_JR) * (RNNT_JV / 32) * 8;
}

__global__ void __launch_bounds__(RNNT_JV * RNNT_JK) rnnt_joint_kernel(RnntJointParams p, RnntState s, int B) {
  extern __shared__ __align__(16) float jsm[];
  float* a = jsm;                                          // [JR][J]
  float* s_acc = a + size_t(RNNT_JR) * p.J;                // [JK - 1][JR][JV]
  float* s_val = s_acc + size_t(RNNT_JK - 1) * RNNT_JR * RNNT_JV;   // [JR][JV / 32]
  int* s_idx = reinterpret_cast<int*>(s_val + RNNT_JR * (RNNT_JV / 32));
  __shared__ int s_t0[RNNT_JU], s_nf[RNNT_JU];
  const int b_first = blockIdx.y * RNNT_JU;
  if (threadIdx.x < RNNT_JU) {
    const int b = b_first + threadIdx.x;
    int t0 = 0, nf = 0;
    if (b < B) { t0 = s.t[b]; nf = max(0, min(RNNT_FB, p.seg_len[b] - t0)); }
    s_t0[threadIdx.x] = t0; s_nf[threadIdx.x] = nf;
  }
  __syncthreads();
  {
    int any = 0;
#pragma unroll
    for (int u = 0; u < RNNT_JU; ++u) any += s_nf[u];
    if (any == 0) return;
  }
  for (int i = threadIdx.x; i < RNNT_JR * p.J; i += RNNT_JV * RNNT_JK) {
    const int r = i / p.J, k = i - r * p.J;
    const int u = r / RNNT_FB, f = r - u * RNNT_FB;
    float val = 0.f;
    if (f < s_nf[u]) {
      const int b = b_first + u;
      val = tanhf(p.E[(p.seg_start[b] + s_t0[u] + f) * (long long)p.J + k] + s.g[size_t(b) * p.J + k]);
    }
    a[i] = val;
  }
  __syncthreads();
  const int tv = threadIdx.x & (RNNT_JV - 1), kq = threadIdx.x / RNNT_JV;   // a warp shares kq: the tile reads broadcast
  const int v = blockIdx.x * RNNT_JV + tv;
  const int vc = min(v, p.V - 1);
  float acc[RNNT_JR];
#pragma unroll
  for (int r = 0; r < RNNT_JR; ++r) acc[r] = 0.f;
  // slice kq takes the 4-wide k groups kq, kq + JK, ...; the 64 weights of a 256-column phase are all requested before the
  // first FMA (two L2 round trips for J = 512 instead of one per step)
  const float* wp = p.WoT + vc;
  constexpr int GROUPS = 16;
  for (int k0 = 0; k0 < p.J; k0 += GROUPS * 4 * RNNT_JK) {
    float w[GROUPS][4];
#pragma unroll
    for (int m = 0; m < GROUPS; ++m) {
      const int km = k0 + 4 * (kq + RNNT_JK * m);
#pragma unroll
      for (int i = 0; i < 4; ++i) w[m][i] = km < p.J ? __ldg(wp + size_t(km + i) * p.V) : 0.f;
    }
#pragma unroll
    for (int m = 0; m < GROUPS; ++m) {
      const int km = k0 + 4 * (kq + RNNT_JK * m);
      if (km < p.J) {
#pragma unroll
        for (int r = 0; r < RNNT_JR; ++r) {
          const float4 x = *reinterpret_cast<const float4*>(&a[size_t(r) * p.J + km]);
          acc[r] = fmaf(x.x, w[m][0], acc[r]); acc[r] = fmaf(x.y, w[m][1], acc[r]);
          acc[r] = fmaf(x.z, w[m][2], acc[r]); acc[r] = fmaf(x.w, w[m][3], acc[r]);
        }
      }
    }
  }
  if (kq > 0) {
#pragma unroll
    for (int r = 0; r < RNNT_JR; ++r) s_acc[(size_t(kq - 1) * RNNT_JR + r) * RNNT_JV + tv] = acc[r];
  }
  __syncthreads();
  if (kq == 0) {
    const float bias = __ldg(p.bo + vc);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int r = 0; r < RNNT_JR; ++r) {
      float sum = acc[r];
#pragma unroll
      for (int q = 0; q < RNNT_JK - 1; ++q) sum += s_acc[(size_t(q) * RNNT_JR + r) * RNNT_JV + tv];
      float val = v < p.V ? sum + bias : -INFINITY;
      int idx = v;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, val, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov > val || (ov == val && oi < idx)) { val = ov; idx = oi; }
      }
      if (lane == 0) { s_val[r * (RNNT_JV / 32) + warp] = val; s_idx[r * (RNNT_JV / 32) + warp] = idx; }
    }
  }
  __syncthreads();
  if (threadIdx.x < RNNT_JR) {
    const int r = threadIdx.x, u = r / RNNT_FB, f = r - u * RNNT_FB;
    if (f < s_nf[u]) {
      float val = s_val[r * (RNNT_JV / 32)]; int idx = s_idx[r * (RNNT_JV / 32)];
#pragma unroll
      for (int w = 1; w < RNNT_JV / 32; ++w) {
        const float ov = s_val[r * (RNNT_JV / 32) + w]; const int oi = s_idx[r * (RNNT_JV / 32) + w];
        if (ov > val || (ov == val && oi < idx)) { val = ov; idx = oi; }
      }
      const size_t at = (size_t(b_first + u) * RNNT_FB + f) * p.n_vtiles + blockIdx.x;
      s.part_val[at] = val;
      s.part_idx[at] = idx;
    }
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
  const int lane = threadIdx.x & 31;
  for (int b = threadIdx.x >> 5; b < p.B; b += blockDim.x >> 5) {       // warp per utterance
    int t = s.t[b];
    const int len = p.seg_len[b];
    if (t >= len) continue;
    // lane = 4 f + q: the four lanes of frame f split its vocabulary tiles, then combine (lowest index wins ties)
    static_assert(RNNT_FB == 8, "lane layout assumes 8 frames x 4 lanes");
    const int f_l = lane >> 2, q_l = lane & 3;
    const size_t at = (size_t(b) * RNNT_FB + f_l) * p.n_vtiles;
    float val = -INFINITY; int idx = 0x7fffffff;
    for (int w = q_l; w < p.n_vtiles; w += 4) {
      const float ov = s.part_val[at + w]; const int oi = s.part_idx[at + w];
      if (ov > val || (ov == val && oi < idx)) { val = ov; idx = oi; }
    }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, val, o);
      const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ov > val || (ov == val && oi < idx)) { val = ov; idx = oi; }
    }
    // first frame of the block (still inside the utterance) whose symbol is not blank
    const unsigned nb = __ballot_sync(0xffffffffu, q_l == 0 && idx != p.blank && t + f_l < len);
    const int limit = min(RNNT_FB, len - t);
    const int f_hit = nb ? (__ffs(nb) - 1) >> 2 : limit;
    const int sym = __shfl_sync(0xffffffffu, idx, (f_hit < RNNT_FB ? f_hit : 0) * 4);
    if (lane == 0) {
      int step = s.step[b];
      bool emitted = false;
      if (f_hit > 0) step = 1;                  // blanks before the hit: next frame, slot 1
      t += f_hit;
      if (f_hit < limit) {
        const int n = s.count[b];
        if (n >= p.cap) { *s.overflow = 1; t = len; }
        else {
          p.out_tokens[size_t(b) * p.cap + n] = sym;
          p.out_frames[size_t(b) * p.cap + n] = t;
          s.count[b] = n + 1;
          s.token[b] = sym;
          s.cur[b] ^= 1;
          if (++step > p.n_steps) { ++t; step = 1; }
          emitted = true;
        }
      }
      s.t[b] = t; s.step[b] = step;
      if (t < len) {
        atomicAdd(&s_rem, 1);
        if (emitted) s.active[atomicAdd(&s_nact, 1)] = b;
      } else {
        p.out_counts[b] = s.count[b];
      }
    }
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
