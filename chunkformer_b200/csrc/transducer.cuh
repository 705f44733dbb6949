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
#include <algorithm>

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
  int* active;                     // [B] utterances that need a predictor step this iteration (compact, -1 = empty slot)
  int* act_cur;                    // [B] committed state half of active[i]   } read in ONE round trip by the predictor
  int* act_tok;                    // [B] predictor input token of active[i]  } kernels (no n_active -> active -> cur chain)
  int* need_g;                     // [B] 1 when the utterance emitted a symbol this iteration (its predictor vector is stale)
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
  // compact list of the utterances that need a predictor step: (utterance, committed state half, input token); entries beyond
  // *cnt are stale (cnt == null: the list ends at the first -1)
  const int* act; const int* act_cur; const int* act_tok; const int* cnt;
};

CF_DEVINL float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(128) rnnt_lstm_kernel(RnntLstmParams p, RnntState s) {
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (j >= p.H) return;
  const size_t half = size_t(p.layers) * p.B * p.H;
  // gate biases of this unit: no dependence on the utterances, requested first
  float bias[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) bias[q] = __ldg(p.b_ih + q * p.H + j) + __ldg(p.b_hh + q * p.H + j);
  // tiles of RNNT_BT utterances run side by side (blockIdx.y): the step is a chain of L2 round trips, not throughput
  for (int b0 = blockIdx.y * RNNT_BT; b0 < p.B; b0 += gridDim.y * RNNT_BT) {
    int bs[RNNT_BT], curs[RNNT_BT], toks[RNNT_BT];
    const int n_list = p.cnt ? *p.cnt : p.B;              // read in the same round trip as the entries
#pragma unroll
    for (int u = 0; u < RNNT_BT; ++u) {
      const bool in = b0 + u < p.B;
      bs[u] = in ? p.act[b0 + u] : -1;
      curs[u] = in ? p.act_cur[b0 + u] : 0;
      toks[u] = in ? p.act_tok[b0 + u] : 0;
    }
#pragma unroll
    for (int u = 0; u < RNNT_BT; ++u)
      if (b0 + u >= n_list) bs[u] = -1;
    if (bs[0] < 0) return;                      // the list is compact: an empty first slot ends the work
    float acc[RNNT_BT][4];
#pragma unroll
    for (int u = 0; u < RNNT_BT; ++u)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[u][q] = 0.f;
    const float* xin[RNNT_BT]; const float* hin[RNNT_BT];
    float c_old[RNNT_BT];
#pragma unroll
    for (int u = 0; u < RNNT_BT; ++u) {
      const int b = bs[u] >= 0 ? bs[u] : bs[0];
      const int cur = bs[u] >= 0 ? curs[u] : curs[0];
      const int tok = bs[u] >= 0 ? toks[u] : toks[0];
      hin[u] = s.h + cur * half + (size_t(p.layer) * p.B + b) * p.H;
      xin[u] = p.embed ? p.embed + size_t(tok) * p.In
                       : s.h + (cur ^ 1) * half + (size_t(p.layer - 1) * p.B + b) * p.H;
      c_old[u] = s.c[cur * half + (size_t(p.layer) * p.B + b) * p.H + j];
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
      if (lane == u && bs[u] >= 0) {
        const int b = bs[u], cur = curs[u];
        float gate[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) gate[q] = acc[u][q] + bias[q];
        const size_t at = (size_t(p.layer) * p.B + b) * p.H + j;
        const float c_new = sigmoidf_acc(gate[1]) * c_old[u] + sigmoidf_acc(gate[0]) * tanhf(gate[2]);
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
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (r >= J) return;
  const size_t half = size_t(layers) * B * H;
  const float bias = __ldg(bc + r);
  for (int b0 = blockIdx.y * RNNT_BT; b0 < B; b0 += gridDim.y * RNNT_BT) {
    int bs[RNNT_BT], curs[RNNT_BT];
#pragma unroll
    for (int u = 0; u < RNNT_BT; ++u) {
      const bool in = b0 + u < B;
      bs[u] = in ? s.active[b0 + u] : -1;
      curs[u] = in ? s.act_cur[b0 + u] : 0;
    }
    if (bs[0] < 0) return;
    float acc[RNNT_BT];
    const float* hin[RNNT_BT];
#pragma unroll
    for (int u = 0; u < RNNT_BT; ++u) {
      acc[u] = 0.f;
      const int b = bs[u] >= 0 ? bs[u] : bs[0];
      const int cur = bs[u] >= 0 ? curs[u] : curs[0];
      hin[u] = s.h + (cur ^ 1) * half + (size_t(layers - 1) * B + b) * H;
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
      if (lane == u && bs[u] >= 0) s.g[size_t(bs[u]) * J + r] = acc[u] + bias;
  }
}

// ---------------------------------------------------------------------------------------------
// Joint over the next RNNT_FB frames of every unfinished utterance with its current predictor vector
// (joint.py:94-101: tanh(enc_ffn(enc) + pred_ffn(pred)) -> ffn_out), per-tile argmax (lowest index on ties, torch.argmax).
// grid (ceil(V / 128), B); thread = one vocabulary entry, ffn_out stored transposed [J, V] for coalesced reads.
// ---------------------------------------------------------------------------------------------
struct RnntJointParams {
  const float* E;            // [rows, J] enc_ffn output
  const float* WoT;          // [J / 4][V][4]: ffn_out transposed in groups of four k, one 16-byte load per (k group, v)
  const float* bo;           // [V]
  const long long* seg_start; const int* seg_len;   // device [B]
  int J, V, n_vtiles;
#ifdef CF_ABLATION
  int debug;                 // ablation build only (CF_RNNT_DEBUG): 1 = no tanh / E loads, 2 = no K loop
#define CF_RNNT_DBG(p, bit) ((p).debug & (bit))
#else
#define CF_RNNT_DBG(p, bit) 0
#endif
  // fused projection (cluster launch): g_b = Wc . h_top'(b) + bc computed by the cluster of the utterance's joint CTAs
  const float* Wc; const float* bc; int H, layers;
  // in-cluster greedy control (FUSED): the decide step of the utterance runs in CTA 0 of its cluster
  int* list_b; int* list_cur; int* list_tok; int* cnt; int* rem;        // this iteration's output list (appended atomically)
  int* cnt_next; int* rem_next;                                         // zeroed here for the next iteration
  long long* out_tokens; int* out_frames; int* out_counts; int n_steps, cap, blank;
};

constexpr int RNNT_VPT = 4;       // vocabulary entries per thread: every activation read from shared memory feeds 16 FMAs
constexpr int RNNT_JK = 16;       // K slices (warps) per CTA; a warp covers 32 x VPT = RNNT_JV vocabulary entries
constexpr int RNNT_JR = RNNT_FB;  // activation rows per CTA (one utterance; sharing the ffn_out tile between utterances leaves
                                  // SMs idle and was slower)
constexpr int RNNT_JTHREADS = 32 * RNNT_JK;
static_assert(RNNT_JV == 32 * RNNT_VPT, "a warp covers the CTA's vocabulary tile");

inline size_t rnnt_joint_smem_bytes(int J) {
  return size_t(RNNT_JR) * J * 4 + size_t(RNNT_JK) * RNNT_JR * RNNT_JV * 4 + size_t(J) * 4 + 8 * RNNT_FB * 8;
}

// Phase timings of the first version (thread = one vocabulary entry x one of 4 K-slices, 4-byte weight loads): 20 us =
// 5 launch / prologue / reductions + 3.7 activation tile (encoder rows, tanh) + 11.4 K loop; the K loop was bound by the
// shared-memory pipe (one 16-byte broadcast read per 4 FMAs), not by the weight loads (16-byte loads alone: 18.6 us).
//
// FUSED (launched as one thread-block cluster per utterance = its vocabulary-tile CTAs, 2 / 4 / 8 of them): the projection
// g_b = Wc . h_top'(b) + bc of an utterance that just emitted is computed by the cluster itself, J / n_vtiles rows per CTA, and
// handed to every CTA of the cluster through distributed shared memory (one cluster barrier) instead of a separate launch
// (a dependent launch costs about 4.5 us on this path whatever it computes).
template <bool FUSED>
__global__ void __launch_bounds__(RNNT_JTHREADS) rnnt_joint_kernel(RnntJointParams p, RnntState s, int B) {
  extern __shared__ __align__(16) float jsm[];
  float* a = jsm;                                          // [JR][J]
  float* s_acc = a + size_t(RNNT_JR) * p.J;                // [JK][JR][JV]
  float* g_s = s_acc + size_t(RNNT_JK) * RNNT_JR * RNNT_JV;   // [J]
  float* pv_s = g_s + p.J;                                  // [8 ranks][FB] per-tile maxima   } filled through DSMEM in
  int* pi_s = reinterpret_cast<int*>(pv_s + 8 * RNNT_FB);   // [8 ranks][FB] their indices     } CTA 0 of the cluster
  const int b = blockIdx.y;
  if (FUSED && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) { *p.cnt_next = 0; *p.rem_next = 0; }
  const int t0 = s.t[b];
  const int nf = b < B ? max(0, min(RNNT_FB, p.seg_len[b] - t0)) : 0;
  if (FUSED) cluster_sync_all();                            // every CTA of the cluster runs before its shared memory is written
  if (nf == 0) return;                                      // the same for every CTA of the utterance's cluster
  const int lane = threadIdx.x & 31, kq = threadIdx.x >> 5;
  if (FUSED && s.need_g[b]) {
    const int nv = gridDim.x, rank = blockIdx.x;            // cluster = the grid's x extent
    const int rows_per = (p.J + nv - 1) / nv;
    const size_t half = size_t(p.layers) * B * p.H;
    const float* hin = s.h + (s.cur[b] ^ 1) * half + (size_t(p.layers - 1) * B + b) * p.H;
    const int r_end = min(p.J, (rank + 1) * rows_per);
    // four rows per warp at a time, all their weight vectors requested before the first FMA
    for (int r0 = rank * rows_per + kq; r0 < r_end; r0 += 4 * RNNT_JK) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int k0 = 0; k0 < p.H; k0 += 512) {
        float4 w[4][4], x[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int k = k0 + m * 128 + lane * 4;
          x[m] = k < p.H ? *reinterpret_cast<const float4*>(hin + k) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = r0 + i * RNNT_JK;
            w[i][m] = (k < p.H && r < r_end) ? __ldg(reinterpret_cast<const float4*>(p.Wc + size_t(r) * p.H + k))
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int m = 0; m < 4; ++m)
            acc[i] += w[i][m].x * x[m].x + w[i][m].y * x[m].y + w[i][m].z * x[m].z + w[i][m].w * x[m].w;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = r0 + i * RNNT_JK;
        const float val = warp_sum(acc[i]) + (r < r_end ? __ldg(p.bc + r) : 0.f);
        if (r < r_end) {
          if (lane < nv) {                                   // lane i hands the value to CTA i of the cluster
            const uint32_t remote = mapa_rank(smem_u32(g_s + r), uint32_t(lane));
            asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(val) : "memory");
          }
          if (lane == 0) s.g[size_t(b) * p.J + r] = val;     // kept for the iterations in which the utterance does not emit
        }
      }
    }
    cluster_sync_all();
  } else {
    for (int k = threadIdx.x; k < p.J; k += RNNT_JTHREADS) g_s[k] = s.g[size_t(b) * p.J + k];
    __syncthreads();
  }
  {
    const float* e = p.E + (p.seg_start[b] + t0) * (long long)p.J;
    for (int i = threadIdx.x; i < RNNT_JR * p.J; i += RNNT_JTHREADS) {
      const int f = i / p.J, k = i - f * p.J;
      a[i] = f < nf ? (CF_RNNT_DBG(p, 1) ? 0.5f : tanhf(e[(long long)f * p.J + k] + g_s[k])) : 0.f;
    }
  }
  __syncthreads();
  const int v0 = blockIdx.x * RNNT_JV + lane;             // this thread's entries: v0 + 32 i
  float acc[RNNT_JR][RNNT_VPT];
#pragma unroll
  for (int r = 0; r < RNNT_JR; ++r)
#pragma unroll
    for (int i = 0; i < RNNT_VPT; ++i) acc[r][i] = 0.f;
  // warp kq takes the 4-wide k groups kq, kq + JK, ...; four groups (16 weight vectors of 16 bytes) in flight per step
  const float4* wp = reinterpret_cast<const float4*>(p.WoT);
  for (int k0 = 0; k0 < (CF_RNNT_DBG(p, 2) ? 0 : p.J); k0 += 4 * 4 * RNNT_JK) {
    float4 w[4][RNNT_VPT];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int km = k0 + 4 * (kq + RNNT_JK * m);
#pragma unroll
      for (int i = 0; i < RNNT_VPT; ++i) {
        const int v = min(v0 + 32 * i, p.V - 1);
        w[m][i] = km < p.J ? __ldg(wp + size_t(km >> 2) * p.V + v) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int km = k0 + 4 * (kq + RNNT_JK * m);
      if (km < p.J) {
#pragma unroll
        for (int r = 0; r < RNNT_JR; ++r) {
          const float4 x = *reinterpret_cast<const float4*>(&a[size_t(r) * p.J + km]);
#pragma unroll
          for (int i = 0; i < RNNT_VPT; ++i) {
            acc[r][i] = fmaf(x.x, w[m][i].x, acc[r][i]); acc[r][i] = fmaf(x.y, w[m][i].y, acc[r][i]);
            acc[r][i] = fmaf(x.z, w[m][i].z, acc[r][i]); acc[r][i] = fmaf(x.w, w[m][i].w, acc[r][i]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < RNNT_JR; ++r)
#pragma unroll
    for (int i = 0; i < RNNT_VPT; ++i) s_acc[(size_t(kq) * RNNT_JR + r) * RNNT_JV + lane + 32 * i] = acc[r][i];
  __syncthreads();
  // warp r (r < JR) finishes frame r: sum the K slices for its 128 entries, add the bias, argmax (lowest index on ties)
  if (kq < RNNT_JR) {
    const int r = kq;
    float val = -INFINITY; int idx = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < RNNT_VPT; ++i) {
      const int v = v0 + 32 * i;
      float sum = 0.f;
#pragma unroll
      for (int q = 0; q < RNNT_JK; ++q) sum += s_acc[(size_t(q) * RNNT_JR + r) * RNNT_JV + lane + 32 * i];
      const float cand = v < p.V ? sum + __ldg(p.bo + v) : -INFINITY;
      if (cand > val || (cand == val && v < idx)) { val = cand; idx = v; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, val, o);
      const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ov > val || (ov == val && oi < idx)) { val = ov; idx = oi; }
    }
    if (!FUSED) {
      if (lane == 0 && r < nf) {
        const size_t at = (size_t(b) * RNNT_FB + r) * p.n_vtiles + blockIdx.x;
        s.part_val[at] = val;
        s.part_idx[at] = idx;
      }
    } else if (lane == 0) {                                  // frame r's maximum of this tile -> CTA 0 of the cluster
      const uint32_t rv = mapa_rank(smem_u32(pv_s + blockIdx.x * RNNT_FB + r), 0u);
      const uint32_t ri = mapa_rank(smem_u32(pi_s + blockIdx.x * RNNT_FB + r), 0u);
      asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(rv), "f"(val) : "memory");
      asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(ri), "r"(idx) : "memory");
    }
  }
  if (!FUSED) return;
  cluster_sync_all();                                        // CTA 0 was running all along (it took part in the K loop)
  if (blockIdx.x != 0 || kq != 0) return;
  // ---- greedy control of utterance b (same rules as rnnt_decide_kernel), one warp
  {
    const int nv = gridDim.x;
    int t = t0;
    const int len = p.seg_len[b];
    const int f_l = lane >> 2, q_l = lane & 3;
    float val = -INFINITY; int idx = 0x7fffffff;
    for (int w = q_l; w < nv; w += 4) {
      const float ov = pv_s[w * RNNT_FB + f_l]; const int oi = pi_s[w * RNNT_FB + f_l];
      if (ov > val || (ov == val && oi < idx)) { val = ov; idx = oi; }
    }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, val, o);
      const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ov > val || (ov == val && oi < idx)) { val = ov; idx = oi; }
    }
    const unsigned nb = __ballot_sync(0xffffffffu, q_l == 0 && idx != p.blank && t + f_l < len);
    const int limit = min(RNNT_FB, len - t);
    const int f_hit = nb ? (__ffs(nb) - 1) >> 2 : limit;
    const int sym = __shfl_sync(0xffffffffu, idx, (f_hit < RNNT_FB ? f_hit : 0) * 4);
    if (lane == 0) {
      int step = s.step[b];
      bool emitted = false;
      if (f_hit > 0) step = 1;
      t += f_hit;
      if (f_hit < limit) {
        const int n = s.count[b];
        if (n >= p.cap) { *s.overflow = 1; t = len; }
        else {
          p.out_tokens[size_t(b) * p.cap + n] = sym;
          p.out_frames[size_t(b) * p.cap + n] = t;
          s.count[b] = n + 1;
          s.token[b] = sym;
          const int cur = s.cur[b] ^ 1;
          s.cur[b] = cur;
          if (++step > p.n_steps) { ++t; step = 1; }
          emitted = true;
          if (t < len) {
            const int slot = atomicAdd(p.cnt, 1);
            p.list_b[slot] = b; p.list_cur[slot] = cur; p.list_tok[slot] = sym;
          }
        }
      }
      s.t[b] = t; s.step[b] = step;
      s.need_g[b] = emitted ? 1 : 0;
      if (t < len) atomicAdd(p.rem, 1);
      else p.out_counts[b] = s.count[b];
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
      s.need_g[b] = emitted ? 1 : 0;
      if (t < len) {
        atomicAdd(&s_rem, 1);
        if (emitted) {
          const int slot = atomicAdd(&s_nact, 1);
          s.active[slot] = b; s.act_cur[slot] = s.cur[b]; s.act_tok[slot] = sym;
        }
      } else {
        p.out_counts[b] = s.count[b];
      }
    }
  }
  __syncthreads();
  for (int i = s_nact + threadIdx.x; i < p.B; i += blockDim.x) s.active[i] = -1;     // empty slots of the compact list
  if (threadIdx.x == 0) { *s.n_active = s_nact; *s.remaining = s_rem; }
}

// start of a search: zero state, blank token, every non-empty utterance queued for its first predictor step
__global__ void rnnt_init_kernel(RnntState s, const int* seg_len, int* out_counts, int B, int blank, size_t state_floats,
                                 int* list_b, int* list_cur, int* list_tok, int* cnt2, int* rem2) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  for (size_t k = i; k < state_floats; k += size_t(gridDim.x) * blockDim.x) { s.h[k] = 0.f; s.c[k] = 0.f; }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int n = 0;
    for (int b = 0; b < B; ++b) {
      s.t[b] = 0; s.step[b] = 1; s.token[b] = blank; s.cur[b] = 0; s.count[b] = 0; out_counts[b] = 0;
      s.need_g[b] = seg_len[b] > 0 ? 1 : 0;
      if (seg_len[b] > 0) { s.active[n] = b; s.act_cur[n] = 0; s.act_tok[n] = blank; ++n; }
    }
    for (int i = n; i < B; ++i) s.active[i] = -1;
    // in-cluster control: iteration 0 reads the list of parity 1 and appends to parity 0
    for (int i = 0; i < n; ++i) { list_b[B + i] = s.active[i]; list_cur[B + i] = 0; list_tok[B + i] = blank; }
    cnt2[1] = n; rem2[1] = n; cnt2[0] = 0; rem2[0] = 0;
    *s.n_active = n; *s.remaining = n; *s.overflow = 0;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// The whole search as ONE persistent cooperative kernel (round 2).  The launch-per-phase version above spends most of an
// iteration on dependent-launch latency (about 4.5 us per launch) and on re-reading the predictor / joint weights from L2 behind
// every launch (41.5 us per lock-step iteration).  Here every CTA stays resident for the whole search, keeps its slice of every
// weight matrix in shared memory (fp32: LSTM gate rows of its hidden units, rows of the composed pred_ffn o projection, a
// 32-entry vocabulary tile of ffn_out stored [k][entry]), and the phases of an iteration are separated by grid barriers
// (one atomic per CTA on a monotonic counter, bounded spin):
//   LSTM layer 0 .. L-1  (weight-stationary: CTA c owns hidden units c, c + G, ...; all utterances that just emitted)   | barrier each
//   projection g_b = Wc h_top'(b) + bc   (CTA c owns rows c, c + G, ...)                                                   | barrier
//   joint: CTA (vocabulary tile vt, utterance group ug): tanh(E[t_b .. t_b + 8) + g_b) . ffn_out[tile]^T, per-tile argmax  | barrier
//   decide: greedy control of utterance b in CTA b mod G (same rules as rnnt_decide_kernel), next iteration's list         | barrier
// Everything that another CTA wrote is read with ld.global.cg (L1 is not coherent); arithmetic is the same fp32 as above.
// ---------------------------------------------------------------------------------------------------------------------
struct RnntPersistParams {
  const float* embed;                       // [V, E]
  const float* w_ih[8]; const float* w_hh[8]; const float* b_ih[8]; const float* b_hh[8];
  const float* Wc; const float* bc;         // [J, H], [J]
  const float* Wo; const float* bo;         // [V, J] (row-major, as in the checkpoint), [V]
  const float* E;                           // [rows, J] enc_ffn output
  const long long* seg_start; const int* seg_len;
  int layers, Em, H, J, V, B, n_steps, cap, blank;
  int NV, UG;                               // vocabulary tiles of 32 entries, utterance groups (NV * UG <= grid)
  int* list_b; int* list_cur; int* list_tok; int* cnt2; int* rem2;     // [2][B] x 3, [2], [2]: double-buffered by iteration parity
  long long* out_tokens; int* out_frames; int* out_counts;
  unsigned* barrier;                        // [1] zeroed before the launch
  unsigned* jdone;                          // [B] vocabulary tiles delivered per utterance (monotonic), zeroed before the launch
  int* iterations;                          // [1] iterations run (for the host)
  long long max_iters;
};

constexpr int RNNT_PMAXB = 128;    // utterances the persistent kernel handles (its active list lives in shared memory)
constexpr int RNNT_PUT = 8;        // utterances per LSTM input tile (shared memory)
constexpr int RNNT_PJT = 16;       // utterances per projection input tile
inline size_t rnnt_persist_smem_bytes(int layers, int Em, int H, int J, int G) {
  const size_t upc = (H + G - 1) / G, rpc = (J + G - 1) / G;
  size_t fl = 0;
  for (int l = 0; l < layers; ++l) fl += upc * 4 * size_t((l == 0 ? Em : H) + H);
  fl += rpc * size_t(H);                    // Wc rows
  fl += size_t(J) * 32;                     // ffn_out tile [k][32]
  // scratch: LSTM input vectors of RNNT_PUT utterances | projection inputs of RNNT_PJT utterances | 2 x 8 activation rows
  fl += std::max(std::max(RNNT_PUT * size_t(std::max(Em, H) + H), RNNT_PJT * size_t(H)), 2 * RNNT_FB * size_t(J));
  fl += 8 * 8 * 32;                         // K-slice partial sums [8 warps][8 frames][32 entries]
  return fl * 4 + 3 * 4 * RNNT_PMAXB;       // + the iteration's active list
}

CF_DEVINL unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
CF_DEVINL void rnnt_grid_barrier(unsigned* counter, unsigned& epoch, unsigned G) {
  __syncthreads();
  ++epoch;
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    const unsigned target = epoch * G;
    unsigned spins = 0;
    while (ld_acquire_u32(counter) < target)
      if (++spins > (1u << 25)) __trap();   // a scheduling bug must trap, never hang the GPU
    __threadfence();
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256, 1) rnnt_persistent_kernel(RnntPersistParams p, RnntState s) {
  extern __shared__ __align__(16) float psm[];
  __shared__ int s_last;
  const int G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int H = p.H, J = p.J, B = p.B;
  const int upc = (H + G - 1) / G, rpc = (J + G - 1) / G;
  // ---- carve shared memory and load this CTA's weight slices once
  float* sW[8];
  float* cur_ptr = psm;
  for (int l = 0; l < p.layers; ++l) { sW[l] = cur_ptr; cur_ptr += size_t(upc) * 4 * ((l == 0 ? p.Em : H) + H); }
  float* sWc = cur_ptr; cur_ptr += size_t(rpc) * H;
  float* sWo = cur_ptr; cur_ptr += size_t(J) * 32;
  float* sX = cur_ptr;                                       // scratch (see rnnt_persist_smem_bytes)
  {
    const size_t a = RNNT_PUT * size_t(max(p.Em, H) + H), b = RNNT_PJT * size_t(H), c = 2 * RNNT_FB * size_t(J);
    cur_ptr += max(max(a, b), c);
  }
  float* sRed = cur_ptr; cur_ptr += 8 * 8 * 32;              // [8 warps][8 frames][32 entries]
  int* sLb = reinterpret_cast<int*>(cur_ptr);                // this iteration's active list: utterance, committed half, token
  int* sLc = sLb + RNNT_PMAXB;
  int* sLt = sLc + RNNT_PMAXB;
  for (int l = 0; l < p.layers; ++l) {
    const int In = l == 0 ? p.Em : H, K = In + H;
    for (int i = 0; i < upc; ++i) {
      const int j = cta + G * i;
      for (int q = 0; q < 4; ++q) {
        float* dst = sW[l] + (size_t(i) * 4 + q) * K;
        for (int k = tid; k < K; k += 256)
          dst[k] = j < H ? (k < In ? __ldg(p.w_ih[l] + size_t(q * H + j) * In + k) : __ldg(p.w_hh[l] + size_t(q * H + j) * H + (k - In))) : 0.f;
      }
    }
  }
  for (int i = 0; i < rpc; ++i) {
    const int r = cta + G * i;
    for (int k = tid; k < H; k += 256) sWc[size_t(i) * H + k] = r < J ? __ldg(p.Wc + size_t(r) * H + k) : 0.f;
  }
  const int vt = cta % p.NV, ug = cta / p.NV;
  const bool joint_cta = ug < p.UG;
  if (joint_cta)
    for (int i = tid; i < J * 32; i += 256) {
      const int k = i >> 5, e = i & 31, v = vt * 32 + e;
      sWo[i] = v < p.V ? __ldg(p.Wo + size_t(v) * J + k) : 0.f;
    }
  __syncthreads();

  const size_t half = size_t(p.layers) * B * H;
  unsigned epoch = 0;
  long long it = 0;
  for (;; ++it) {
    const int par = int(it & 1);
    const int n_act = __ldcg(p.cnt2 + (par ^ 1));
    if (cta == 0 && tid == 0) { p.cnt2[par] = 0; p.rem2[par] = 0; }      // appended to by this iteration's decide step
    __syncthreads();
    for (int i = tid; i < n_act; i += 256) {                              // one round trip for the whole list
      sLb[i] = __ldcg(p.list_b + (par ^ 1) * B + i);
      sLc[i] = __ldcg(p.list_cur + (par ^ 1) * B + i);
      sLt[i] = __ldcg(p.list_tok + (par ^ 1) * B + i);
    }
    __syncthreads();
    // ------------------------------------------------ predictor: LSTM layers for the utterances that just emitted
    for (int l = 0; l < p.layers; ++l) {
      const int In = l == 0 ? p.Em : H, K = In + H, K4 = K / 4, In4 = In / 4;
      for (int u0 = 0; u0 < n_act; u0 += RNNT_PUT) {
        const int nu = min(RNNT_PUT, n_act - u0);
        __syncthreads();                                      // sX free
        for (int i = tid; i < nu * K4; i += 256) {            // independent 16-byte loads: one L2 round trip for the tile
          const int u = i / K4, k4 = i - u * K4;
          const int b = sLb[u0 + u], cur = sLc[u0 + u];
          float4 v;
          if (k4 < In4) v = l == 0 ? __ldg(reinterpret_cast<const float4*>(p.embed + size_t(sLt[u0 + u]) * In) + k4)
                                   : __ldcg(reinterpret_cast<const float4*>(s.h + (cur ^ 1) * half + (size_t(l - 1) * B + b) * H) + k4);
          else v = __ldcg(reinterpret_cast<const float4*>(s.h + cur * half + (size_t(l) * B + b) * H) + (k4 - In4));
          reinterpret_cast<float4*>(sX)[size_t(u) * K4 + k4] = v;
        }
        __syncthreads();
        for (int item = warp; item < upc * nu; item += 8) {   // (unit slot, utterance of the tile)
          const int i = item % upc, u = item / upc;
          const int j = cta + G * i;
          if (j >= H) continue;
          const int b = sLb[u0 + u], cur = sLc[u0 + u];
          const size_t at = (size_t(l) * B + b) * H + j;
          float c_old = 0.f, bias[4] = {0.f, 0.f, 0.f, 0.f};
          if (lane == 0) {                                    // requested before the dot products, consumed after them
            c_old = __ldcg(s.c + cur * half + at);
#pragma unroll
            for (int q = 0; q < 4; ++q) bias[q] = __ldg(p.b_ih[l] + q * H + j) + __ldg(p.b_hh[l] + q * H + j);
          }
          const float4* w4 = reinterpret_cast<const float4*>(sW[l] + size_t(i) * 4 * K);
          const float4* x4 = reinterpret_cast<const float4*>(sX + size_t(u) * K);
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
          for (int k4 = lane; k4 < K4; k4 += 32) {
            const float4 x = x4[k4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 w = w4[size_t(q) * K4 + k4];
              acc[q] += w.x * x.x + w.y * x.y + w.z * x.z + w.w * x.w;
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[q] = warp_sum(acc[q]);
          if (lane == 0) {
            float gate[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) gate[q] = acc[q] + bias[q];
            const float c_new = sigmoidf_acc(gate[1]) * c_old + sigmoidf_acc(gate[0]) * tanhf(gate[2]);
            s.c[(cur ^ 1) * half + at] = c_new;
            s.h[(cur ^ 1) * half + at] = sigmoidf_acc(gate[3]) * tanhf(c_new);
          }
        }
      }
      rnnt_grid_barrier(p.barrier, epoch, G);
    }
    // ------------------------------------------------ projection rows of this CTA for the same utterances
    for (int u0 = 0; u0 < n_act; u0 += RNNT_PJT) {
      const int nu = min(RNNT_PJT, n_act - u0), H4 = H / 4;
      __syncthreads();
      for (int i = tid; i < nu * H4; i += 256) {
        const int u = i / H4, k4 = i - u * H4;
        const int b = sLb[u0 + u], cur = sLc[u0 + u];
        reinterpret_cast<float4*>(sX)[size_t(u) * H4 + k4] =
            __ldcg(reinterpret_cast<const float4*>(s.h + (cur ^ 1) * half + (size_t(p.layers - 1) * B + b) * H) + k4);
      }
      __syncthreads();
      for (int item = warp; item < rpc * nu; item += 8) {
        const int i = item % rpc, u = item / rpc;
        const int r = cta + G * i;
        if (r >= J) continue;
        const float4* x4 = reinterpret_cast<const float4*>(sX + size_t(u) * H);
        const float4* w4 = reinterpret_cast<const float4*>(sWc + size_t(i) * H);
        float acc = 0.f;
        for (int k4 = lane; k4 < H4; k4 += 32) {
          const float4 x = x4[k4], w = w4[k4];
          acc += w.x * x.x + w.y * x.y + w.z * x.z + w.w * x.w;
        }
        acc = warp_sum(acc);
        if (lane == 0) s.g[size_t(sLb[u0 + u]) * J + r] = acc + __ldg(p.bc + r);
      }
    }
    rnnt_grid_barrier(p.barrier, epoch, G);
    // ------------------------------------------------ joint over the next 8 frames of the unfinished utterances of this group, two
    // utterances per round (warps 0-3: first, 4-7: second; a warp = one quarter of K, lane = vocabulary entry of the tile);
    // the CTA that delivers the last vocabulary tile of an utterance runs its greedy control (greedy_search.py:24-76)
    if (joint_cta) {
      for (int b0 = ug; b0 < B; b0 += 2 * p.UG) {
        int bb[2], t0[2], nf[2];
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          bb[h2] = b0 + h2 * p.UG;
          t0[h2] = bb[h2] < B ? __ldcg(s.t + bb[h2]) : 0;
          nf[h2] = bb[h2] < B ? max(0, min(RNNT_FB, p.seg_len[bb[h2]] - t0[h2])) : 0;
        }
        if (nf[0] == 0 && nf[1] == 0) continue;               // uniform over the CTA
        __syncthreads();                                      // sX / sRed free
        for (int i = tid; i < 2 * RNNT_FB * J; i += 256) {
          const int h2 = i / (RNNT_FB * J), r = i - h2 * RNNT_FB * J, f = r / J, k = r - f * J;
          float v = 0.f;
          if (f < nf[h2]) {
            const float* e = p.E + (p.seg_start[bb[h2]] + t0[h2] + f) * (long long)J;
            v = tanhf(__ldg(e + k) + __ldcg(s.g + size_t(bb[h2]) * J + k));
          }
          sX[i] = v;
        }
        __syncthreads();
        {
          const int h2 = warp >> 2, ks = warp & 3, kper = J / 4;
          const float* ax = sX + size_t(h2) * RNNT_FB * J;
          float acc[RNNT_FB];
#pragma unroll
          for (int f = 0; f < RNNT_FB; ++f) acc[f] = 0.f;
          if (nf[h2] > 0)
            for (int k0 = ks * kper; k0 < (ks + 1) * kper; k0 += 4) {
              const float w0 = sWo[(k0 + 0) * 32 + lane], w1 = sWo[(k0 + 1) * 32 + lane];
              const float w2 = sWo[(k0 + 2) * 32 + lane], w3 = sWo[(k0 + 3) * 32 + lane];
#pragma unroll
              for (int f = 0; f < RNNT_FB; ++f) {
                const float4 a = *reinterpret_cast<const float4*>(ax + size_t(f) * J + k0);
                acc[f] = fmaf(a.x, w0, acc[f]); acc[f] = fmaf(a.y, w1, acc[f]);
                acc[f] = fmaf(a.z, w2, acc[f]); acc[f] = fmaf(a.w, w3, acc[f]);
              }
            }
#pragma unroll
          for (int f = 0; f < RNNT_FB; ++f) sRed[(warp * RNNT_FB + f) * 32 + lane] = acc[f];
        }
        __syncthreads();
        // warp w finishes frames w and w + ... : 16 (utterance, frame) pairs over 8 warps
        for (int pf = warp; pf < 2 * RNNT_FB; pf += 8) {
          const int h2 = pf / RNNT_FB, f = pf - h2 * RNNT_FB, v = vt * 32 + lane;
          if (f >= nf[h2]) continue;
          float sum = 0.f;
#pragma unroll
          for (int q = 0; q < 4; ++q) sum += sRed[((h2 * 4 + q) * RNNT_FB + f) * 32 + lane];
          float val = v < p.V ? sum + __ldg(p.bo + v) : -INFINITY;
          int idx = v < p.V ? v : 0x7fffffff;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, val, o);
            const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
            if (ov > val || (ov == val && oi < idx)) { val = ov; idx = oi; }
          }
          if (lane == 0) {
            const size_t at = (size_t(bb[h2]) * RNNT_FB + f) * p.NV + vt;
            s.part_val[at] = val;
            s.part_idx[at] = idx;
          }
        }
        // last vocabulary tile of an utterance to arrive -> this CTA decides for it
        __syncthreads();
        if (tid == 0) {
          __threadfence();
          int last = 0;
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2)
            if (nf[h2] > 0 && (atomicAdd(p.jdone + bb[h2], 1u) % unsigned(p.NV)) == unsigned(p.NV - 1)) last |= 1 << h2;
          if (last) __threadfence();
          s_last = last;
        }
        __syncthreads();
        const int last = s_last;
        if (last && warp == 0) {
#pragma unroll 1
          for (int h2 = 0; h2 < 2; ++h2) {
            if (!((last >> h2) & 1)) continue;
            const int b = bb[h2];
            int t = t0[h2];
            const int len = p.seg_len[b];
            const int f_l = lane >> 2, q_l = lane & 3;
            const size_t at = (size_t(b) * RNNT_FB + f_l) * p.NV;
            float val = -INFINITY; int idx = 0x7fffffff;
            for (int w = q_l; w < p.NV; w += 4) {
              const float ov = __ldcg(s.part_val + at + w); const int oi = __ldcg(s.part_idx + at + w);
              if (ov > val || (ov == val && oi < idx)) { val = ov; idx = oi; }
            }
#pragma unroll
            for (int o = 1; o <= 2; o <<= 1) {
              const float ov = __shfl_xor_sync(0xffffffffu, val, o);
              const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
              if (ov > val || (ov == val && oi < idx)) { val = ov; idx = oi; }
            }
            const unsigned nb = __ballot_sync(0xffffffffu, q_l == 0 && idx != p.blank && t + f_l < len);
            const int limit = min(RNNT_FB, len - t);
            const int f_hit = nb ? (__ffs(nb) - 1) >> 2 : limit;
            const int sym = __shfl_sync(0xffffffffu, idx, (f_hit < RNNT_FB ? f_hit : 0) * 4);
            if (lane == 0) {
              int step = __ldcg(s.step + b);
              int n_out = __ldcg(s.count + b);
              const int cur_old = __ldcg(s.cur + b);
              if (f_hit > 0) step = 1;                  // blanks before the hit: next frame, slot 1
              t += f_hit;
              if (f_hit < limit) {
                if (n_out >= p.cap) { *s.overflow = 1; t = len; }
                else {
                  p.out_tokens[size_t(b) * p.cap + n_out] = sym;
                  p.out_frames[size_t(b) * p.cap + n_out] = t;
                  s.count[b] = ++n_out;
                  const int cur = cur_old ^ 1;          // the candidate predictor state becomes the committed one
                  s.cur[b] = cur;
                  if (++step > p.n_steps) { ++t; step = 1; }
                  if (t < len) {
                    const int slot = atomicAdd(p.cnt2 + par, 1);
                    p.list_b[par * B + slot] = b; p.list_cur[par * B + slot] = cur; p.list_tok[par * B + slot] = sym;
                  }
                }
              }
              s.t[b] = t; s.step[b] = step;
              if (t < len) atomicAdd(p.rem2 + par, 1);
              else p.out_counts[b] = n_out;
            }
          }
        }
      }
    }
    rnnt_grid_barrier(p.barrier, epoch, G);
    if (__ldcg(p.rem2 + par) == 0 || it + 1 >= p.max_iters) break;
  }
  if (cta == 0 && tid == 0) *p.iterations = int(it + 1);
}

}  // namespace cf
