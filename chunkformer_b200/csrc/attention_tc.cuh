// Fused chunked relative-position attention on tcgen05 / TMEM, fed by TMA (attention.py:420-505 + 104-150).
//
// Specialisation: chunk 64, d_k 64, left/right context multiples of 64 with l + r <= 256 (benchmark 64/128/128).
// One work item = (pair of consecutive chunks g0 = 2p, g1 = 2p+1; head h): the 128 query rows of the pair fill one
// UMMA M=128 tile and share the union key window  rows [64 g0, 64 g0 + W + 64)  of the flat K/V buffer (window of
// chunk g = rows [64 g, 64 g + W); the 5x unfold copy of the reference never exists).  Keys are processed in blocks
// of 128 union slots with an online softmax:
//     S_ac = (Q+u) K_blk^T                      UMMA 128x128x64  -> TMEM cols [0,128)
//     S_bd = (Q+v) P[128b-64 .. 128b+192)^T     UMMA 128x256x64  -> TMEM cols [128,384)
//     s[rho, kk] = S_ac[rho, kk] + S_bd[rho, 127 - rho + kk]      (rel_shift, attention.py:242-266: the same skew
//                                                                  formula holds for both chunks of the pair)
//     P = exp2(s*scale*log2e - m)  (bf16, kept in TMEM as the A operand),  O += P V_blk  UMMA 128x64x128
// The row-dependent skew goes through a thread-private shared-memory row (each thread owns one query row after
// tcgen05.ld 32x32b), so no cross-thread synchronisation is needed for it.  The score matrix never reaches HBM.
//
//   warp 0 : TMA producer (Q tiles per item, K/V blocks in a 2-stage ring, the head's position table once)
//   warp 1 : TMEM allocator + single-thread MMA issuer (S of block j+1 is issued before waiting for P of block j)
//   warps 2..9 : softmax / correction / output; two threads share a query row (64 of each block's 128 keys each),
//                exchanging row maxima / sums through shared memory; fp16 staging of the pre-scaled S_bd values
#pragma once
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "attention_simt.cuh"
#include "gemm_host.cuh"

namespace cf {

constexpr bool kAttentionTcReady = true;

constexpr int ATC_THREADS = 320;                 // TMA warp, MMA warp, 8 softmax warps
constexpr int ATC_STAGE_PITCH = 144;             // bytes per thread-private skew row (64 fp16 + pad; 16-byte stores conflict free)
constexpr uint32_t ATC_PTAB_ROWS = 448;          // table rows -64 .. 383 of the head
constexpr uint32_t ATC_PTAB_BYTES = ATC_PTAB_ROWS * 128;
constexpr uint32_t ATC_TILE_BYTES = 128 * 128;   // 128 rows x 64 bf16
constexpr int ATC_KV_STAGES = 2;
constexpr size_t ATC_SMEM_BYTES = ATC_PTAB_BYTES + 2 * ATC_TILE_BYTES /*Qu,Qv*/ + 2 * ATC_KV_STAGES * ATC_TILE_BYTES /*K,V ring*/ +
                                  256 * ATC_STAGE_PITCH + 2048 + 1024 + 1024 + 128;

CF_DEVINL float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
CF_DEVINL uint32_t pack_half2(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

CF_DEVINL void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
CF_DEVINL void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

CF_DEVINL bool mbar_test(uint64_t* bar, uint32_t parity) {     // non-blocking probe
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}

__device__ const int kSkewRowPerm[32] = {20, 3, 6, 13, 31, 18, 17, 0, 7, 25, 11, 5, 10, 22, 24, 28,
                                       9, 15, 29, 27, 30, 2, 8, 4, 12, 21, 26, 19, 14, 1, 16, 23};

// The 32 skew values a thread needs start at half (31 - lane) of its private row.  Read as 17 aligned 32-bit words and realign
// with one funnel shift per pair: under kSkewRowPerm the word reads are bank-conflict free (17 wavefronts per warp instead of
// the 48 of thirty-two 2-byte reads; tools/skew_row_perm.py).
struct SkewWords {
  uint32_t w[17];
  uint32_t sh;
  CF_DEVINL void load(const uint8_t* row, int lane) {
    const uint32_t* p32 = reinterpret_cast<const uint32_t*>(row) + ((31 - lane) >> 1);
    sh = uint32_t((31 - lane) & 1) * 16u;
#pragma unroll
    for (int i = 0; i < 17; ++i) w[i] = p32[i];
  }
  CF_DEVINL float2 pair(int j) const {              // halves (31 - lane) + 2 j and + 2 j + 1
    const uint32_t pr = __funnelshift_r(w[j], w[j + 1], sh);
    return __half22float2(*reinterpret_cast<const __half2*>(&pr));
  }
};

struct AttnTcParams {
  const int2* range;      // [n_chunks + 2] valid key slots per chunk (entries beyond n_chunks are empty)
  __nv_bfloat16* ctx;     // [n_chunks * 64, d]
  int n_chunks, n_pairs, l, d, heads, nb;   // n_pairs = tiles of 128 query rows; nb = key blocks of 128 union slots
  int c_log2, tab_row0, n_last;              // chunk size (log2); first resident table row (c - 128); S_bd columns of the last block
  int items_per_cta_stride;                  // CTAs per head
  float scale_log2e;
#ifdef CF_ABLATION
  long long* prof = nullptr;                 // tools build: [grid][16] cycles per phase of softmax warp 0 / the MMA warp
  int debug = 0;                             // tools build: 1 = no skew (S_bd read unshifted from TMEM, no shared-memory round trip; wrong results)
#endif
};
#ifdef CF_ABLATION
#define ATC_MARK(k) do { const long long _n = clock64(); seg[k] += _n - seg_t; seg_t = _n; } while (0)
#else
#define ATC_MARK(k) do {} while (0)
#endif

// PRE: Q+u / Q+v already carry (1/sqrt(d_k)) * log2(e) (folded into the fused QKV projection at weight load).
template <bool PRE>
__global__ void __launch_bounds__(ATC_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tma_qkv, const __grid_constant__ CUtensorMap tma_pos, AttnTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
  uint8_t* s_ptab = smem;                        // smem row rr <-> table row rr - 64
  uint8_t* s_qu = s_ptab + ATC_PTAB_BYTES;
  uint8_t* s_qv = s_qu + ATC_TILE_BYTES;
  uint8_t* s_k = s_qv + ATC_TILE_BYTES;                      // [ATC_KV_STAGES]
  uint8_t* s_v = s_k + ATC_KV_STAGES * ATC_TILE_BYTES;       // [ATC_KV_STAGES]
  uint8_t* s_stage = s_v + ATC_KV_STAGES * ATC_TILE_BYTES;   // 256 thread-private rows
  float* s_xch = reinterpret_cast<float*>(s_stage + 256 * ATC_STAGE_PITCH);   // [2 parity][2 set][128] row maxima
  float* s_lx = s_xch + 512;                                                   // [2 set][128] row sums
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_lx + 256);
  uint64_t* ptab_full = bars + 0;
  uint64_t* q_full = bars + 1;
  uint64_t* q_empty = bars + 2;
  uint64_t* kv_full = bars + 3;    // [ATC_KV_STAGES]
  uint64_t* kv_empty = bars + 6;   // [ATC_KV_STAGES]
  uint64_t* s_full = bars + 9;
  uint64_t* s_free = bars + 10;
  uint64_t* p_full = bars + 11;
  uint64_t* pv_done = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp index provably uniform
  const int h = blockIdx.x % p.heads;
  const int first_pair = blockIdx.x / p.heads;
  const int d = p.d;
  const int nb = p.nb;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tma_qkv);
    tma_prefetch_desc(&tma_pos);
    mbar_init(ptab_full, 1);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int s = 0; s < ATC_KV_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    mbar_init(s_full, 1);
    mbar_init(s_free, 8);      // one arrival per softmax warp
    mbar_init(p_full, 8);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t TM_AC = 0, TM_BD = 128, TM_O = 384, TM_P = 448;   // P (bf16, 128 keys = 64 packed columns) stays in TMEM

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (convergent warp, elected lane issues)
    {
      if (elect_one()) {
        mbar_arrive_expect_tx(ptab_full, ATC_PTAB_BYTES);
        for (int i = 0; i < 7; ++i)                                            // table rows c-128 .. c+319 (rows < 0 read as 0)
          tma_load_2d(s_ptab + i * 64 * 128, &tma_pos, ptab_full, h * 64, p.tab_row0 + 64 * i);
      }
      __syncwarp();
      uint32_t item = 0, blk = 0;
      for (int pair = first_pair; pair < p.n_pairs; pair += p.items_per_cta_stride, ++item) {
        const int trow = 128 * pair;                 // first flat row of the tile = first buffer row of its union key window
        mbar_wait(q_empty, (item & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(q_full, 2 * ATC_TILE_BYTES);
          tma_load_2d(s_qu, &tma_qkv, q_full, h * 64, p.l + trow);
          tma_load_2d(s_qv, &tma_qkv, q_full, d + h * 64, p.l + trow);
        }
        __syncwarp();
        for (int b = 0; b < nb; ++b, ++blk) {
          const uint32_t st = blk % ATC_KV_STAGES, ph = (blk / ATC_KV_STAGES) & 1;
          mbar_wait(&kv_empty[st], ph ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&kv_full[st], 2 * ATC_TILE_BYTES);
            tma_load_2d(s_k + st * ATC_TILE_BYTES, &tma_qkv, &kv_full[st], 2 * d + h * 64, trow + 128 * b);
            tma_load_2d(s_v + st * ATC_TILE_BYTES, &tma_qkv, &kv_full[st], 3 * d + h * 64, trow + 128 * b);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (convergent warp, elected lane issues)
    {
      constexpr uint32_t idesc_ac = make_idesc_bf16(128, 128);
      constexpr uint32_t idesc_bd = make_idesc_bf16(128, 256);
      const uint32_t idesc_bd_last = p.n_last == 192 ? make_idesc_bf16(128, 192) : make_idesc_bf16(128, 256);   // 192: the last block never reads resident table rows >= 448
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);   // A = P from TMEM, B (= V) is MN-major
      mbar_wait(ptab_full, 0);
      const uint64_t dqu = make_sw128_desc(smem_u32(s_qu)), dqv = make_sw128_desc(smem_u32(s_qv));
      const uint64_t dk0 = make_sw128_desc(smem_u32(s_k)), dv0 = make_sw128_desc(smem_u32(s_v));
      const uint64_t dp0 = make_sw128_desc(smem_u32(s_ptab));
      uint32_t item = 0, blk = 0;
      bool have_prev = false;
      uint32_t prev_st = 0, prev_b = 0;
      auto issue_pv = [&](uint32_t pblk, uint32_t st, uint32_t b) {
        mbar_wait(p_full, pblk & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t db = dv0 + uint64_t((st * ATC_TILE_BYTES) >> 4);
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            // A: 16 keys = 8 packed P columns in TMEM; B (MN-major): 16 key rows = 2048 B
            umma_bf16_ts(tmem_base + TM_O, tmem_base + TM_P + 8 * t, db + uint64_t((t * 2048) >> 4), idesc_pv, (b | t) != 0);
          }
          umma_commit(pv_done);
          umma_commit(&kv_empty[st]);
        }
        __syncwarp();
      };
      for (int pair = first_pair; pair < p.n_pairs; pair += p.items_per_cta_stride, ++item) {
        mbar_wait(q_full, item & 1);
        for (int b = 0; b < nb; ++b, ++blk) {
          const uint32_t st = blk % ATC_KV_STAGES, ph = (blk / ATC_KV_STAGES) & 1;
          mbar_wait(&kv_full[st], ph);
          mbar_wait(s_free, (blk & 1) ^ 1);            // softmax finished reading the previous S block
          tc_fence_after();
          if (elect_one()) {
            const uint64_t dk = dk0 + uint64_t((st * ATC_TILE_BYTES) >> 4);
            const uint64_t dp = dp0 + uint64_t((b * 128 * 128) >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base + TM_AC, dqu + 2 * k, dk + 2 * k, idesc_ac, k != 0);
            const uint32_t idbd = (b == nb - 1) ? idesc_bd_last : idesc_bd;
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base + TM_BD, dqv + 2 * k, dp + 2 * k, idbd, k != 0);
            umma_commit(s_full);
            if (b == nb - 1) umma_commit(q_empty);     // Q tiles may be overwritten once these MMAs retire
          }
          __syncwarp();
          if (have_prev) issue_pv(blk - 1, prev_st, prev_b);
          have_prev = true; prev_st = st; prev_b = b;
        }
      }
      if (have_prev) issue_pv(blk - 1, prev_st, prev_b);
    }
  } else {
    // ------------------------------------------------------------------ softmax warps
    // thread = (query row rho, key half `set` of every 128-key block); two threads share a row
    const int sw = warp - 2;
    const int quad = warp & 3;                         // TMEM lane quadrant this warp may access
    const int set = sw >> 2;
    const int rho = quad * 32 + lane;
    const int cj = rho >> p.c_log2;                    // chunk of this row inside the tile
    const int uoff = cj << p.c_log2;                   // its window starts at this union slot
    const uint32_t lane_addr = uint32_t(quad * 32) << 16;
    // skew row of this thread: a lane permutation inside the warp's 32 rows (found by local search, tools/skew_row_perm.py)
    // cuts the bank-conflict replays of the 2-byte skew reads from 2.0 to 1.5 wavefronts per load at the 144-byte pitch
    // while the 16-byte staging stores stay conflict free
    uint8_t* stage = s_stage + ((warp - 2) * 32 + kSkewRowPerm[lane]) * ATC_STAGE_PITCH;
    const int cb_thread = 96 - 32 * quad + 64 * set;   // first S_bd column this warp stages (warp-uniform)
    uint32_t blk = 0;
    int ep_g = -1;                                     // chunk (of this row) whose item is finished but not yet written out
    long long ep_row0 = 0;                             // first flat row of that item's tile
    float ep_l = 0.f;
    // O / l -> ctx for the finished item (this set writes 32 of the head's 64 output columns).  Called once the item's
    // last P V has retired and before this thread lets the next item's first P V start, so it costs no extra wait.
    auto write_out = [&]() {
      const float l_tot = ep_l + s_lx[(set ^ 1) * 128 + rho];
      const float inv = l_tot > 0.f ? 1.0f / l_tot : 0.f;        // no valid key: zero context (attention.py:133-136)
      uint32_t r[32];
      tmem_ld32(tmem_base + lane_addr + TM_O + 32 * set, r);
      tmem_ld_wait();
      if (ep_g < p.n_chunks) {
        __nv_bfloat16* orow = p.ctx + ((long long)ep_row0 + rho) * d + h * 64 + 32 * set;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          uint32_t o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = pack_bf16(__uint_as_float(r[16 * q + 2 * e]) * inv, __uint_as_float(r[16 * q + 2 * e + 1]) * inv);
          asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(orow + 16 * q), "r"(o[0]), "r"(o[1]), "r"(o[2]),
                       "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7])
                       : "memory");
        }
      }
      ep_g = -1;
    };
#ifdef CF_ABLATION
    long long seg[8] = {0, 0, 0, 0, 0, 0, 0, 0}, seg_t = clock64();
    const long long t_begin = seg_t;
#endif
    for (int pair = first_pair; pair < p.n_pairs; pair += p.items_per_cta_stride) {
      const int g = (pair << (7 - p.c_log2)) + cj;
      const int2 rg = p.range[g];
      const int ulo = rg.x + uoff, uhi = rg.y + uoff;   // valid union slots for this row
      float m_run = -1e30f, l_run = 0.f;
      for (int b = 0; b < nb; ++b, ++blk) {
        ATC_MARK(7);
        mbar_wait(s_full, blk & 1);
        tc_fence_after();
        ATC_MARK(0);
        float s[64];
        float mx = -1e30f;
        uint4 keep[4];                                 // packed fp16 S_bd columns [32, 64) of this thread's window:
                                                       // staged as the upper half for sub-block 0, lower half for sub-block 1
#pragma unroll
        for (int sb = 0; sb < 2; ++sb) {
          uint32_t r0[32];
          const uint32_t cbase = TM_BD + cb_thread;
#ifdef CF_ABLATION
          if (p.debug & 1) {                           // timing experiment: what the skew through shared memory costs
            uint32_t rb[32];
            tmem_ld32(tmem_base + lane_addr + cbase + 32 * sb, rb);
            tmem_ld32(tmem_base + lane_addr + TM_AC + 64 * set + 32 * sb, r0);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              const float v = __uint_as_float(r0[k]) + __uint_as_float(rb[k]);
              s[32 * sb + k] = v;
              mx = fmaxf(mx, v);
            }
            continue;
          }
#endif
          const float sc = PRE ? 1.0f : p.scale_log2e;
          auto pack4 = [&](int q) {
            uint4 w0;
            w0.x = pack_half2(__uint_as_float(r0[8 * q]) * sc, __uint_as_float(r0[8 * q + 1]) * sc);
            w0.y = pack_half2(__uint_as_float(r0[8 * q + 2]) * sc, __uint_as_float(r0[8 * q + 3]) * sc);
            w0.z = pack_half2(__uint_as_float(r0[8 * q + 4]) * sc, __uint_as_float(r0[8 * q + 5]) * sc);
            w0.w = pack_half2(__uint_as_float(r0[8 * q + 6]) * sc, __uint_as_float(r0[8 * q + 7]) * sc);
            return w0;
          };
          if (sb == 0) {                               // window columns [0, 32) and [32, 64): two TMEM loads
            tmem_ld32(tmem_base + lane_addr + cbase, r0);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(stage + 16 * q) = pack4(q);
            tmem_ld32(tmem_base + lane_addr + cbase + 32, r0);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 4; ++q) { keep[q] = pack4(q); *reinterpret_cast<uint4*>(stage + 64 + 16 * q) = keep[q]; }
          } else {                                     // window columns [32, 64) come from registers, [64, 96) from TMEM
#pragma unroll
            for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(stage + 16 * q) = keep[q];
            tmem_ld32(tmem_base + lane_addr + cbase + 64, r0);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(stage + 64 + 16 * q) = pack4(q);
          }
          tmem_ld32(tmem_base + lane_addr + TM_AC + 64 * set + 32 * sb, r0);
          tmem_ld_wait();
          const int u0 = 128 * b + 64 * set + 32 * sb;
          const bool edge = (u0 < ulo) || (u0 + 32 > uhi);
          SkewWords sk;
          sk.load(stage, lane);
          if (__any_sync(0xffffffffu, edge)) {
#pragma unroll
            for (int k = 0; k < 32; k += 2) {
              const float2 bd = sk.pair(k >> 1);
              float v0 = PRE ? __uint_as_float(r0[k]) + bd.x : fmaf(__uint_as_float(r0[k]), p.scale_log2e, bd.x);
              float v1 = PRE ? __uint_as_float(r0[k + 1]) + bd.y : fmaf(__uint_as_float(r0[k + 1]), p.scale_log2e, bd.y);
              v0 = (u0 + k >= ulo && u0 + k < uhi) ? v0 : -INFINITY;
              v1 = (u0 + k + 1 >= ulo && u0 + k + 1 < uhi) ? v1 : -INFINITY;
              s[32 * sb + k] = v0; s[32 * sb + k + 1] = v1;
              mx = fmaxf(mx, fmaxf(v0, v1));
            }
          } else {
#pragma unroll
            for (int k = 0; k < 32; k += 2) {
              const float2 bd = sk.pair(k >> 1);
              const float v0 = PRE ? __uint_as_float(r0[k]) + bd.x : fmaf(__uint_as_float(r0[k]), p.scale_log2e, bd.x);
              const float v1 = PRE ? __uint_as_float(r0[k + 1]) + bd.y : fmaf(__uint_as_float(r0[k + 1]), p.scale_log2e, bd.y);
              s[32 * sb + k] = v0; s[32 * sb + k + 1] = v1;
              mx = fmaxf(mx, fmaxf(v0, v1));
            }
          }
        }
        ATC_MARK(1);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free);            // this warp's part of S is in registers
        float* xch = s_xch + (blk & 1) * 256;
        xch[set * 128 + rho] = mx;
        named_bar_sync(1 + quad, 64);
        ATC_MARK(2);      // only the two warps that share this quadrant's rows exchange anything
        const float m_new = fmaxf(m_run, fmaxf(mx, xch[(set ^ 1) * 128 + rho]));
        const float alpha = fast_exp2(m_run - m_new);
        float sum = 0.f;
        ATC_MARK(3);
        if (blk > 0) mbar_wait(pv_done, (blk - 1) & 1);   // previous P V retired: P tile and O are ours again
        tc_fence_after();
        ATC_MARK(4);
        if (ep_g >= 0) write_out();                    // previous item: its row sums were published before the barrier above
        {
          uint32_t pk[32];                             // this thread's 64 probabilities as 32 packed bf16 pairs
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float p0 = fast_exp2(s[2 * e] - m_new), p1 = fast_exp2(s[2 * e + 1] - m_new);
            sum += p0 + p1;
            pk[e] = pack_bf16(p0, p1);
          }
          tmem_st32(tmem_base + lane_addr + TM_P + 32 * set, pk);
        }
        ATC_MARK(5);
        l_run = l_run * alpha + sum;
        m_run = m_new;
        if (b > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {   // rescale this set's half of the running output
          uint32_t r[32];
          tmem_ld32(tmem_base + lane_addr + TM_O + 32 * set, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) * alpha);
          tmem_st32(tmem_base + lane_addr + TM_O + 32 * set, r);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        ATC_MARK(6);
      }
      // ---- item finished: publish the row sum (read by the partner thread after the next named barrier) and defer the
      // output to the next block's P V wait
      s_lx[set * 128 + rho] = l_run;
      ep_g = g; ep_l = l_run; ep_row0 = 128LL * pair;
    }
    if (ep_g >= 0) {
      named_bar_sync(1 + quad, 64);      // only the two warps that share this quadrant's rows exchange anything
      mbar_wait(pv_done, (blk - 1) & 1);
      tc_fence_after();
      write_out();
      tc_fence_before();
    }
#ifdef CF_ABLATION
    if (p.prof && (warp == 2 || warp == 7) && lane == 0) {
      long long* pr = p.prof + (size_t(blockIdx.x) * 2 + (warp == 7)) * 16;
      for (int k = 0; k < 8; ++k) pr[k] = seg[k];
      pr[8] = clock64() - t_begin; pr[9] = blk;
    }
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------------------
// General ring kernel: d_k = 64 or 128, any chunk size that divides 128 (tile = 128 / c chunks) or is >= 128 (tiles of 128
// rows inside one chunk: chunk 128 / 256 configurations, full attention = one chunk per utterance), any context sizes.
// Used for everything the 128-key-block kernel above does not cover (chunkformer-rnnt-large / classification geometry with
// d_k = 128, long contexts, long chunks).  Per item (tile, head) the operands no longer fit beside a resident position
// table, so keys go in blocks of 64 (S_ac N = 64, S_bd N = 192) and the projected position table streams through a ring of
// four 64-row chunks: block b reads table rows [tb + 64 b, tb + 64 b + 192) = chunks b, b+1, b+2, of which only b+2 is
// new.  K is double-buffered, V single-buffered, every operand has its own full / empty barrier pair and one polling
// thread issues each TMA load the moment its buffer is free.  P never touches shared memory (TMEM, A operand of P V).
//   score(rho, u) = S_ac[rho, u] + S_bd[rho, 127 - rho + (u - 64 b)]   with table row = tb + 64 b + column, where
//   tb = c - 128 (multi-chunk tiles) or c - 128 - 128 t (tile t of a long chunk);   u = key slot in the tile's window.
// TMEM map: S_ac [0,64)  S_bd [64,256)  O [256,256+DK)  P [256+DK, 288+DK).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int ATCR_THREADS = 320;
constexpr uint32_t ATCR_ATOM = 64 * 128;            // 64 rows x 64 bf16
template <int DK>
constexpr size_t atcr_smem_bytes() {
  return size_t(DK / 64) * (2 + 2 + 2 + 1 + 4) * ATCR_ATOM /*Qu, Qv (128 rows each) + K x2 + V + table ring x4*/ + 256 * ATC_STAGE_PITCH +
         2048 + 1024 + 256 + 1024;
}

struct AttnRingParams {
  const int2* range;      // [n_chunks (+ phantoms)] valid key slots per chunk
  __nv_bfloat16* ctx;     // [n_chunks * c, d]
  int n_chunks, n_tiles, c, l, d, heads, nb;   // nb = key blocks of 64 window slots
  int c_log2;             // c < 128: log2(c) (tile = 128 / c chunks);  c >= 128: -1 (tiles_per_chunk tiles inside one chunk)
  int tiles_per_chunk;    // c >= 128
  int last_chunks;        // table chunks read by the last block (2 when no valid score there needs columns >= 128)
  int items_per_cta_stride;
  float scale_log2e;
};

struct AttnRingTile { int qrow0, krow0, tb; };
CF_DEVINL AttnRingTile atcr_tile(const AttnRingParams& p, int ti) {
  AttnRingTile t;
  if (p.c_log2 >= 0) { t.qrow0 = 128 * ti; t.krow0 = 128 * ti; t.tb = p.c - 128; }
  else {
    const int g = ti / p.tiles_per_chunk, tt = ti - g * p.tiles_per_chunk;
    t.qrow0 = p.c * g + 128 * tt; t.krow0 = p.c * g; t.tb = p.c - 128 - 128 * tt;
  }
  return t;
}

template <int DK, bool PRE>
__global__ void __launch_bounds__(ATCR_THREADS, 1)
attention_ring_kernel(const __grid_constant__ CUtensorMap tma_q /*box 128 x 64*/, const __grid_constant__ CUtensorMap tma_kv /*box 64 x 64*/,
                      const __grid_constant__ CUtensorMap tma_pos /*box 64 x 64*/, AttnRingParams p) {
  constexpr int NA = DK / 64;                            // 64-column swizzle atoms per operand row
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_qu = smem;                                 // [NA atoms][128 rows x 128 B]
  uint8_t* s_qv = s_qu + NA * 2 * ATCR_ATOM;
  uint8_t* s_k = s_qv + NA * 2 * ATCR_ATOM;             // [2 stages][NA atoms][64 keys x 128 B]
  uint8_t* s_v = s_k + 2 * NA * ATCR_ATOM;              // [NA atoms][64 keys x 128 B]   (MN-major B operand)
  uint8_t* s_pt = s_v + NA * ATCR_ATOM;                 // [4 slots][NA atoms][64 table rows x 128 B]
  uint8_t* s_stage = s_pt + 4 * NA * ATCR_ATOM;         // 256 thread-private skew rows
  float* s_xch = reinterpret_cast<float*>(s_stage + 256 * ATC_STAGE_PITCH);   // [2 parity][2 set][128] row maxima
  float* s_lx = s_xch + 512;                                                   // [2 set][128] row sums
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_lx + 256);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* k_full = bars + 2;     // [2]
  uint64_t* k_empty = bars + 4;    // [2]
  uint64_t* v_full = bars + 6;
  uint64_t* v_empty = bars + 7;
  uint64_t* pt_full = bars + 8;    // [4]
  uint64_t* pt_empty = bars + 12;  // [4]
  uint64_t* s_full = bars + 16;
  uint64_t* s_free = bars + 17;
  uint64_t* p_full = bars + 18;
  uint64_t* pv_done = bars + 19;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int h = blockIdx.x % p.heads;
  const int first_tile = blockIdx.x / p.heads;
  const int d = p.d;
  const int nb = p.nb;
  const int nch = nb + 2;                               // table chunks per item
  const int stride = p.items_per_cta_stride;
  const int n_items = first_tile < p.n_tiles ? (p.n_tiles - first_tile + stride - 1) / stride : 0;
  const uint32_t total = uint32_t(n_items) * uint32_t(nb);
  const uint32_t total_ch = uint32_t(n_items) * uint32_t(nch);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_kv);
    tma_prefetch_desc(&tma_pos);
    mbar_init(q_full, 1); mbar_init(q_empty, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
    mbar_init(v_full, 1); mbar_init(v_empty, 1);
    for (int s = 0; s < 4; ++s) { mbar_init(&pt_full[s], 1); mbar_init(&pt_empty[s], 1); }
    mbar_init(s_full, 1);
    mbar_init(s_free, 8);
    mbar_init(p_full, 8);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t TM_AC = 0, TM_BD = 64, TM_O = 256, TM_P = 256 + DK;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer: four streams polled by one thread
    if (elect_one()) {
      uint32_t kb = 0, vb = 0, cb = 0;
      int qi = 0;
      while (kb < total || vb < total || cb < total_ch || qi < n_items) {
        bool issued = false;
        if (qi < n_items && mbar_test(q_empty, (qi & 1) ^ 1)) {
          const int row = p.l + atcr_tile(p, first_tile + qi * stride).qrow0;
          mbar_arrive_expect_tx(q_full, NA * 4 * ATCR_ATOM);
#pragma unroll
          for (int a = 0; a < NA; ++a) {
            tma_load_2d(s_qu + a * 2 * ATCR_ATOM, &tma_q, q_full, h * DK + 64 * a, row);
            tma_load_2d(s_qv + a * 2 * ATCR_ATOM, &tma_q, q_full, d + h * DK + 64 * a, row);
          }
          ++qi; issued = true;
        }
        if (kb < total && mbar_test(&k_empty[kb & 1], ((kb >> 1) & 1) ^ 1)) {
          const int it = int(kb) / nb, b = int(kb) - it * nb;
          const int row = atcr_tile(p, first_tile + it * stride).krow0 + 64 * b;
          mbar_arrive_expect_tx(&k_full[kb & 1], NA * ATCR_ATOM);
#pragma unroll
          for (int a = 0; a < NA; ++a)
            tma_load_2d(s_k + ((kb & 1) * NA + a) * ATCR_ATOM, &tma_kv, &k_full[kb & 1], 2 * d + h * DK + 64 * a, row);
          ++kb; issued = true;
        }
        if (cb < total_ch && mbar_test(&pt_empty[cb & 3], ((cb >> 2) & 1) ^ 1)) {
          const int it = int(cb / uint32_t(nch)), ct = int(cb) - it * nch;
          const int row = atcr_tile(p, first_tile + it * stride).tb + 64 * ct;        // rows outside [0, R) read as zeros
          mbar_arrive_expect_tx(&pt_full[cb & 3], NA * ATCR_ATOM);
#pragma unroll
          for (int a = 0; a < NA; ++a)
            tma_load_2d(s_pt + ((cb & 3) * NA + a) * ATCR_ATOM, &tma_pos, &pt_full[cb & 3], h * DK + 64 * a, row);
          ++cb; issued = true;
        }
        if (vb < total && mbar_test(v_empty, (vb & 1) ^ 1)) {
          const int it = int(vb) / nb, b = int(vb) - it * nb;
          const int row = atcr_tile(p, first_tile + it * stride).krow0 + 64 * b;
          mbar_arrive_expect_tx(v_full, NA * ATCR_ATOM);
#pragma unroll
          for (int a = 0; a < NA; ++a)
            tma_load_2d(s_v + a * ATCR_ATOM, &tma_kv, v_full, 3 * d + h * DK + 64 * a, row);
          ++vb; issued = true;
        }
        if (!issued) __nanosleep(64);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (convergent warp, elected lane issues)
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 64);
    constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);   // A = P from TMEM, B (= V) MN-major
    const uint64_t dqu = make_sw128_desc(smem_u32(s_qu)), dqv = make_sw128_desc(smem_u32(s_qv));
    const uint64_t dk0 = make_sw128_desc(smem_u32(s_k)), dv0 = make_sw128_desc(smem_u32(s_v)), dpt0 = make_sw128_desc(smem_u32(s_pt));
    auto issue_pv = [&](uint32_t pblk, uint32_t pb) {
      mbar_wait(v_full, pblk & 1);
      mbar_wait(p_full, pblk & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int t = 0; t < 4; ++t)          // 16 keys per step: 8 packed P columns; V: 16 key rows = 2048 B inside an atom
#pragma unroll
          for (int hn = 0; hn < NA; ++hn)
            umma_bf16_ts(tmem_base + TM_O + 64 * hn, tmem_base + TM_P + 8 * t, dv0 + uint64_t((hn * ATCR_ATOM + t * 2048) >> 4),
                         idesc_pv, (pb | uint32_t(t)) != 0);
        umma_commit(pv_done);
        umma_commit(v_empty);
      }
      __syncwarp();
    };
    uint32_t blk = 0, ch0 = 0;                            // ch0: global index of the item's chunk 0
    for (int item = 0; item < n_items; ++item, ch0 += uint32_t(nch)) {
      mbar_wait(q_full, item & 1);
      for (int b = 0; b < nb; ++b, ++blk) {
        const uint32_t st = blk & 1;
        const int nchunks = (b == nb - 1) ? p.last_chunks : 3;
        mbar_wait(&k_full[st], (blk >> 1) & 1);
        for (int j = (b == 0 ? 0 : 2); j < 3; ++j) {     // chunks b, b+1 were waited for by the previous block; the item's last
          const uint32_t gc = ch0 + uint32_t(b + j);     // chunk may not be read but must have landed before its slot is released
          mbar_wait(&pt_full[gc & 3], (gc >> 2) & 1);
        }
        mbar_wait(s_free, (blk & 1) ^ 1);                 // softmax finished reading the previous S block
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int a = 0; a < NA; ++a)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_ss(tmem_base + TM_AC, dqu + uint64_t((a * 2 * ATCR_ATOM) >> 4) + 2 * k,
                           dk0 + uint64_t(((st * NA + a) * ATCR_ATOM) >> 4) + 2 * k, idesc_s, (a | k) != 0);
          for (int j = 0; j < nchunks; ++j) {
            const uint32_t slot = (ch0 + uint32_t(b + j)) & 3;
#pragma unroll
            for (int a = 0; a < NA; ++a)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_ss(tmem_base + TM_BD + 64 * j, dqv + uint64_t((a * 2 * ATCR_ATOM) >> 4) + 2 * k,
                             dpt0 + uint64_t(((slot * NA + a) * ATCR_ATOM) >> 4) + 2 * k, idesc_s, (a | k) != 0);
          }
          umma_commit(s_full);
          umma_commit(&k_empty[st]);
          umma_commit(&pt_empty[(ch0 + uint32_t(b)) & 3]);      // chunk b is not read by later blocks
          if (b == nb - 1) {
            umma_commit(&pt_empty[(ch0 + uint32_t(b + 1)) & 3]);  // ... nor are the item's last chunks
            umma_commit(&pt_empty[(ch0 + uint32_t(b + 2)) & 3]);
            umma_commit(q_empty);
          }
        }
        __syncwarp();
        if (blk > 0) issue_pv(blk - 1, uint32_t(b == 0 ? nb - 1 : b - 1));
      }
    }
    if (total > 0) issue_pv(total - 1, uint32_t(nb - 1));
  } else {
    // ------------------------------------------------------------------ softmax warps: thread = (query row, 32 of the block's 64 keys)
    const int sw = warp - 2;
    const int quad = warp & 3;
    const int set = sw >> 2;
    const int rho = quad * 32 + lane;
    const uint32_t lane_addr = uint32_t(quad * 32) << 16;
    uint8_t* stage = s_stage + ((warp - 2) * 32 + kSkewRowPerm[lane]) * ATC_STAGE_PITCH;   // see attention_tc_kernel
    const int cbw = 96 - 32 * quad + 32 * set;          // first S_bd column this warp stages (warp-uniform)
    constexpr int OC = DK / 2;                          // output columns per thread
    uint32_t blk = 0;
    long long ep_row = -1;                              // flat output row of the finished item (-1: none / not a real row)
    bool ep_pending = false;
    float ep_l = 0.f;
    auto write_out = [&]() {                            // O / l -> ctx: this thread's half of the head's output columns
      const float l_tot = ep_l + s_lx[(set ^ 1) * 128 + rho];
      const float inv = l_tot > 0.f ? 1.0f / l_tot : 0.f;      // no valid key: zero context (attention.py:133-136)
#pragma unroll
      for (int hc = 0; hc < OC / 32; ++hc) {
        uint32_t r[32];
        tmem_ld32(tmem_base + lane_addr + TM_O + OC * set + 32 * hc, r);
        tmem_ld_wait();
        if (ep_row >= 0) {
          __nv_bfloat16* orow = p.ctx + ep_row * d + h * DK + OC * set + 32 * hc;
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            uint32_t o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = pack_bf16(__uint_as_float(r[16 * q + 2 * e]) * inv, __uint_as_float(r[16 * q + 2 * e + 1]) * inv);
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(orow + 16 * q), "r"(o[0]), "r"(o[1]),
                         "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7])
                         : "memory");
          }
        }
      }
      ep_pending = false;
    };
    for (int ti = first_tile; ti < p.n_tiles; ti += stride) {
      // this row's chunk, its valid key slots inside the tile's window and its flat output row
      int g, uoff;
      bool real;
      const AttnRingTile tl = atcr_tile(p, ti);
      if (p.c_log2 >= 0) {
        const int cj = rho >> p.c_log2;
        g = (ti << (7 - p.c_log2)) + cj; uoff = cj << p.c_log2; real = g < p.n_chunks;
      } else {
        g = ti / p.tiles_per_chunk; uoff = 0; real = (tl.qrow0 - tl.krow0) + rho < p.c;
      }
      const int2 rg = p.range[g];
      const int ulo = rg.x + uoff, uhi = rg.y + uoff;
      float m_run = -1e30f, l_run = 0.f;
      for (int b = 0; b < nb; ++b, ++blk) {
        mbar_wait(s_full, blk & 1);
        tc_fence_after();
        float s[32];
        float mx = -1e30f;
        {
          const float sc = PRE ? 1.0f : p.scale_log2e;
#pragma unroll
          for (int hb = 0; hb < 2; ++hb) {
            uint32_t rb[32];
            tmem_ld32(tmem_base + lane_addr + TM_BD + cbw + 32 * hb, rb);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 w0;
              w0.x = pack_half2(__uint_as_float(rb[8 * q]) * sc, __uint_as_float(rb[8 * q + 1]) * sc);
              w0.y = pack_half2(__uint_as_float(rb[8 * q + 2]) * sc, __uint_as_float(rb[8 * q + 3]) * sc);
              w0.z = pack_half2(__uint_as_float(rb[8 * q + 4]) * sc, __uint_as_float(rb[8 * q + 5]) * sc);
              w0.w = pack_half2(__uint_as_float(rb[8 * q + 6]) * sc, __uint_as_float(rb[8 * q + 7]) * sc);
              *reinterpret_cast<uint4*>(stage + 64 * hb + 16 * q) = w0;
            }
          }
          uint32_t ra[32];
          tmem_ld32(tmem_base + lane_addr + TM_AC + 32 * set, ra);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_free);            // this warp's part of S is in registers / shared memory
          const int u0 = 64 * b + 32 * set;
          const bool edge = (u0 < ulo) || (u0 + 32 > uhi);
          SkewWords sk;
          sk.load(stage, lane);
          if (__any_sync(0xffffffffu, edge)) {
#pragma unroll
            for (int k = 0; k < 32; k += 2) {
              const float2 bd = sk.pair(k >> 1);
              float v0 = PRE ? __uint_as_float(ra[k]) + bd.x : fmaf(__uint_as_float(ra[k]), p.scale_log2e, bd.x);
              float v1 = PRE ? __uint_as_float(ra[k + 1]) + bd.y : fmaf(__uint_as_float(ra[k + 1]), p.scale_log2e, bd.y);
              v0 = (u0 + k >= ulo && u0 + k < uhi) ? v0 : -INFINITY;
              v1 = (u0 + k + 1 >= ulo && u0 + k + 1 < uhi) ? v1 : -INFINITY;
              s[k] = v0; s[k + 1] = v1;
              mx = fmaxf(mx, fmaxf(v0, v1));
            }
          } else {
#pragma unroll
            for (int k = 0; k < 32; k += 2) {
              const float2 bd = sk.pair(k >> 1);
              const float v0 = PRE ? __uint_as_float(ra[k]) + bd.x : fmaf(__uint_as_float(ra[k]), p.scale_log2e, bd.x);
              const float v1 = PRE ? __uint_as_float(ra[k + 1]) + bd.y : fmaf(__uint_as_float(ra[k + 1]), p.scale_log2e, bd.y);
              s[k] = v0; s[k + 1] = v1;
              mx = fmaxf(mx, fmaxf(v0, v1));
            }
          }
        }
        float* xch = s_xch + (blk & 1) * 256;
        xch[set * 128 + rho] = mx;
        named_bar_sync(1 + quad, 64);      // only the two warps that share this quadrant's rows exchange anything
        const float m_new = fmaxf(m_run, fmaxf(mx, xch[(set ^ 1) * 128 + rho]));
        const float alpha = fast_exp2(m_run - m_new);
        float sum = 0.f;
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float p0 = fast_exp2(s[2 * e] - m_new), p1 = fast_exp2(s[2 * e + 1] - m_new);
          sum += p0 + p1;
          pk[e] = pack_bf16(p0, p1);
        }
        l_run = l_run * alpha + sum;
        m_run = m_new;
        if (blk > 0) mbar_wait(pv_done, (blk - 1) & 1);   // previous P V retired: the P columns and O are ours again
        tc_fence_after();
        if (ep_pending) write_out();                     // previous item (its row sums were published before the barrier above)
        tmem_st16(tmem_base + lane_addr + TM_P + 16 * set, pk);
        if (b > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {   // rescale this thread's columns of the running output
#pragma unroll
          for (int hc = 0; hc < OC / 32; ++hc) {
            uint32_t r[32];
            tmem_ld32(tmem_base + lane_addr + TM_O + OC * set + 32 * hc, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) * alpha);
            tmem_st32(tmem_base + lane_addr + TM_O + OC * set + 32 * hc, r);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
      }
      s_lx[set * 128 + rho] = l_run;                     // read by the partner thread after the next named barrier
      ep_pending = true; ep_l = l_run; ep_row = real ? (long long)tl.qrow0 + rho : -1;
    }
    if (ep_pending) {
      named_bar_sync(1 + quad, 64);      // only the two warps that share this quadrant's rows exchange anything
      mbar_wait(pv_done, (blk - 1) & 1);
      tc_fence_after();
      write_out();
      tc_fence_before();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <int DK>
inline bool launch_attention_ring_dk(const AttnParams& a, cudaStream_t st, std::string* err) {
  const int W = a.l + a.c + a.r;
  const int R = 2 * a.c + a.l + a.r - 1;
  const int Rpad = ((R + 127) / 128) * 128;
  AttnRingParams p{};
  p.range = a.range; p.ctx = a.ctx; p.n_chunks = a.n_chunks; p.c = a.c; p.l = a.l; p.d = a.d; p.heads = a.heads;
  p.scale_log2e = a.scale * 1.4426950408889634f;
  int U;
  if (a.c < 128) {
    int c_log2 = 0;
    while ((1 << c_log2) < a.c) ++c_log2;
    p.c_log2 = c_log2; p.tiles_per_chunk = 1;
    p.n_tiles = (a.n_chunks + (128 / a.c) - 1) / (128 / a.c);
    U = a.l + 128 + a.r;
  } else {
    p.c_log2 = -1; p.tiles_per_chunk = (a.c + 127) / 128;
    p.n_tiles = a.n_chunks * p.tiles_per_chunk;
    U = W;
  }
  p.nb = (U + 63) / 64;
  p.last_chunks = (W - 64 * (p.nb - 1) <= 1) ? 2 : 3;   // no valid score of the last block needs S_bd columns >= 128
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int per_head = sms / a.heads;
  if (per_head < 1) per_head = 1;
  if (per_head > p.n_tiles) per_head = p.n_tiles;
  p.items_per_cta_stride = per_head;
  const uint64_t qkv_rows = uint64_t(a.l) + uint64_t(a.n_chunks) * a.c + uint64_t(a.r) + 2 * uint64_t(a.c) + 128;   // = the buffer the caller allocates
  CUtensorMap tq, tk, tp;
  if (!make_tma_2d_bf16(&tq, a.qkv, qkv_rows, uint64_t(4) * a.d, uint64_t(4) * a.d, 128, 64, err)) return false;
  if (!make_tma_2d_bf16(&tk, a.qkv, qkv_rows, uint64_t(4) * a.d, uint64_t(4) * a.d, 64, 64, err)) return false;
  if (!make_tma_2d_bf16(&tp, a.pos, uint64_t(Rpad), uint64_t(a.d), uint64_t(a.d), 64, 64, err)) return false;
  const size_t smem = atcr_smem_bytes<DK>();
  if (!ensure_smem_optin(attention_ring_kernel<DK, true>, smem, err, "attention_ring")) return false;
  if (!ensure_smem_optin(attention_ring_kernel<DK, false>, smem, err, "attention_ring")) return false;
  if (a.prescaled) attention_ring_kernel<DK, true><<<per_head * a.heads, ATCR_THREADS, smem, st>>>(tq, tk, tp, p);
  else attention_ring_kernel<DK, false><<<per_head * a.heads, ATCR_THREADS, smem, st>>>(tq, tk, tp, p);
  ++g_kernel_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { if (err) *err = std::string("attention_ring launch: ") + cudaGetErrorString(e); return false; }
  return true;
}
inline bool launch_attention_ring(const AttnParams& a, cudaStream_t st, std::string* err) {
  return (a.d / a.heads) == 128 ? launch_attention_ring_dk<128>(a, st, err) : launch_attention_ring_dk<64>(a, st, err);
}

inline bool launch_attention_tc(const AttnParams& a, cudaStream_t st, std::string* err) {
  // tile = 128 query rows = 128 / c consecutive chunks; their union key window has l + 128 + r slots
  const int W = a.l + a.c + a.r;
  const int U = a.l + 128 + a.r;
  const int R = 2 * a.c + a.l + a.r - 1;
  const int Rpad = ((R + 127) / 128) * 128;
  int c_log2 = 0;
  while ((1 << c_log2) < a.c) ++c_log2;
  const int cpt = 128 / a.c;
  AttnTcParams p{};
  p.range = a.range; p.ctx = a.ctx; p.n_chunks = a.n_chunks; p.n_pairs = (a.n_chunks + cpt - 1) / cpt; p.l = a.l; p.d = a.d;
  p.heads = a.heads; p.nb = (U + 127) / 128; p.scale_log2e = a.scale * 1.4426950408889634f;
  p.c_log2 = c_log2; p.tab_row0 = a.c - 128;
  p.n_last = (W - 128 * (p.nb - 1) <= 65) ? 192 : 256;   // widest S_bd column a valid score of the last block can need
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int per_head = sms / a.heads;
  if (per_head < 1) per_head = 1;
  if (per_head > p.n_pairs) per_head = p.n_pairs;
  p.items_per_cta_stride = per_head;
  // tensor maps: flat QKV buffer [rows, 4d] (box 128 rows x 64 cols) and this layer's position table [Rpad, d] (box 64 x 64)
  const uint64_t qkv_rows = uint64_t(a.l) + uint64_t(a.n_chunks) * a.c + uint64_t(a.r) + 2 * a.c + 128;
  CUtensorMap tq, tp;
  if (!make_tma_2d_bf16(&tq, a.qkv, qkv_rows, uint64_t(4) * a.d, uint64_t(4) * a.d, 128, 64, err)) return false;
  if (!make_tma_2d_bf16(&tp, a.pos, uint64_t(Rpad), uint64_t(a.d), uint64_t(a.d), 64, 64, err)) return false;
  if (!ensure_smem_optin(attention_tc_kernel<true>, ATC_SMEM_BYTES, err, "attention_tc")) return false;
  if (!ensure_smem_optin(attention_tc_kernel<false>, ATC_SMEM_BYTES, err, "attention_tc")) return false;
#ifdef CF_ABLATION
  const int prof_ctas = per_head * a.heads;
  { const char* e = getenv("CF_ATTN_DEBUG"); p.debug = e ? atoi(e) : 0; }
  if (getenv("CF_ATTN_PROF")) { cudaMalloc(&p.prof, size_t(prof_ctas) * 32 * 8); cudaMemsetAsync(p.prof, 0, size_t(prof_ctas) * 32 * 8, st); }
#endif
  if (a.prescaled) attention_tc_kernel<true><<<per_head * a.heads, ATC_THREADS, ATC_SMEM_BYTES, st>>>(tq, tp, p);
  else attention_tc_kernel<false><<<per_head * a.heads, ATC_THREADS, ATC_SMEM_BYTES, st>>>(tq, tp, p);
#ifdef CF_ABLATION
  if (p.prof) {
    std::vector<long long> hp(size_t(prof_ctas) * 32);
    cudaStreamSynchronize(st);
    cudaMemcpy(hp.data(), p.prof, hp.size() * 8, cudaMemcpyDeviceToHost);
    cudaFree(p.prof);
    static const char* names[10] = {"wait S (MMA)", "TMEM loads, staging, skew add, max", "s_free + max exchange barrier", "exponentials",
                                    "wait previous P V", "write-out of previous item + P store", "O rescale + p_full", "loop overhead",
                                    "total", "blocks"};
    for (int w = 0; w < 2; ++w) {
      fprintf(stderr, "attention_tc softmax warp %d (cycles per key block, mean over %d CTAs):\n", w ? 7 : 2, prof_ctas);
      double blocks = 0;
      for (int c = 0; c < prof_ctas; ++c) blocks += double(hp[(size_t(c) * 2 + w) * 16 + 9]);
      for (int f = 0; f < 9; ++f) {
        double sum = 0;
        for (int c = 0; c < prof_ctas; ++c) sum += double(hp[(size_t(c) * 2 + w) * 16 + f]);
        fprintf(stderr, "  %-40s %10.0f\n", names[f], blocks > 0 ? sum / blocks : 0.0);
      }
    }
  }
#endif
  ++g_kernel_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { if (err) *err = std::string("attention_tc launch: ") + cudaGetErrorString(e); return false; }
  return true;
}

}  // namespace cf
