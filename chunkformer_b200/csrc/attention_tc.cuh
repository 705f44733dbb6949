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
//     P = exp2(s*scale*log2e - m)  (bf16, written to smem as the K-major A operand),  O += P V_blk  UMMA 128x64x128
// The row-dependent skew goes through a thread-private shared-memory row (each thread owns one query row after
// tcgen05.ld 32x32b), so no cross-thread synchronisation is needed for it.  The score matrix never reaches HBM.
//
//   warp 0 : TMA producer (Q tiles per item, K/V blocks in a 2-stage ring, the head's position table once)
//   warp 1 : TMEM allocator + single-thread MMA issuer (S of block j+1 is issued before waiting for P of block j)
//   warps 2..9 : softmax / correction / output; two threads share a query row (64 of each block's 128 keys each),
//                exchanging row maxima / sums through shared memory; fp16 staging of the pre-scaled S_bd values
#pragma once
#include <cuda_fp16.h>

#include <string>

#include "attention_simt.cuh"
#include "gemm_host.cuh"

namespace cf {

constexpr bool kAttentionTcReady = true;

constexpr int ATC_THREADS = 320;                 // TMA warp, MMA warp, 8 softmax warps
constexpr int ATC_STAGE_PITCH = 144;             // bytes per thread-private skew row (64 fp16 + pad; 16-byte stores conflict free)
constexpr uint32_t ATC_PTAB_ROWS = 448;          // table rows -64 .. 383 of the head
constexpr uint32_t ATC_PTAB_BYTES = ATC_PTAB_ROWS * 128;
constexpr uint32_t ATC_TILE_BYTES = 128 * 128;   // 128 rows x 64 bf16
constexpr size_t ATC_SMEM_BYTES = ATC_PTAB_BYTES + 2 * ATC_TILE_BYTES /*Qu,Qv*/ + 4 * ATC_TILE_BYTES /*K,V x2*/ +
                                  2 * ATC_TILE_BYTES /*P probs*/ + 256 * ATC_STAGE_PITCH + 2048 + 1024 + 1024 + 128;

CF_DEVINL float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
CF_DEVINL uint32_t pack_half2(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

struct AttnTcParams {
  const int2* range;      // [n_chunks + 2] valid key slots per chunk (entries beyond n_chunks are empty)
  __nv_bfloat16* ctx;     // [n_chunks * 64, d]
  int n_chunks, n_pairs, l, d, heads, nb;   // nb = key blocks of 128 union slots
  int items_per_cta_stride;                  // CTAs per head
  float scale_log2e;
};

// PRE: Q+u / Q+v already carry (1/sqrt(d_k)) * log2(e) (folded into the fused QKV projection at weight load).
template <bool PRE>
__global__ void __launch_bounds__(ATC_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tma_qkv, const __grid_constant__ CUtensorMap tma_pos, AttnTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
  uint8_t* s_ptab = smem;                        // smem row rr <-> table row rr - 64
  uint8_t* s_qu = s_ptab + ATC_PTAB_BYTES;
  uint8_t* s_qv = s_qu + ATC_TILE_BYTES;
  uint8_t* s_k = s_qv + ATC_TILE_BYTES;          // [2]
  uint8_t* s_v = s_k + 2 * ATC_TILE_BYTES;       // [2]
  uint8_t* s_pp = s_v + 2 * ATC_TILE_BYTES;      // 2 atoms of 64 keys
  uint8_t* s_stage = s_pp + 2 * ATC_TILE_BYTES;  // 256 thread-private rows
  float* s_xch = reinterpret_cast<float*>(s_stage + 256 * ATC_STAGE_PITCH);   // [2 parity][2 set][128] row maxima
  float* s_lx = s_xch + 512;                                                   // [2 set][128] row sums
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_lx + 256);
  uint64_t* ptab_full = bars + 0;
  uint64_t* q_full = bars + 1;
  uint64_t* q_empty = bars + 2;
  uint64_t* kv_full = bars + 3;    // [2]
  uint64_t* kv_empty = bars + 5;   // [2]
  uint64_t* s_full = bars + 7;
  uint64_t* s_free = bars + 8;
  uint64_t* p_full = bars + 9;
  uint64_t* pv_done = bars + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
    for (int s = 0; s < 2; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    mbar_init(s_full, 1);
    mbar_init(s_free, 256);
    mbar_init(p_full, 256);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t TM_AC = 0, TM_BD = 128, TM_O = 384;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(ptab_full, ATC_PTAB_BYTES);
      for (int i = 0; i < 7; ++i)                                              // table rows -64 .. 383 (rows < 0 read as 0)
        tma_load_2d(s_ptab + i * 64 * 128, &tma_pos, ptab_full, h * 64, -64 + 64 * i);
      uint32_t item = 0, blk = 0;
      for (int pair = first_pair; pair < p.n_pairs; pair += p.items_per_cta_stride, ++item) {
        const int g0 = 2 * pair;
        mbar_wait(q_empty, (item & 1) ^ 1);
        mbar_arrive_expect_tx(q_full, 2 * ATC_TILE_BYTES);
        tma_load_2d(s_qu, &tma_qkv, q_full, h * 64, p.l + 64 * g0);
        tma_load_2d(s_qv, &tma_qkv, q_full, d + h * 64, p.l + 64 * g0);
        for (int b = 0; b < nb; ++b, ++blk) {
          const uint32_t st = blk & 1, ph = (blk >> 1) & 1;
          mbar_wait(&kv_empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&kv_full[st], 2 * ATC_TILE_BYTES);
          tma_load_2d(s_k + st * ATC_TILE_BYTES, &tma_qkv, &kv_full[st], 2 * d + h * 64, 64 * g0 + 128 * b);
          tma_load_2d(s_v + st * ATC_TILE_BYTES, &tma_qkv, &kv_full[st], 3 * d + h * 64, 64 * g0 + 128 * b);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_ac = make_idesc_bf16(128, 128);
      constexpr uint32_t idesc_bd = make_idesc_bf16(128, 256);
      constexpr uint32_t idesc_bd_last = make_idesc_bf16(128, 192);   // the last block never needs table rows >= 384
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);   // B (= V) is MN-major
      mbar_wait(ptab_full, 0);
      uint32_t item = 0, blk = 0;
      bool have_prev = false;
      uint32_t prev_st = 0, prev_b = 0;
      auto issue_pv = [&](uint32_t pblk, uint32_t st, uint32_t b) {
        mbar_wait(p_full, pblk & 1);
        tc_fence_after();
        const uint64_t da = make_sw128_desc(smem_u32(s_pp));
        const uint64_t db = make_sw128_desc(smem_u32(s_v + st * ATC_TILE_BYTES));
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          // A: 16 keys = 32 B inside a 64-key atom, atoms 16 KB apart; B (MN-major): 16 key rows = 2048 B
          umma_bf16_ss(tmem_base + TM_O, da + uint64_t(((t >> 2) * ATC_TILE_BYTES + (t & 3) * 32) >> 4),
                       db + uint64_t((t * 2048) >> 4), idesc_pv, (b | t) != 0);
        }
        umma_commit(pv_done);
        umma_commit(&kv_empty[st]);
      };
      for (int pair = first_pair; pair < p.n_pairs; pair += p.items_per_cta_stride, ++item) {
        mbar_wait(q_full, item & 1);
        for (int b = 0; b < nb; ++b, ++blk) {
          const uint32_t st = blk & 1, ph = (blk >> 1) & 1;
          mbar_wait(&kv_full[st], ph);
          mbar_wait(s_free, (blk & 1) ^ 1);            // softmax finished reading the previous S block
          tc_fence_after();
          const uint64_t dqu = make_sw128_desc(smem_u32(s_qu)), dqv = make_sw128_desc(smem_u32(s_qv));
          const uint64_t dk = make_sw128_desc(smem_u32(s_k + st * ATC_TILE_BYTES));
          const uint64_t dp = make_sw128_desc(smem_u32(s_ptab + b * 128 * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base + TM_AC, dqu + 2 * k, dk + 2 * k, idesc_ac, k != 0);
          const uint32_t idbd = (b == nb - 1) ? idesc_bd_last : idesc_bd;
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base + TM_BD, dqv + 2 * k, dp + 2 * k, idbd, k != 0);
          umma_commit(s_full);
          if (b == nb - 1) umma_commit(q_empty);       // Q tiles may be overwritten once these MMAs retire
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
    const int half = rho >> 6, qi = rho & 63;
    const uint32_t lane_addr = uint32_t(quad * 32) << 16;
    uint8_t* stage = s_stage + (threadIdx.x - 64) * ATC_STAGE_PITCH;
    const __half* stage_rd = reinterpret_cast<const __half*>(stage) + (31 - lane);
    const int cb_thread = 96 - 32 * quad + 64 * set;   // first S_bd column this warp stages (warp-uniform)
    uint8_t* pp_row = s_pp + set * ATC_TILE_BYTES + rho * 128;
    uint32_t blk = 0;
    for (int pair = first_pair; pair < p.n_pairs; pair += p.items_per_cta_stride) {
      const int g = 2 * pair + half;
      const int2 rg = p.range[g];
      const int ulo = rg.x + 64 * half, uhi = rg.y + 64 * half;   // valid union slots for this row
      float m_run = -1e30f, l_run = 0.f;
      for (int b = 0; b < nb; ++b, ++blk) {
        mbar_wait(s_full, blk & 1);
        tc_fence_after();
        float s[64];
        float mx = -1e30f;
        uint4 keep[4];                                 // packed fp16 S_bd columns [32, 64) of this thread's window:
                                                       // staged as the upper half for sub-block 0, lower half for sub-block 1
#pragma unroll
        for (int sb = 0; sb < 2; ++sb) {
          uint32_t r0[32];
          const uint32_t cbase = TM_BD + cb_thread;
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
          if (__any_sync(0xffffffffu, edge)) {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              const int uq = u0 + k;
              float v = PRE ? __uint_as_float(r0[k]) + __half2float(stage_rd[k]) : fmaf(__uint_as_float(r0[k]), p.scale_log2e, __half2float(stage_rd[k]));
              v = (uq >= ulo && uq < uhi) ? v : -INFINITY;
              s[32 * sb + k] = v;
              mx = fmaxf(mx, v);
            }
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              const float v = PRE ? __uint_as_float(r0[k]) + __half2float(stage_rd[k]) : fmaf(__uint_as_float(r0[k]), p.scale_log2e, __half2float(stage_rd[k]));
              s[32 * sb + k] = v;
              mx = fmaxf(mx, v);
            }
          }
        }
        tc_fence_before();
        mbar_arrive(s_free);                           // this thread's part of S is in registers
        float* xch = s_xch + (blk & 1) * 256;
        xch[set * 128 + rho] = mx;
        named_bar_sync(1, 256);
        const float m_new = fmaxf(m_run, fmaxf(mx, xch[(set ^ 1) * 128 + rho]));
        const float alpha = fast_exp2(m_run - m_new);
        float sum = 0.f;
        if (blk > 0) mbar_wait(pv_done, (blk - 1) & 1);   // previous P V retired: P tile and O are ours again
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < 8; ++j) {                  // 8 x (8 keys -> 16 bytes) into this set's 64-key K-major atom
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float p0 = fast_exp2(s[8 * j + 2 * e] - m_new), p1 = fast_exp2(s[8 * j + 2 * e + 1] - m_new);
            sum += p0 + p1;
            w[e] = pack_bf16(p0, p1);
          }
          *reinterpret_cast<uint4*>(pp_row + ((j ^ (rho & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        l_run = l_run * alpha + sum;
        m_run = m_new;
        if (b > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {   // rescale this set's half of the running output
          uint32_t r[32];
          tmem_ld32(tmem_base + lane_addr + TM_O + 32 * set, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) * alpha);
          tmem_st32(tmem_base + lane_addr + TM_O + 32 * set, r);
          tmem_st_wait();
        }
        tc_fence_before();
        fence_proxy_async();                           // P tile (generic-proxy stores) -> visible to the MMA (async proxy)
        mbar_arrive(p_full);
      }
      // ---- item epilogue: O / l -> ctx (this set writes 32 of the head's 64 output columns)
      s_lx[set * 128 + rho] = l_run;
      named_bar_sync(1, 256);
      const float l_tot = l_run + s_lx[(set ^ 1) * 128 + rho];
      mbar_wait(pv_done, (blk - 1) & 1);
      tc_fence_after();
      const float inv = l_tot > 0.f ? 1.0f / l_tot : 0.f;        // no valid key: zero context (attention.py:133-136)
      __nv_bfloat16* orow = p.ctx + ((long long)g * 64 + qi) * d + h * 64 + 32 * set;
      {
        uint32_t r[32];
        tmem_ld32(tmem_base + lane_addr + TM_O + 32 * set, r);
        tmem_ld_wait();
        if (g < p.n_chunks) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 o;
            o.x = pack_bf16(__uint_as_float(r[8 * q]) * inv, __uint_as_float(r[8 * q + 1]) * inv);
            o.y = pack_bf16(__uint_as_float(r[8 * q + 2]) * inv, __uint_as_float(r[8 * q + 3]) * inv);
            o.z = pack_bf16(__uint_as_float(r[8 * q + 4]) * inv, __uint_as_float(r[8 * q + 5]) * inv);
            o.w = pack_bf16(__uint_as_float(r[8 * q + 6]) * inv, __uint_as_float(r[8 * q + 7]) * inv);
            *reinterpret_cast<uint4*>(orow + 8 * q) = o;
          }
        }
      }
      tc_fence_before();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

inline bool launch_attention_tc(const AttnParams& a, cudaStream_t st, std::string* err) {
  const int W = a.l + a.c + a.r;
  const int U = W + 64;
  const int R = 2 * a.c + a.l + a.r - 1;
  const int Rpad = ((R + 127) / 128) * 128;
  AttnTcParams p{};
  p.range = a.range; p.ctx = a.ctx; p.n_chunks = a.n_chunks; p.n_pairs = (a.n_chunks + 1) / 2; p.l = a.l; p.d = a.d;
  p.heads = a.heads; p.nb = (U + 127) / 128; p.scale_log2e = a.scale * 1.4426950408889634f;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int per_head = sms / a.heads;
  if (per_head < 1) per_head = 1;
  if (per_head > p.n_pairs) per_head = p.n_pairs;
  p.items_per_cta_stride = per_head;
  // tensor maps: flat QKV buffer [rows, 4d] (box 128 rows x 64 cols) and this layer's position table [Rpad, d] (box 64 x 64)
  const uint64_t qkv_rows = uint64_t(a.l) + uint64_t(a.n_chunks) * 64 + uint64_t(a.r) + 2 * 64 + 128;
  CUtensorMap tq, tp;
  if (!make_tma_2d_bf16(&tq, a.qkv, qkv_rows, uint64_t(4) * a.d, uint64_t(4) * a.d, 128, 64, err)) return false;
  if (!make_tma_2d_bf16(&tp, a.pos, uint64_t(Rpad), uint64_t(a.d), uint64_t(a.d), 64, 64, err)) return false;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(ATC_SMEM_BYTES));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(ATC_SMEM_BYTES));
    if (e != cudaSuccess) { if (err) *err = std::string("cudaFuncSetAttribute(attention_tc): ") + cudaGetErrorString(e); return false; }
    attr_set = true;
  }
  if (a.prescaled) attention_tc_kernel<true><<<per_head * a.heads, ATC_THREADS, ATC_SMEM_BYTES, st>>>(tq, tp, p);
  else attention_tc_kernel<false><<<per_head * a.heads, ATC_THREADS, ATC_SMEM_BYTES, st>>>(tq, tp, p);
  ++g_kernel_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { if (err) *err = std::string("attention_tc launch: ") + cudaGetErrorString(e); return false; }
  return true;
}

}  // namespace cf
