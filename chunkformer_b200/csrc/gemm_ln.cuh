// Residual GEMM with the following LayerNorm(s) fused into the epilogue, on a pair of CTAs (thread-block cluster of 2):
//
//     x_new = resid + rowmask * alpha * (A W^T + bias)                         (encoder_layer.py:190-246, every "x = x + ..." line)
//     LNM_Y     : x <- x_new,             y = LN1(x_new)                        (norm_mha / norm_conv / norm_ff, first norm_ff_macaron)
//     LNM_XY    : x <- LN1(x_new),        y = LN2(LN1(x_new))                   (norm_final of layer i + norm_ff_macaron of layer i + 1)
//     LNM_FINAL : out = LN2(LN1(x_new))   fp32 and / or bf16                    (norm_final of the last layer + after_norm, encoder.py:670-671)
//
// A LayerNorm row spans all N = d_model output columns, so one cluster owns a 128-row block: CTA r of the pair computes
// columns [r * NC, r * NC + NC), NC = N / 2 (UMMA 128 x NC x 16, cta_group::1, accumulators double-buffered in TMEM across row
// blocks), and the two CTAs exchange per-row partial statistics (count, mean, M2: Chan's parallel variance, no E[x^2] - mean^2
// cancellation) through distributed shared memory: each epilogue thread (one row, NC / 2 columns) writes its partial into its
// own CTA's table and the peer's (st.shared::cluster) and arrives on both CTAs' mbarrier; after the wait every thread merges
// the four partials of its row in a fixed order, so both CTAs normalise with bit-identical mean / rstd.
//
// x_new never makes a round trip through memory before it is normalised: pass 1 writes it to global memory (TMA store) AND
// back into the accumulator's TMEM columns (tcgen05.st), pass 2 re-reads it from TMEM.  The stand-alone LayerNorm kernels
// (one HBM read of x and one write of y per LayerNorm: 10 % of the round-1 step at 100 % of HBM bandwidth) disappear.
//
//   warp 0 : TMA producer (A 128 x 64 and W NC x 64 per stage)      warp 1 : TMEM allocator + MMA issuer
//   warps 2..9 : epilogue, two groups of four warps (one per TMEM lane quadrant), group g owns NC / 2 columns of the CTA's tile;
//                residual sub-tiles (128 rows x 32 fp32) arrive by TMA in the group's two 16 KB staging slots, one round ahead
#pragma once
#include "gemm.cuh"

namespace cf {

enum GemmLnMode : int { LNM_Y = 1, LNM_XY = 2, LNM_FINAL = 3 };

struct GemmLnParams {
  const float* bias = nullptr;         // [N]
  int has_resid = 0;                   // residual sub-tiles are read through tma_r
  float alpha = 1.0f;
  const int2* row_range = nullptr;     // row valid iff range[row / rows_per_chunk].x <= row % rows_per_chunk < .y (convolution.py:253)
  int rows_per_chunk = 1;
  int mode = LNM_Y;
  const float* ln1_w = nullptr; const float* ln1_b = nullptr;
  const float* ln2_w = nullptr; const float* ln2_b = nullptr;
  const int* row_limit = nullptr;      // LNM_Y: rows with (row % rows_per_seq) >= limit[row / rows_per_seq] give y = 0
  int rows_per_seq = 1;
  int store_f32 = 1, store_bf16 = 1;   // LNM_FINAL: which outputs exist
  // Gathered A (compact streaming, api.cu; gemm_ln_split_kernel / gemm_ln_quad_kernel; gather_rows > 0, tma_a is then a 3-D map
  // {K, rows, groups}): row m of the GEMM is row gather_row0 + m % gather_rows of group m / gather_rows (gather_rows divides 128).
  int gather_rows = 0, gather_row0 = 0;
#ifdef CF_ABLATION
  long long* prof = nullptr;           // tools build: [grid][16] cycles per role spent waiting / working (gemm_ln_split_kernel)
  int debug = 0;                       // tools build: phase ablation bits (gemm_ln_split_kernel; results are wrong)
#endif
};

#ifdef CF_ABLATION
#define CF_PROF_T0() const long long _pt0 = clock64()
#define CF_PROF_ADD(var) (var) += clock64() - _pt0
#define CF_PROF_DECL(...) long long __VA_ARGS__
#define CF_LN_DBG(ep, bit) (((ep).debug & (bit)) != 0)
#define CF_PROF_MARK(k) do { const long long _n = clock64(); seg[k] += _n - seg_t; seg_t = _n; } while (0)
#else
#define CF_LN_DBG(ep, bit) (false)
#define CF_PROF_MARK(k) do {} while (0)
#define CF_PROF_T0() do {} while (0)
#define CF_PROF_ADD(var) do {} while (0)
#define CF_PROF_DECL(...) do {} while (0)
#endif

template <int NC> __host__ __device__ constexpr int gemmln_stages() { return NC == 256 ? 3 : 4; }
template <int NC> constexpr size_t gemmln_smem_bytes() {
  return size_t(gemmln_stages<NC>()) * (GEMM_BM * 128 + NC * 128) + 4 * GEMM_STAGING_BYTES + 2 * 4 * 128 * sizeof(float2) + 1024 + 256;
}

CF_DEVINL void st_cluster_f32x2(uint32_t cluster_addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(cluster_addr), "f"(a), "f"(b) : "memory");
}
// Bounded wait with cluster-scope acquire (the releases come from the peer CTA's threads).
CF_DEVINL void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}

// Running (count, mean, M2) of a row segment; `add32` folds in 32 values with a two-pass chunk variance, `merge` is Chan's rule.
struct RowStats {
  float n = 0.f, mean = 0.f, m2 = 0.f;
  CF_DEVINL void merge(float nb, float mb, float m2b) {
    const float nt = n + nb;
    const float delta = mb - mean;
    const float f = nb / nt;
    mean = fmaf(delta, f, mean);
    m2 = m2 + m2b + delta * delta * n * f;
    n = nt;
  }
  CF_DEVINL void add32(const float (&v)[32]) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) s += v[j];
    const float mc = s * (1.0f / 32.0f);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) { const float dlt = v[j] - mc; q = fmaf(dlt, dlt, q); }
    if (n == 0.f) { n = 32.f; mean = mc; m2 = q; } else merge(32.f, mc, q);
  }
};

// ---------------------------------------------------------------------------------------------------------------------
// The LayerNorm epilogue of one group of four warps (128 threads = 128 rows, one TMEM lane quadrant per warp) over NCG
// accumulator columns: shared by gemm_ln_kernel (two groups per CTA, NCG = NC / 2) and ffn_fused_kernel (one group per CTA,
// NCG = NC).  The group owns two 16 KB staging slots used alternately (`q` counts the uses): residual sub-tiles (128 rows x
// 32 fp32) land in them by TMA one round ahead, results leave through them by TMA store.
// ---------------------------------------------------------------------------------------------------------------------
template <int NCG>
struct LnEpilogue {
  static constexpr int NR = NCG / 32;   // 32-column rounds per pass
  // wiring (constant for the kernel)
  const CUtensorMap* tma_x; const CUtensorMap* tma_r; const CUtensorMap* tma_y;
  const GemmLnParams* ep;
  uint8_t* stg;            // this group's two staging slots
  uint64_t* rfull;         // [2] residual-arrival barriers of the two slots
  uint64_t* stat_bar;      // [2] statistics barriers (n_part warps-worth of arrivals each: every epilogue warp of both CTAs)
  float2* s_stat;          // [2 buffers][n_part][128 rows] (mean, M2)
  int n_part;              // partial statistics per row = epilogue groups per CTA * 2
  int part_id;             // this group's slot in the table: rank * groups_per_cta + group
  uint32_t rank;
  int bar_id, trow, lane, M, n_total;
  bool issuer;
  // state
  uint32_t q = 0;                 // staging uses so far (slot = q & 1)
  uint32_t nl0 = 0, nl1 = 0;      // residual loads issued into each slot (mbarrier phase bookkeeping, identical in every thread)
  CF_DEVINL void count_load(uint32_t slot) { if (slot) ++nl1; else ++nl0; }
  CF_DEVINL uint32_t loads(uint32_t slot) const { return slot ? nl1 : nl0; }
  uint32_t xr = 0;                // statistics exchanges so far

  // a slot may be rewritten by the threads once the TMA store that last read it (two uses ago) has drained it
  CF_DEVINL void acquire_slot() {
    if (issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    named_bar_sync(bar_id, 128);
  }
  CF_DEVINL void issue_resid(uint32_t slot, int col0, int row0) {      // issuer only
    mbar_arrive_expect_tx(&rfull[slot], GEMM_STAGING_BYTES);
    tma_load_2d(stg + slot * GEMM_STAGING_BYTES, tma_r, &rfull[slot], col0, row0);
  }
  // residual sub-tile of the first round of the row block at row0 into the slot its first use will take
  CF_DEVINL void prefetch_first(int gcol0, int row0, bool after_stores) {
    if (!ep->has_resid) return;
    if (issuer) {
      if (after_stores) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      issue_resid(q & 1u, gcol0, row0);
    }
    count_load(q & 1u);
  }
  // per-row statistics over all columns of the row: this thread's partial + the other groups' of both CTAs.
  // Exchange k uses table / barrier k & 1: an arrival for exchange k + 2 can only be made by a warp that has passed exchange
  // k + 1, i.e. after every warp of both CTAs has arrived for (and therefore finished reading) exchange k.
  CF_DEVINL void exchange(const RowStats& mine, float& mean, float& rstd) {
    const uint32_t buf = xr & 1u;
    float2* p = s_stat + (buf * uint32_t(n_part) + uint32_t(part_id)) * 128u + trow;
    *p = make_float2(mine.mean, mine.m2);
    st_cluster_f32x2(mapa_rank(smem_u32(p), rank ^ 1u), mine.mean, mine.m2);
    __syncwarp();
    if (lane == 0) {
      mbar_arrive_cluster(mapa_rank(smem_u32(&stat_bar[buf]), rank ^ 1u));
      mbar_arrive_cluster(mapa_rank(smem_u32(&stat_bar[buf]), rank));
    }
    mbar_wait_cluster(&stat_bar[buf], (xr >> 1) & 1u);
    RowStats tot;
    for (int sl = 0; sl < n_part; ++sl) {
      const float2 v = s_stat[(buf * uint32_t(n_part) + sl) * 128u + trow];
      if (sl == 0) { tot.n = float(NCG); tot.mean = v.x; tot.m2 = v.y; } else tot.merge(float(NCG), v.x, v.y);
    }
    mean = tot.mean;
    rstd = rsqrtf(tot.m2 / float(n_total) + 1e-5f);
    ++xr;
  }

  // One 128-row block: the accumulator slab at `taddr` (this thread's TMEM lane, NCG fp32 columns), global columns
  // [gcol0, gcol0 + NCG), rows [row0, row0 + 128).  next_row0 >= 0: the row block this group handles next (residual prefetch).
  // The caller waits for the accumulator before and releases it after.
  CF_DEVINL void tile(uint32_t taddr, int row0, int gcol0, int next_row0) {
    const GemmLnParams& e = *ep;
    const bool has_res = e.has_resid != 0;
    const int row = row0 + trow;
    bool keep = true;
    if (e.row_range != nullptr && row < M) {
      const int ch = row / e.rows_per_chunk;
      const int rr = row - ch * e.rows_per_chunk;
      const int2 rg = e.row_range[ch];
      keep = (rr >= rg.x && rr < rg.y);
    }
    // the residual slab of the row block this group handles next -> L2 now, so that its sub-tile loads (issued only one round
    // ahead, there are two staging slots) are L2 hits instead of HBM round trips
    if (issuer && has_res && next_row0 >= 0) {
#pragma unroll 1
      for (int cc = 0; cc < NR; ++cc) tma_prefetch_2d(tma_r, gcol0 + cc * 32, next_row0);
    }
    // ---------------- pass 1: x_new = resid + keep * alpha * (acc + bias) -> TMEM (+ global for LNM_Y), statistics
    RowStats st1;
#pragma unroll 1
    for (int cc = 0; cc < NR; ++cc, ++q) {
      const int col0 = gcol0 + cc * 32;
      const uint32_t slot = q & 1u;
      uint8_t* tl = stg + slot * GEMM_STAGING_BYTES;
      uint32_t r[32];
      tmem_ld32(taddr + cc * 32, r);
      if (has_res) {
        if (cc + 1 < NR) {                       // next round's residual into the other slot
          if (issuer) { tma_store_wait_read(); issue_resid(slot ^ 1u, col0 + 32, row0); }
          count_load(slot ^ 1u);
        }
      } else if (e.mode == LNM_Y) {
        acquire_slot();
      }
      float4 b[8];
#pragma unroll
      for (int qd = 0; qd < 8; ++qd) b[qd] = __ldg(reinterpret_cast<const float4*>(e.bias + col0) + qd);
      if (has_res) mbar_wait(&rfull[slot], (loads(slot) - 1u) & 1u);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int qd = 0; qd < 8; ++qd) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has_res) {
          const uint4 xr4 = stage_load16(tl, trow, qd);
          x = make_float4(__uint_as_float(xr4.x), __uint_as_float(xr4.y), __uint_as_float(xr4.z), __uint_as_float(xr4.w));
        }
        v[4 * qd] = keep ? fmaf(e.alpha, __uint_as_float(r[4 * qd]) + b[qd].x, x.x) : x.x;
        v[4 * qd + 1] = keep ? fmaf(e.alpha, __uint_as_float(r[4 * qd + 1]) + b[qd].y, x.y) : x.y;
        v[4 * qd + 2] = keep ? fmaf(e.alpha, __uint_as_float(r[4 * qd + 2]) + b[qd].z, x.z) : x.z;
        v[4 * qd + 3] = keep ? fmaf(e.alpha, __uint_as_float(r[4 * qd + 3]) + b[qd].w, x.w) : x.w;
        if (e.mode == LNM_Y)
          stage_store16(tl, trow, qd, make_uint4(__float_as_uint(v[4 * qd]), __float_as_uint(v[4 * qd + 1]),
                                                 __float_as_uint(v[4 * qd + 2]), __float_as_uint(v[4 * qd + 3])));
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(v[j]);
      tmem_st32(taddr + cc * 32, r);
      st1.add32(v);
      if (e.mode == LNM_Y) fence_proxy_async();
      named_bar_sync(bar_id, 128);               // every thread is done with this slot (and with the one reloaded next round)
      if (e.mode == LNM_Y && issuer) { tma_store_2d(tma_x, tl, col0, row0); tma_store_commit(); }
    }
    tmem_st_wait();
    float mean1, rstd1;
    exchange(st1, mean1, rstd1);

    // ---------------- LNM_XY / LNM_FINAL: x_mid = LN1(x_new) -> TMEM (+ global for LNM_XY), statistics of x_mid
    float mean2 = 0.f, rstd2 = 1.f;
    if (e.mode != LNM_Y) {
      RowStats st2;
#pragma unroll 1
      for (int cc = 0; cc < NR; ++cc) {
        const int col0 = gcol0 + cc * 32;
        uint32_t r[32];
        tmem_ld32(taddr + cc * 32, r);
        uint8_t* tl = stg + (q & 1u) * GEMM_STAGING_BYTES;
        if (e.mode == LNM_XY) acquire_slot();
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int qd = 0; qd < 8; ++qd) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(e.ln1_w + col0) + qd);
          const float4 bb = __ldg(reinterpret_cast<const float4*>(e.ln1_b + col0) + qd);
          v[4 * qd] = fmaf((__uint_as_float(r[4 * qd]) - mean1) * rstd1, w.x, bb.x);
          v[4 * qd + 1] = fmaf((__uint_as_float(r[4 * qd + 1]) - mean1) * rstd1, w.y, bb.y);
          v[4 * qd + 2] = fmaf((__uint_as_float(r[4 * qd + 2]) - mean1) * rstd1, w.z, bb.z);
          v[4 * qd + 3] = fmaf((__uint_as_float(r[4 * qd + 3]) - mean1) * rstd1, w.w, bb.w);
          if (e.mode == LNM_XY)
            stage_store16(tl, trow, qd, make_uint4(__float_as_uint(v[4 * qd]), __float_as_uint(v[4 * qd + 1]),
                                                   __float_as_uint(v[4 * qd + 2]), __float_as_uint(v[4 * qd + 3])));
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(v[j]);
        tmem_st32(taddr + cc * 32, r);
        st2.add32(v);
        if (e.mode == LNM_XY) {
          fence_proxy_async();
          named_bar_sync(bar_id, 128);
          if (issuer) { tma_store_2d(tma_x, tl, col0, row0); tma_store_commit(); }
          ++q;
        }
      }
      tmem_st_wait();
      exchange(st2, mean2, rstd2);
    }

    // ---------------- last pass: the normalised rows.  fp32 (LNM_FINAL only), then bf16
    const float mean_l = e.mode == LNM_Y ? mean1 : mean2, rstd_l = e.mode == LNM_Y ? rstd1 : rstd2;
    const float* lw = e.mode == LNM_Y ? e.ln1_w : e.ln2_w;
    const float* lb = e.mode == LNM_Y ? e.ln1_b : e.ln2_b;
    bool zero = false;
    if (e.mode == LNM_Y && e.row_limit != nullptr && row < M) {
      const int sq = row / e.rows_per_seq;
      zero = (row - sq * e.rows_per_seq) >= e.row_limit[sq];
    }
    if (e.mode == LNM_FINAL && e.store_f32) {
#pragma unroll 1
      for (int cc = 0; cc < NR; ++cc, ++q) {
        const int col0 = gcol0 + cc * 32;
        uint8_t* tl = stg + (q & 1u) * GEMM_STAGING_BYTES;
        uint32_t r[32];
        tmem_ld32(taddr + cc * 32, r);
        acquire_slot();
        tmem_ld_wait();
#pragma unroll
        for (int qd = 0; qd < 8; ++qd) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(lw + col0) + qd);
          const float4 bb = __ldg(reinterpret_cast<const float4*>(lb + col0) + qd);
          stage_store16(tl, trow, qd,
                        make_uint4(__float_as_uint(fmaf((__uint_as_float(r[4 * qd]) - mean_l) * rstd_l, w.x, bb.x)),
                                   __float_as_uint(fmaf((__uint_as_float(r[4 * qd + 1]) - mean_l) * rstd_l, w.y, bb.y)),
                                   __float_as_uint(fmaf((__uint_as_float(r[4 * qd + 2]) - mean_l) * rstd_l, w.z, bb.z)),
                                   __float_as_uint(fmaf((__uint_as_float(r[4 * qd + 3]) - mean_l) * rstd_l, w.w, bb.w))));
        }
        fence_proxy_async();
        named_bar_sync(bar_id, 128);
        if (issuer) { tma_store_2d(tma_x, tl, col0, row0); tma_store_commit(); }
      }
    }
    if (e.mode != LNM_FINAL || e.store_bf16) {
#pragma unroll 1
      for (int u = 0; u < NR / 2; ++u, ++q) {        // 64 bf16 columns (128 bytes per row) per staging use
        uint8_t* tl = stg + (q & 1u) * GEMM_STAGING_BYTES;
        acquire_slot();
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int cc = 2 * u + hf;
          const int col0 = gcol0 + cc * 32;
          uint32_t r[32];
          tmem_ld32(taddr + cc * 32, r);
          float4 w[8], bb[8];
#pragma unroll
          for (int qd = 0; qd < 8; ++qd) {
            w[qd] = __ldg(reinterpret_cast<const float4*>(lw + col0) + qd);
            bb[qd] = __ldg(reinterpret_cast<const float4*>(lb + col0) + qd);
          }
          tmem_ld_wait();
          uint32_t o[16];
#pragma unroll
          for (int qd = 0; qd < 8; ++qd) {
            const float y0 = fmaf((__uint_as_float(r[4 * qd]) - mean_l) * rstd_l, w[qd].x, bb[qd].x);
            const float y1 = fmaf((__uint_as_float(r[4 * qd + 1]) - mean_l) * rstd_l, w[qd].y, bb[qd].y);
            const float y2 = fmaf((__uint_as_float(r[4 * qd + 2]) - mean_l) * rstd_l, w[qd].z, bb[qd].z);
            const float y3 = fmaf((__uint_as_float(r[4 * qd + 3]) - mean_l) * rstd_l, w[qd].w, bb[qd].w);
            o[2 * qd] = zero ? 0u : pack_bf16(y0, y1);
            o[2 * qd + 1] = zero ? 0u : pack_bf16(y2, y3);
          }
#pragma unroll
          for (int qd = 0; qd < 4; ++qd)
            stage_store16(tl, trow, 4 * hf + qd, make_uint4(o[4 * qd], o[4 * qd + 1], o[4 * qd + 2], o[4 * qd + 3]));
        }
        fence_proxy_async();
        named_bar_sync(bar_id, 128);
        if (issuer) { tma_store_2d(tma_y, tl, gcol0 + 64 * u, row0); tma_store_commit(); }
      }
    }
  }
};

template <int NC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_ln_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const __grid_constant__ CUtensorMap tma_x, const __grid_constant__ CUtensorMap tma_r,
               const __grid_constant__ CUtensorMap tma_y, int M, int K, GemmLnParams ep) {
  constexpr int STAGES = gemmln_stages<NC>();
  constexpr uint32_t A_BYTES = GEMM_BM * 128;
  constexpr uint32_t B_BYTES = NC * 128;
  constexpr uint32_t TMEM_COLS = 2 * NC;
  constexpr int NCG = NC / 2;          // columns per epilogue group
  constexpr int NR = NCG / 32;         // 32-column rounds per group and pass
  constexpr int N = 2 * NC;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint8_t* sStage = sB + STAGES * B_BYTES;                                   // [2 groups][2 slots] x 16 KB
  float2* s_stat = reinterpret_cast<float2*>(sStage + 4 * GEMM_STAGING_BYTES);   // [2 buffers][4 partials][128 rows] (mean, M2)
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_stat + 2 * 4 * 128);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;      // [2]
  uint64_t* res_full = tempty_bar + 2;       // [2 groups][2 slots]
  uint64_t* stat_bar = res_full + 4;         // [2] used alternately; 16 warp arrivals per exchange: 8 local + 8 from the peer CTA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stat_bar + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int m_tiles = (M + GEMM_BM - 1) / GEMM_BM;
  const int k_blocks = K / GEMM_BK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    tma_prefetch_desc(&tma_x);
    tma_prefetch_desc(&tma_y);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], GEMM_EPI_WARPS); }
    for (int s = 0; s < 4; ++s) mbar_init(&res_full[s], 1);
    mbar_init(&stat_bar[0], 2 * GEMM_EPI_WARPS);
    mbar_init(&stat_bar[1], 2 * GEMM_EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();                          // the peer's barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------ TMA producer
    uint32_t stage = 0, phase = 0;
    for (int m_blk = cluster_id; m_blk < m_tiles; m_blk += num_clusters) {
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[stage], A_BYTES + B_BYTES);
          tma_load_2d(sA + stage * A_BYTES, &tma_a, &full_bar[stage], kb * GEMM_BK, m_blk * GEMM_BM);
          tma_load_2d(sB + stage * B_BYTES, &tma_b, &full_bar[stage], kb * GEMM_BK, int(rank) * NC);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(GEMM_BM, NC);
    const uint64_t da0 = make_sw128_desc(smem_u32(sA));
    const uint64_t db0 = make_sw128_desc(smem_u32(sB));
    uint32_t stage = 0, phase = 0;
    int it = 0;
    for (int m_blk = cluster_id; m_blk < m_tiles; m_blk += num_clusters, ++it) {
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * NC;
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t da = da0 + uint64_t((stage * A_BYTES) >> 4);
          const uint64_t db = db0 + uint64_t((stage * B_BYTES) >> 4);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) umma_bf16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&empty_bar[stage]);
          if (kb == k_blocks - 1) umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------ epilogue
    const int ew = warp - 2;
    const int quad = warp & 3;
    const int grp = ew >> 2;
    LnEpilogue<NCG> le;
    le.tma_x = &tma_x; le.tma_r = &tma_r; le.tma_y = &tma_y; le.ep = &ep;
    le.stg = sStage + grp * 2 * GEMM_STAGING_BYTES; le.rfull = &res_full[grp * 2]; le.stat_bar = stat_bar; le.s_stat = s_stat;
    le.n_part = 4; le.part_id = int(rank) * 2 + grp; le.rank = rank; le.bar_id = 1 + grp; le.trow = quad * 32 + lane;
    le.lane = lane; le.M = M; le.n_total = N; le.issuer = ((ew & 3) == 0) && lane == 0;
    const int gcol0 = int(rank) * NC + grp * NCG;       // first global column of this group's slab
    if (cluster_id < m_tiles) le.prefetch_first(gcol0, cluster_id * GEMM_BM, false);
    int it = 0;
    for (int m_blk = cluster_id; m_blk < m_tiles; m_blk += num_clusters, ++it) {
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + acc * NC + grp * NCG;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int next_blk = m_blk + num_clusters;
      le.tile(taddr, m_blk * GEMM_BM, gcol0, next_blk < m_tiles ? next_blk * GEMM_BM : -1);
      // the accumulator (and the x kept in it) is free again
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (next_blk < m_tiles) le.prefetch_first(gcol0, next_blk * GEMM_BM, true);
    }
    if (le.issuer) tma_store_wait_all();            // global writes complete before the CTA exits
  }

  tc_fence_before();
  cluster_sync_all();      // the peer may still write this CTA's statistics table / signal its barrier
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------------------
// Split-epilogue version (the default): the normalisation passes run on their OWN warps.
//
// In gemm_ln_kernel above the eight epilogue warps do everything in sequence for a row block: pass 1 (residual add, x store,
// statistics), the DSMEM exchange, then the normalisation pass(es).  They are latency-bound (one row per thread, long
// dependent chains, 2 warps per scheduler), so the kernel takes exactly the time of the residual GEMM plus the time of the
// stand-alone LayerNorm: while the normalisation passes run no residual load is in flight and the HBM read pipe idles.
// Here the work of a row block is a three-stage pipeline over the two TMEM accumulators:
//     MMA warp          : accumulator b <- A W^T                                     (waits for tempty[b])
//     warps 2..9  (P1)  : x_new = resid + mask * alpha * (acc + bias) -> TMEM b (+ TMA store for LNM_Y), per-row partial
//                         statistics published to both CTAs, then straight on to the next row block (no wait)
//     warps 10..  (P2)  : wait for the four partials of the row, normalise from TMEM b, store with 32-byte register stores
//                         (a thread owns a row: one full sector per lane and instruction), LNM_XY / LNM_FINAL: second
//                         statistics exchange among the P2 warps of both CTAs; then release accumulator b
// so pass 1 of row block i + 1 (the HBM-bound part) runs concurrently with the normalisation of row block i.
// Statistics tables are double-buffered by row-block parity.  The peer's P1 may run up to three row blocks ahead of this
// CTA's P2 (P1_peer(i+3) <- MMA_peer(i+3) <- P2_peer(i+1) <- P1_self(i+1), all possible while P2_self(i) is still busy), so a
// P1 warp may only overwrite the entry it wrote two row blocks ago in the PEER's table once the peer's P2 warps have read it:
// they return a credit (remote mbarrier arrive) right after merging the partials.  A CTA's own entries need no credit
// (P1_self(i+2) <- MMA_self(i+2) <- P2_self(i) finished).
// ---------------------------------------------------------------------------------------------------------------------
template <int NC> __host__ __device__ constexpr int gemmln_split_p2_groups() { return 2; }
template <int NC> constexpr int gemmln_split_threads() { return 64 + 256 + 128 * gemmln_split_p2_groups<NC>(); }
template <int NC> constexpr size_t gemmln_split_smem_bytes() {
  return size_t(gemmln_stages<NC>()) * (GEMM_BM * 128 + NC * 128) + 4 * GEMM_STAGING_BYTES +
         (2 * 2 + 2 * 2) * 128 * sizeof(float2) /*pass-1 partials: own groups, peer's groups; [2 parities] each*/ +
         2 * 2 * gemmln_split_p2_groups<NC>() * 128 * sizeof(float2) /*second exchange*/ + 1024 + 256;
}

// Remote (or local) shared-memory store whose completion is counted on an mbarrier of the destination CTA (complete_tx), like a
// TMA write: the reader needs only a CTA-scope wait on that barrier.  A cluster-scope acquire / release pair instead costs an L1
// invalidation (CCTL.IVALL) at every wait and a drain of all outstanding stores (ERRBAR) at every arrive: 20 % + 13 % of the
// warp-stall samples of this kernel (profiles/README.md).
CF_DEVINL void st_async_f32x2(uint32_t cluster_addr, float a, float b, uint32_t cluster_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(cluster_addr), "f"(a),
               "f"(b), "r"(cluster_bar)
               : "memory");
}
// `after`: a value computed from the shared-memory reads this arrival gives credit for.  It is an (unused) operand, so the
// arrival cannot be scheduled before those reads have returned their data (a warp issues in order).
CF_DEVINL void mbar_arrive_cluster_relaxed(uint32_t cluster_addr, float after) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];  // after %1" ::"r"(cluster_addr), "f"(after) : "memory");
}
CF_DEVINL void mbar_wait_cluster_relaxed(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.relaxed.cluster.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}

template <int NC> constexpr size_t gemmln_quad_smem_bytes() {
  return size_t(4) * (GEMM_BM * 128 + (NC / 2) * 128) + 4 * GEMM_STAGING_BYTES + (2 * 2 + 2 * 2) * 128 * sizeof(float2) +
         2 * 2 * gemmln_split_p2_groups<NC>() * 128 * sizeof(float2) + 1024 + 256;
}
CF_DEVINL void umma_commit_2sm_mask(uint64_t* bar, uint16_t mask) {   // arrive on `bar` (same offset) in the CTAs of `mask`
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}

CF_DEVINL void tma_load_3d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

CF_DEVINL void st_global_v8(void* p, const uint32_t (&o)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]),
               "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7])
               : "memory");
}

template <int NC, bool QUAD>
CF_DEVINL void gemm_ln_split_body(const CUtensorMap& tma_a, const CUtensorMap& tma_b, const CUtensorMap& tma_x, const CUtensorMap& tma_r,
                                  int M, int K, const GemmLnParams& ep, float* __restrict__ x_out, long long ldx,
                                  __nv_bfloat16* __restrict__ y_out, long long ldy) {
  constexpr int STAGES = QUAD ? 4 : gemmln_stages<NC>();
  constexpr int BMC = QUAD ? 256 : GEMM_BM;          // rows per cluster and row block
  constexpr uint32_t A_BYTES = GEMM_BM * 128;
  constexpr uint32_t B_BYTES = (QUAD ? NC / 2 : NC) * 128;   // QUAD: this CTA's half of the pair's NC weight rows
  constexpr uint32_t TMEM_COLS = 2 * NC;
  constexpr int NCG = NC / 2;          // columns per pass-1 group
  constexpr int NR = NCG / 32;
  constexpr int P2G = gemmln_split_p2_groups<NC>();
  constexpr int NC2 = NC / P2G;        // columns per pass-2 group
  constexpr int NR2 = NC2 / 32;
  constexpr int N = 2 * NC;
  constexpr int P1_WARPS = 8, P2_WARPS = 4 * P2G;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint8_t* sStage = sB + STAGES * B_BYTES;                                         // [2 groups][2 slots] x 16 KB
  float2* s_own = reinterpret_cast<float2*>(sStage + 4 * GEMM_STAGING_BYTES);      // [2][2 groups][128] (mean, M2) of this CTA's P1 groups
  float2* s_peer = s_own + 2 * 2 * 128;                                            // [2][2 groups][128] written by the peer's P1 groups
  float2* s_st2 = s_peer + 2 * 2 * 128;                                            // [2][2 * P2G][128] second exchange
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_st2 + 2 * 2 * P2G * 128);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;      // [2]   P2 warps release the accumulator
  uint64_t* res_full = tempty_bar + 2;       // [2 groups][2 slots]
  uint64_t* stat_bar = res_full + 4;         // [2]   this CTA's 8 P1 warps arrive, each expecting its peer counterpart's 32 x 8 bytes
  uint64_t* st2_bar = stat_bar + 2;          // [2]   same for the P2 warps
  uint64_t* cred_bar = st2_bar + 2;          // [2]   P2_WARPS arrivals from the PEER's P2 warps: "your partials have been read"
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(cred_bar + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  // pair kernel: rank = column half, statistics partner = rank ^ 1.  QUAD (cluster of four = two cta_group::2 pairs): CTA
  // (pair p = column half, half i = rows [128 i, 128 i + 128) of the cluster's 256-row block) has cluster rank 2 p + i; the pair
  // runs UMMA 256 x NC x 16 issued by its even-rank CTA, each CTA feeding its own 128 rows of A and half of the pair's weight
  // rows (32 KB per k-block and CTA instead of 48); the statistics partner is the CTA with the same rows, rank ^ 2.
  const uint32_t crank = cluster_ctarank();
  const uint32_t rank = QUAD ? (crank >> 1) : crank;          // column half
  const uint32_t half = QUAD ? (crank & 1u) : 0u;             // row half
  const uint32_t peer = QUAD ? (crank ^ 2u) : (crank ^ 1u);   // cluster rank of the statistics partner
  const uint32_t lead = crank & ~1u;                          // QUAD: cluster rank of this pair's MMA-issuing CTA
  const uint16_t pair_mask = uint16_t(3u << (crank & 2u));
  const bool leader = !QUAD || half == 0;
  const int cluster_id = QUAD ? (blockIdx.x >> 2) : (blockIdx.x >> 1);
  const int num_clusters = QUAD ? (gridDim.x >> 2) : (gridDim.x >> 1);
  const int m_tiles = (M + BMC - 1) / BMC;
  const int k_blocks = K / GEMM_BK;
  const int row_off = int(half) * GEMM_BM;                    // this CTA's first row inside the cluster's row block

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    tma_prefetch_desc(&tma_x);
    tma_prefetch_desc(&tma_r);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], QUAD ? 2 * P2_WARPS : P2_WARPS); }
    for (int s = 0; s < 4; ++s) mbar_init(&res_full[s], 1);
    for (int s = 0; s < 2; ++s) mbar_init(&stat_bar[s], P1_WARPS);     // + 256 bytes of remote partials per arrival (expect_tx)
    for (int s = 0; s < 2; ++s) mbar_init(&st2_bar[s], P2_WARPS);
    for (int s = 0; s < 2; ++s) mbar_init(&cred_bar[s], P2_WARPS);
    fence_barrier_init();
  }
  if (warp == 1) { if (QUAD) tmem_alloc_2sm(tmem_slot, TMEM_COLS); else tmem_alloc(tmem_slot, TMEM_COLS); }
  tc_fence_before();
  cluster_sync_all();                          // the peer's barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------ TMA producer
    uint32_t stage = 0, phase = 0;
    CF_PROF_DECL(w_empty = 0, t_all = clock64());
    for (int m_blk = cluster_id; m_blk < m_tiles; m_blk += num_clusters) {
      for (int kb = 0; kb < k_blocks; ++kb) {
        { CF_PROF_T0(); mbar_wait(&empty_bar[stage], phase ^ 1); CF_PROF_ADD(w_empty); }
        if (elect_one()) {
          if (QUAD) {                                  // completion of both CTAs' loads lands on the leader's barrier
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * (A_BYTES + B_BYTES));
            const uint32_t fb = mapa_rank(smem_u32(&full_bar[stage]), lead);
            const int arow = m_blk * BMC + row_off;
            if (ep.gather_rows > 0) tma_load_3d_2sm(sA + stage * A_BYTES, &tma_a, fb, kb * GEMM_BK, ep.gather_row0, arow / ep.gather_rows);
            else tma_load_2d_2sm(sA + stage * A_BYTES, &tma_a, fb, kb * GEMM_BK, arow);
            tma_load_2d_2sm(sB + stage * B_BYTES, &tma_b, fb, kb * GEMM_BK, int(rank) * NC + int(half) * (NC / 2));
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], A_BYTES + B_BYTES);
            if (ep.gather_rows > 0)
              tma_load_3d(sA + stage * A_BYTES, &tma_a, &full_bar[stage], kb * GEMM_BK, ep.gather_row0, m_blk * GEMM_BM / ep.gather_rows);
            else tma_load_2d(sA + stage * A_BYTES, &tma_a, &full_bar[stage], kb * GEMM_BK, m_blk * GEMM_BM);
            tma_load_2d(sB + stage * B_BYTES, &tma_b, &full_bar[stage], kb * GEMM_BK, int(rank) * NC);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
#ifdef CF_ABLATION
    if (ep.prof && lane == 0) { ep.prof[blockIdx.x * 16 + 0] = clock64() - t_all; ep.prof[blockIdx.x * 16 + 1] = w_empty; }
#endif
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (QUAD: the pair's even-rank CTA only)
    constexpr uint32_t idesc = make_idesc_bf16(QUAD ? 256 : GEMM_BM, NC);
    const uint64_t da0 = make_sw128_desc(smem_u32(sA));
    const uint64_t db0 = make_sw128_desc(smem_u32(sB));
    uint32_t stage = 0, phase = 0;
    int it = 0;
    CF_PROF_DECL(w_tempty = 0, w_full = 0, t_all = clock64());
    for (int m_blk = cluster_id; leader && m_blk < m_tiles; m_blk += num_clusters, ++it) {
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      { CF_PROF_T0(); mbar_wait(&tempty_bar[acc], acc_phase ^ 1); CF_PROF_ADD(w_tempty); }
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * NC;
      for (int kb = 0; kb < k_blocks; ++kb) {
        { CF_PROF_T0(); mbar_wait(&full_bar[stage], phase); CF_PROF_ADD(w_full); }
        tc_fence_after();
        if (elect_one()) {
          const uint64_t da = da0 + uint64_t((stage * A_BYTES) >> 4);
          const uint64_t db = db0 + uint64_t((stage * B_BYTES) >> 4);
          if (QUAD) {
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k) umma_bf16_ss_2sm(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
            umma_commit_2sm_mask(&empty_bar[stage], pair_mask);
            if (kb == k_blocks - 1) umma_commit_2sm_mask(&tfull_bar[acc], pair_mask);
          } else {
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k) umma_bf16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
            umma_commit(&empty_bar[stage]);
            if (kb == k_blocks - 1) umma_commit(&tfull_bar[acc]);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
#ifdef CF_ABLATION
    if (ep.prof && lane == 0) { ep.prof[blockIdx.x * 16 + 2] = clock64() - t_all; ep.prof[blockIdx.x * 16 + 3] = w_tempty; ep.prof[blockIdx.x * 16 + 4] = w_full; }
#endif
  } else if (warp < 2 + P1_WARPS) {
    // ------------------------------------------------ pass 1: x_new -> TMEM (+ global), partial statistics -> both CTAs
    const int ew = warp - 2;
    const int quad = warp & 3;
    const int grp = ew >> 2;
    const int trow = quad * 32 + lane;
    const int bar_id = 1 + grp;
    const bool issuer = ((ew & 3) == 0) && lane == 0;
    const bool has_res = ep.has_resid != 0 && !CF_LN_DBG(ep, 1);
    const bool store_x = ep.mode == LNM_Y;
    uint8_t* stg = sStage + grp * 2 * GEMM_STAGING_BYTES;
    uint64_t* rfull = &res_full[grp * 2];
    const int gcol0 = int(rank) * NC + grp * NCG;
    uint32_t q = 0, nl0 = 0, nl1 = 0;          // staging uses; residual loads issued into each slot (phase bookkeeping)
    auto issue_resid = [&](uint32_t slot, int col0, int row0) {
      mbar_arrive_expect_tx(&rfull[slot], GEMM_STAGING_BYTES);
      tma_load_2d(stg + slot * GEMM_STAGING_BYTES, &tma_r, &rfull[slot], col0, row0);
    };
    if (has_res && cluster_id < m_tiles) {
      if (issuer) issue_resid(0, gcol0, cluster_id * BMC + row_off);
      ++nl0;
    }
    int it = 0;
    CF_PROF_DECL(w_tfull = 0, w_res = 0, w_cred = 0, w_bar = 0, w_strd = 0, w_tld = 0, t_all = clock64(), seg[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, seg_t = 0);
    for (int m_blk = cluster_id; m_blk < m_tiles; m_blk += num_clusters, ++it) {
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + acc * NC + grp * NCG;
      const int row0 = m_blk * BMC + row_off;
      const int next_blk = m_blk + num_clusters;
      const int next_row0 = next_blk < m_tiles ? next_blk * BMC + row_off : -1;
      const int row = row0 + trow;
      bool keep = true;
      if (ep.row_range != nullptr && row < M) {
        const int ch = row / ep.rows_per_chunk;
        const int rr = row - ch * ep.rows_per_chunk;
        const int2 rg = ep.row_range[ch];
        keep = (rr >= rg.x && rr < rg.y);
      }
      if (issuer && has_res && next_row0 >= 0) {       // next row block's residual slab -> L2
#pragma unroll 1
        for (int cc = 0; cc < NR; ++cc) tma_prefetch_2d(&tma_r, gcol0 + cc * 32, next_row0);
      }
      { CF_PROF_T0(); mbar_wait(&tfull_bar[acc], acc_phase); CF_PROF_ADD(w_tfull); }
      tc_fence_after();
      RowStats st1;
#pragma unroll 1
      for (int cc = 0; cc < NR; ++cc, ++q) {
        const int col0 = gcol0 + cc * 32;
        const uint32_t slot = q & 1u;
        uint8_t* tl = stg + slot * GEMM_STAGING_BYTES;
        uint32_t r[32];
#ifdef CF_ABLATION
        seg_t = clock64();
#endif
        if (!CF_LN_DBG(ep, 256)) tmem_ld32(taddr + cc * 32, r);
        else {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = uint32_t(j + cc);
        }
        if (has_res) {
          // the residual sub-tile of the next round (or of the next row block's first round) into the other slot
          const bool in_blk = cc + 1 < NR;
          const int nr0 = in_blk ? row0 : next_row0, nc0 = in_blk ? col0 + 32 : gcol0;
          if (nr0 >= 0) {
            if (issuer) { { CF_PROF_T0(); tma_store_wait_read(); CF_PROF_ADD(w_strd); } issue_resid(slot ^ 1u, nc0, nr0); }
            if (slot) ++nl0; else ++nl1;
          }
        } else if (store_x) {
          if (issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          named_bar_sync(bar_id, 128);
        }
        CF_PROF_MARK(0);
        if (has_res) { CF_PROF_T0(); mbar_wait(&rfull[slot], ((slot ? nl1 : nl0) - 1u) & 1u); CF_PROF_ADD(w_res); }
        CF_PROF_MARK(1);
        { CF_PROF_T0(); tmem_ld_wait(); CF_PROF_ADD(w_tld); }
        CF_PROF_MARK(2);
        // Two halves of 16 columns: the four residual reads of a half are issued together and all precede the half's four
        // stores.  (Written as one load - compute - store sequence per 16 bytes the compiler must keep every load behind the
        // previous store to the same tile: eight serialised shared-memory round trips per round, 1 800 of a round's 4 500 cycles.)
        float v[32];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint4 xr4[4];
          float4 b[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int qd = 4 * hf + i;
            xr4[i] = (has_res && !CF_LN_DBG(ep, 4)) ? stage_load16(tl, trow, qd) : make_uint4(0u, 0u, 0u, 0u);
            b[i] = __ldg(reinterpret_cast<const float4*>(ep.bias + col0) + qd);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int qd = 4 * hf + i;
            const float4 x = make_float4(__uint_as_float(xr4[i].x), __uint_as_float(xr4[i].y), __uint_as_float(xr4[i].z), __uint_as_float(xr4[i].w));
            v[4 * qd] = keep ? fmaf(ep.alpha, __uint_as_float(r[4 * qd]) + b[i].x, x.x) : x.x;
            v[4 * qd + 1] = keep ? fmaf(ep.alpha, __uint_as_float(r[4 * qd + 1]) + b[i].y, x.y) : x.y;
            v[4 * qd + 2] = keep ? fmaf(ep.alpha, __uint_as_float(r[4 * qd + 2]) + b[i].z, x.z) : x.z;
            v[4 * qd + 3] = keep ? fmaf(ep.alpha, __uint_as_float(r[4 * qd + 3]) + b[i].w, x.w) : x.w;
          }
          if (store_x && !CF_LN_DBG(ep, 4)) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int qd = 4 * hf + i;
              stage_store16(tl, trow, qd, make_uint4(__float_as_uint(v[4 * qd]), __float_as_uint(v[4 * qd + 1]),
                                                     __float_as_uint(v[4 * qd + 2]), __float_as_uint(v[4 * qd + 3])));
            }
          }
        }
        CF_PROF_MARK(3);
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(v[j]);
        if (!CF_LN_DBG(ep, 128)) tmem_st32(taddr + cc * 32, r);
        CF_PROF_MARK(4);
        if (!CF_LN_DBG(ep, 32)) st1.add32(v);
        CF_PROF_MARK(5);
        if (store_x && !CF_LN_DBG(ep, 64)) fence_proxy_async();
        CF_PROF_MARK(6);
        if ((has_res || store_x) && !CF_LN_DBG(ep, 512)) { CF_PROF_T0(); named_bar_sync(bar_id, 128); CF_PROF_ADD(w_bar); }   // every thread is done with this slot
        CF_PROF_MARK(7);
        if (store_x && issuer && !CF_LN_DBG(ep, 2)) { tma_store_2d(&tma_x, tl, col0, row0); tma_store_commit(); }
        CF_PROF_MARK(8);
      }
      tmem_st_wait();
      // publish (mean, M2) of this thread's NCG columns into both CTAs' tables (entry it & 1)
      {
        const uint32_t par = uint32_t(it) & 1u, use = uint32_t(it) >> 1;
        if (use > 0) { CF_PROF_T0(); mbar_wait_cluster_relaxed(&cred_bar[par], (use - 1u) & 1u); CF_PROF_ADD(w_cred); }   // the peer has read what this thread wrote two row blocks ago
        float2* po = s_own + (par * 2u + uint32_t(grp)) * 128u + trow;
        float2* pp = s_peer + (par * 2u + uint32_t(grp)) * 128u + trow;
        *po = make_float2(st1.mean, st1.m2);
        st_async_f32x2(mapa_rank(smem_u32(pp), peer), st1.mean, st1.m2, mapa_rank(smem_u32(&stat_bar[par]), peer));
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_expect_tx(&stat_bar[par], 32 * sizeof(float2));   // the same warp of the peer sends as much
      }
    }
    if (issuer) tma_store_wait_all();            // global writes complete before the CTA exits
#ifdef CF_ABLATION
    if (ep.prof && ew == 0 && lane == 0) {
      long long* pr = ep.prof + blockIdx.x * 16;
      pr[5] = clock64() - t_all; pr[6] = w_tfull; pr[7] = w_res; pr[8] = w_cred; pr[9] = w_bar; pr[13] = w_strd; pr[14] = w_tld;
    }
    if (ep.prof && ew == 1 && lane == 0) ep.prof[blockIdx.x * 16 + 15] = w_bar;     // a warp that does not issue
    if (ep.prof && ew == 0 && lane == 0) for (int k = 0; k < 9; ++k) ep.prof[(gridDim.x + blockIdx.x) * 16 + k] = seg[k];
    if (ep.prof && ew == 1 && lane == 0) for (int k = 0; k < 9; ++k) ep.prof[(2 * gridDim.x + blockIdx.x) * 16 + k] = seg[k];
#endif
  } else {
    // ------------------------------------------------ pass 2: normalise from TMEM, register stores, release the accumulator
    const int pw = warp - (2 + P1_WARPS);
    const int quad = warp & 3;
    const int g2 = pw >> 2;
    const int trow = quad * 32 + lane;
    const int ccol0 = g2 * NC2;                          // first column of this group inside the CTA's accumulator
    const int gcol0 = int(rank) * NC + ccol0;
    const int part2 = int(rank) * P2G + g2;
    uint32_t xr2 = 0;                                    // second-statistics exchanges so far
    const bool two = ep.mode != LNM_Y;
    int it = 0;
    CF_PROF_DECL(w_stat = 0, w_st2 = 0, t_all = clock64());
    for (int m_blk = cluster_id; m_blk < m_tiles; m_blk += num_clusters, ++it) {
      const uint32_t acc = it & 1;
      const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + acc * NC + ccol0;
      const int row = m_blk * BMC + row_off + trow;
      const bool row_ok = row < M;
      const uint32_t par = uint32_t(it) & 1u;
      { CF_PROF_T0(); mbar_wait(&stat_bar[par], (uint32_t(it) >> 1) & 1u); CF_PROF_ADD(w_stat); }
      tc_fence_after();
      float mean1, rstd1;
      {
        float2 pv[4];
#pragma unroll
        for (int sl = 0; sl < 4; ++sl) {                 // fixed order (global column order): both CTAs get identical bits
          const bool own = (uint32_t(sl) >> 1) == rank;
          pv[sl] = own ? s_own[(par * 2u + (sl & 1)) * 128u + trow] : s_peer[(par * 2u + (sl & 1)) * 128u + trow];
        }
        RowStats tot;
#pragma unroll
        for (int sl = 0; sl < 4; ++sl) {
          if (sl == 0) { tot.n = float(NCG); tot.mean = pv[0].x; tot.m2 = pv[0].y; } else tot.merge(float(NCG), pv[sl].x, pv[sl].y);
        }
        mean1 = tot.mean;
        rstd1 = rsqrtf(tot.m2 / float(N) + 1e-5f);
        // credit: the partials are in registers (the merge above consumed them), the peer may overwrite its entries
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(mapa_rank(smem_u32(&cred_bar[par]), peer), rstd1);
      }
      float mean_l = mean1, rstd_l = rstd1;
      if (two) {
        // x_mid = LN1(x_new) -> TMEM (+ global for LNM_XY), its statistics, second exchange among the P2 warps of both CTAs
        RowStats st2;
        float* xrow = x_out + (long long)row * ldx + gcol0;
#pragma unroll 1
        for (int cc = 0; cc < NR2; ++cc) {
          const int col0 = gcol0 + cc * 32;
          uint32_t r[32];
          tmem_ld32(taddr + cc * 32, r);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int qd = 0; qd < 8; ++qd) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(ep.ln1_w + col0) + qd);
            const float4 bb = __ldg(reinterpret_cast<const float4*>(ep.ln1_b + col0) + qd);
            v[4 * qd] = fmaf((__uint_as_float(r[4 * qd]) - mean1) * rstd1, w.x, bb.x);
            v[4 * qd + 1] = fmaf((__uint_as_float(r[4 * qd + 1]) - mean1) * rstd1, w.y, bb.y);
            v[4 * qd + 2] = fmaf((__uint_as_float(r[4 * qd + 2]) - mean1) * rstd1, w.z, bb.z);
            v[4 * qd + 3] = fmaf((__uint_as_float(r[4 * qd + 3]) - mean1) * rstd1, w.w, bb.w);
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(v[j]);
          tmem_st32(taddr + cc * 32, r);
          if (ep.mode == LNM_XY && row_ok) {
#pragma unroll
            for (int s8 = 0; s8 < 4; ++s8) {
              uint32_t o[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = r[8 * s8 + e];
              st_global_v8(xrow + cc * 32 + 8 * s8, o);
            }
          }
          st2.add32(v);
        }
        tmem_st_wait();
        const uint32_t buf = xr2 & 1u;
        float2* p = s_st2 + (buf * uint32_t(2 * P2G) + uint32_t(part2)) * 128u + trow;
        *p = make_float2(st2.mean, st2.m2);
        st_async_f32x2(mapa_rank(smem_u32(p), peer), st2.mean, st2.m2, mapa_rank(smem_u32(&st2_bar[buf]), peer));
        __syncwarp();
        if (lane == 0) mbar_arrive_expect_tx(&st2_bar[buf], 32 * sizeof(float2));
        { CF_PROF_T0(); mbar_wait(&st2_bar[buf], (xr2 >> 1) & 1u); CF_PROF_ADD(w_st2); }
        RowStats tot;
#pragma unroll
        for (int sl = 0; sl < 2 * P2G; ++sl) {
          const float2 v = s_st2[(buf * uint32_t(2 * P2G) + sl) * 128u + trow];
          if (sl == 0) { tot.n = float(NC2); tot.mean = v.x; tot.m2 = v.y; } else tot.merge(float(NC2), v.x, v.y);
        }
        mean_l = tot.mean;
        rstd_l = rsqrtf(tot.m2 / float(N) + 1e-5f);
        ++xr2;
      }
      // ---------------- last pass: the normalised rows (bf16; LNM_FINAL: fp32 and / or bf16)
      const float* lw = two ? ep.ln2_w : ep.ln1_w;
      const float* lb = two ? ep.ln2_b : ep.ln1_b;
      bool zero = false;
      if (ep.mode == LNM_Y && ep.row_limit != nullptr && row_ok) {
        const int sq = row / ep.rows_per_seq;
        zero = (row - sq * ep.rows_per_seq) >= ep.row_limit[sq];
      }
      const bool st_f32 = ep.mode == LNM_FINAL && ep.store_f32 && row_ok && !CF_LN_DBG(ep, 8);
      const bool st_b16 = (ep.mode != LNM_FINAL || ep.store_bf16) && row_ok && !CF_LN_DBG(ep, 8);
      float* frow = x_out + (long long)row * ldx + gcol0;
      __nv_bfloat16* yrow = y_out + (long long)row * ldy + gcol0;
#pragma unroll 1
      for (int cc = 0; cc < (CF_LN_DBG(ep, 16) ? 0 : NR2); ++cc) {
        const int col0 = gcol0 + cc * 32;
        uint32_t r[32];
        tmem_ld32(taddr + cc * 32, r);
        float4 w[8], bb[8];
#pragma unroll
        for (int qd = 0; qd < 8; ++qd) {
          w[qd] = __ldg(reinterpret_cast<const float4*>(lw + col0) + qd);
          bb[qd] = __ldg(reinterpret_cast<const float4*>(lb + col0) + qd);
        }
        tmem_ld_wait();
        float y[32];
#pragma unroll
        for (int qd = 0; qd < 8; ++qd) {
          y[4 * qd] = fmaf((__uint_as_float(r[4 * qd]) - mean_l) * rstd_l, w[qd].x, bb[qd].x);
          y[4 * qd + 1] = fmaf((__uint_as_float(r[4 * qd + 1]) - mean_l) * rstd_l, w[qd].y, bb[qd].y);
          y[4 * qd + 2] = fmaf((__uint_as_float(r[4 * qd + 2]) - mean_l) * rstd_l, w[qd].z, bb[qd].z);
          y[4 * qd + 3] = fmaf((__uint_as_float(r[4 * qd + 3]) - mean_l) * rstd_l, w[qd].w, bb[qd].w);
        }
        if (st_f32) {
#pragma unroll
          for (int s8 = 0; s8 < 4; ++s8) {
            uint32_t o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = __float_as_uint(y[8 * s8 + e]);
            st_global_v8(frow + cc * 32 + 8 * s8, o);
          }
        }
        if (st_b16) {
#pragma unroll
          for (int s8 = 0; s8 < 2; ++s8) {
            uint32_t o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = zero ? 0u : pack_bf16(y[16 * s8 + 2 * e], y[16 * s8 + 2 * e + 1]);
            st_global_v8(yrow + cc * 32 + 16 * s8, o);
          }
        }
      }
      // the accumulator (and the x kept in it) is free again (QUAD: the pair's MMA issuer waits for both CTAs)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (QUAD) mbar_arrive_cluster_norelease(mapa_rank(smem_u32(&tempty_bar[acc]), lead));
        else mbar_arrive(&tempty_bar[acc]);
      }
    }
#ifdef CF_ABLATION
    if (ep.prof && pw == 0 && lane == 0) {
      long long* pr = ep.prof + blockIdx.x * 16;
      pr[10] = clock64() - t_all; pr[11] = w_stat; pr[12] = w_st2;
    }
#endif
  }

  tc_fence_before();
  cluster_sync_all();      // the peer may still write this CTA's statistics tables / signal its barriers
  if (warp == 1) { if (QUAD) tmem_dealloc_2sm(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS); }
}

template <int NC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(gemmln_split_threads<NC>(), 1)
gemm_ln_split_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                     const __grid_constant__ CUtensorMap tma_x, const __grid_constant__ CUtensorMap tma_r, int M, int K,
                     const __grid_constant__ GemmLnParams ep, float* __restrict__ x_out, long long ldx, __nv_bfloat16* __restrict__ y_out,
                     long long ldy) {
  gemm_ln_split_body<NC, false>(tma_a, tma_b, tma_x, tma_r, M, K, ep, x_out, ldx, y_out, ldy);
}

// Cluster of four CTAs = two cta_group::2 pairs per 256-row block (see gemm_ln_split_body): half the weight-operand traffic
// per CTA and a four-stage ring.  Only 33 clusters of four (132 of 148 SMs) can be resident on a B200 (GPC sizes), so this
// pays only where the pair kernel starves on operands (K = 2048).
template <int NC>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(gemmln_split_threads<NC>(), 1)
gemm_ln_quad_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                    const __grid_constant__ CUtensorMap tma_x, const __grid_constant__ CUtensorMap tma_r, int M, int K,
                    const __grid_constant__ GemmLnParams ep, float* __restrict__ x_out, long long ldx, __nv_bfloat16* __restrict__ y_out,
                    long long ldy) {
  gemm_ln_split_body<NC, true>(tma_a, tma_b, tma_x, tma_r, M, K, ep, x_out, ldx, y_out, ldy);
}


}  // namespace cf
