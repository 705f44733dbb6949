// Persistent warp-specialised tcgen05 GEMM for sm_100a:  C[M,N] = A[M,K] * B[N,K]^T  (bf16 in, fp32 accumulate)
//
//   warp 0      : TMA producer  (cp.async.bulk.tensor 2-D tiles, 128B swizzle, 4-stage mbarrier ring)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x 256 x 16, accumulators in TMEM)
//   warps 2..9  : epilogue, two groups of four warps (one warp per TMEM lane quadrant), each group owns 128 of the
//                 tile's 256 columns: tcgen05.ld -> registers -> fused bias / activation / GLU / residual -> 128B-swizzled
//                 staging tile in shared memory -> TMA store (cp.async.bulk.tensor ... bulk_group), or argmax partials.
//
// TMEM holds two 128x256 fp32 accumulators (512 columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
// Covers every projection of the path (SURVEY.md 2.3 K5,K6,K9,K10,K14,K15,K18,K20): the reference issues these as
// cuBLAS/cuDNN calls followed by separate elementwise kernels (positionwise_feed_forward.py:59, attention.py:99-101,150,
// convolution.py:220-221,250-253, subsampling.py:94-104,163-164, ctc.py:81).
#pragma once
#include "common.cuh"

namespace cf {

enum GemmEpi : int {
  EPI_BF16 = 0,    // out_bf16 = act(acc + bias)
  EPI_GLU = 1,     // weights row-interleaved (value, gate): out_bf16[:, j] = (acc[2j]+b[2j]) * sigmoid(acc[2j+1]+b[2j+1])
  EPI_F32 = 2,     // out_f32 = resid + rowmask * alpha * (acc + bias)      (resid / rowmask optional)
  EPI_ARGMAX = 4,  // per (row, n-tile half): best logit, runner-up, index of best
};
enum GemmAct : int { ACT_NONE = 0, ACT_RELU = 1, ACT_SILU = 2 };

struct GemmEpiParams {
  const float* bias = nullptr;   // [N]
  const float* resid = nullptr;  // EPI_F32
  long long ld_resid = 0;
  float alpha = 1.0f;
  const int2* row_range = nullptr;  // EPI_F32: row valid iff range[row / rows_per_chunk].x <= row % rows_per_chunk < .y
  int rows_per_chunk = 1;
  int resid_tma = 0;             // EPI_F32: residual sub-tiles are TMA-loaded into the staging tiles (set by launch_gemm)
  // Scattered bf16 / GLU output (compact streaming, api.cu; scatter_rows > 0, tma_c is then a 3-D map {columns, rows, groups}):
  // accumulator row m belongs to group m / scatter_rows and lands on row scatter_row0 + m % scatter_rows of that group
  // (scatter_rows divides 128); a 128-row tile is 128 / scatter_rows boxes of scatter_rows rows in one store.
  int scatter_rows = 0, scatter_row0 = 0;
#ifdef CF_ABLATION
  int debug = 0;                 // ablation builds only (tools/, -DCF_ABLATION): 1 = epilogue does nothing, 2 = no loads / MMAs,
                                 // 4 = bf16 epilogue computes but neither stages nor stores, 8 = bf16 epilogue stores straight from registers,
                                 // 16 = bf16 epilogue stores every tile to the first 128 rows (no DRAM write traffic), 32 = no TMA store
  void* raw_out = nullptr; long long raw_ldo = 0;   // output as a plain pointer (debug 8)
#endif
  float* part_best = nullptr;    // EPI_ARGMAX: [M, 2 * n_tiles]
  float* part_second = nullptr;
  int* part_index = nullptr;
};

// Timing / ablation switches exist only in the tools build (python -m chunkformer_b200.build --ablation); in the product
// library CF_DBG folds to 0 and every ablation branch is compiled out.
#ifdef CF_ABLATION
#define CF_DBG(ep, bits) (((ep).debug & (bits)) != 0)
#else
#define CF_DBG(ep, bits) (0)
#endif

constexpr int GEMM_BM = 128;
constexpr int GEMM_BN = 256;
constexpr int GEMM_BK = 64;
constexpr int GEMM_STAGES = 4;
constexpr int GEMM_EPI_WARPS = 8;
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;
constexpr uint32_t GEMM_STAGING_BYTES = 128 * 128;  // 128 rows x 128 bytes, one per epilogue group

// The fp32 epilogues double-buffer their staging tiles (a round fills one slot while the TMA store of the previous round still
// drains the other; with a residual, the residual sub-tile of the next round is TMA-loaded into the free slot) and give up one
// pipeline stage for them.  The bf16 epilogues keep four stages and one slot per group: measured on B200, trading the fourth
// stage for a second slot makes the K = 512 GEMMs 10 % slower (FFN w_1 306 -> 336 us).
__host__ __device__ constexpr int gemm_stages(int epi) { return epi == 2 /*EPI_F32*/ ? 3 : GEMM_STAGES; }
__host__ __device__ constexpr int gemm_staging_slots(int epi) { return epi == 2 ? 4 : 2; }
constexpr size_t gemm_smem_bytes(int epi) {
  return size_t(gemm_stages(epi)) * (GEMM_BM * 128 + GEMM_BN * 128) + gemm_staging_slots(epi) * GEMM_STAGING_BYTES + 1024 /*align slack*/ + 256;
}

// sigmoid / SiLU through one MUFU op (tanh.approx), so the SiLU epilogue stays under the MMA time of a K=512 tile.
CF_DEVINL float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
CF_DEVINL float silu_fast(float x) {
  const float h = 0.5f * x;
  return fmaf(h, tanh_approx(h), h);
}
CF_DEVINL float sigmoid_fast(float x) { return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f); }

CF_DEVINL void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

CF_DEVINL void tma_store_2d(const CUtensorMap* m, const void* smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
CF_DEVINL void tma_store_3d(const CUtensorMap* m, const void* smem_src, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
CF_DEVINL void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
CF_DEVINL void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
CF_DEVINL void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// 16-byte store into a 128B-swizzled [128 rows][128 B] staging tile (matches CU_TENSOR_MAP_SWIZZLE_128B).
CF_DEVINL void stage_store16(uint8_t* tile, int row, int slot, uint4 v) {
  *reinterpret_cast<uint4*>(tile + row * 128 + ((slot ^ (row & 7)) << 4)) = v;
}

// Tile schedule of a persistent CTA (or CTA pair) `cta` of `ncta`: tiles are dealt round-robin in (m, n) order (balanced to one tile).
CF_DEVINL bool gemm_tile_of(int t, int cta, int ncta, int m_tiles, int n_tiles, int& m_blk, int& n_blk) {
  const int tile = cta + t * ncta;
  m_blk = tile / n_tiles;
  n_blk = tile - m_blk * n_tiles;
  return tile < m_tiles * n_tiles;
}

CF_DEVINL uint4 stage_load16(const uint8_t* tile, int row, int slot) {
  return *reinterpret_cast<const uint4*>(tile + row * 128 + ((slot ^ (row & 7)) << 4));
}
CF_DEVINL void tma_prefetch_2d(const CUtensorMap* m, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1)
               : "memory");
}

// fp32 + residual epilogue of one 128 x 128 accumulator slab with the residual moved by TMA (N % 256 == 0, so every group
// of every tile runs exactly four 32-column rounds).  A thread owns one accumulator row; reading its 128-byte residual
// segment straight from global memory costs 32 L1 wavefronts per warp instruction and was the bottleneck of the K = 512
// GEMMs.  Instead the residual sub-tile [128 rows x 32 cols] of round rc+1 is TMA-loaded into the group's other staging
// slot while round rc is processed; threads add their accumulator row to it in shared memory (swizzled, conflict free)
// and the same slot is TMA-stored.  `rc` counts the group's rounds across tiles (slot = rc & 1, mbarrier phase = rc >> 1).
CF_DEVINL void gemm_epilogue_f32_tma(uint32_t taddr, int row0, int trow, int gcol0, int M, uint8_t* stg2, int bar_id, bool issuer,
                                     const CUtensorMap* tma_c, const CUtensorMap* tma_r, const GemmEpiParams& ep,
                                     uint64_t* res_full, uint32_t& rc, int next_row0, int next_gcol0) {
  const int row = row0 + trow;
  bool keep = true;
  if (ep.row_range != nullptr && row < M) {
    const int ch = row / ep.rows_per_chunk;
    const int rr = row - ch * ep.rows_per_chunk;
    const int2 rg = ep.row_range[ch];
    keep = (rr >= rg.x && rr < rg.y);
  }
  // masked rows get exactly +0 added (masked_fill_, convolution.py:253)
  if (issuer && next_row0 >= 0) {
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) tma_prefetch_2d(tma_r, next_gcol0 + cc * 32, next_row0);   // next tile's residual -> L2
  }
#pragma unroll 1
  for (int cc = 0; cc < 4; ++cc, ++rc) {
    const int col0 = gcol0 + cc * 32;
    const uint32_t slot = rc & 1u, ph = (rc >> 1) & 1u;
    uint8_t* tile = stg2 + slot * GEMM_STAGING_BYTES;
    uint32_t r[32];
    tmem_ld32(taddr + cc * 32, r);
    if (issuer) {
      const bool in_tile = cc < 3;
      const int nr0 = in_tile ? row0 : next_row0, nc0 = in_tile ? col0 + 32 : next_gcol0;
      if (nr0 >= 0) {
        tma_store_wait_read();                      // the store that last used the other slot has drained it
        mbar_arrive_expect_tx(&res_full[slot ^ 1u], GEMM_STAGING_BYTES);
        tma_load_2d(stg2 + (slot ^ 1u) * GEMM_STAGING_BYTES, tma_r, &res_full[slot ^ 1u], nc0, nr0);
      }
    }
    float4 b[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) b[q] = __ldg(reinterpret_cast<const float4*>(ep.bias + col0) + q);
    mbar_wait(&res_full[slot], ph);
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const uint4 xr = stage_load16(tile, trow, q);
      float4 v;
      v.x = keep ? fmaf(ep.alpha, __uint_as_float(r[4 * q]) + b[q].x, __uint_as_float(xr.x)) : __uint_as_float(xr.x);
      v.y = keep ? fmaf(ep.alpha, __uint_as_float(r[4 * q + 1]) + b[q].y, __uint_as_float(xr.y)) : __uint_as_float(xr.y);
      v.z = keep ? fmaf(ep.alpha, __uint_as_float(r[4 * q + 2]) + b[q].z, __uint_as_float(xr.z)) : __uint_as_float(xr.z);
      v.w = keep ? fmaf(ep.alpha, __uint_as_float(r[4 * q + 3]) + b[q].w, __uint_as_float(xr.w)) : __uint_as_float(xr.w);
      stage_store16(tile, trow, q, make_uint4(__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w)));
    }
    fence_proxy_async();
    named_bar_sync(bar_id, 128);
    if (issuer) { tma_store_2d(tma_c, tile, col0, row0); tma_store_commit(); }
  }
}

// Epilogue of one 128-row x 128-column accumulator slab (one epilogue group): shared by the 1-CTA and 2-CTA kernels.
//   taddr : TMEM address of the slab's first column for this warp's lane quadrant;  row0 : global row of tile row 0
//   gcol0 : first global accumulator column of the slab;  stg : the group's 16 KB staging tile
template <int EPI, int ACT>
CF_DEVINL void gemm_epilogue_slab(uint32_t taddr, int row0, int trow, int gcol0, int n_blk, int grp, int n_tiles, int M, int N,
                                  uint8_t* stg2, int bar_id, bool issuer, const CUtensorMap* tma_c, const GemmEpiParams& ep,
                                  uint32_t& rc, const float* pf_next = nullptr, uint32_t pf_bytes = 0) {
  // stg2: the group's staging slots (two for EPI_F32, one otherwise), used alternately (rc counts the group's store rounds);
  // with two slots only the store that used the slot two rounds ago has to be drained before it is refilled
  const int row = row0 + trow;
  const bool row_ok = row < M;

  if (EPI == EPI_ARGMAX) {
    // running (best, runner-up) without branches; the index is recovered only for chunks that raise the maximum
    float best = -INFINITY, second = -INFINITY;
    int best_idx = 0;
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
      const int col0 = gcol0 + cc * 32;
      if (col0 >= N) break;
      uint32_t r[32];
      tmem_ld32(taddr + cc * 32, r);
      float v[32];
      if (col0 + 32 <= N) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(ep.bias + col0) + q);
          v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = (col0 + j < N) ? __ldg(ep.bias + col0 + j) : -INFINITY;
      }
      tmem_ld_wait();
      const float before = best;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] += __uint_as_float(r[j]);
        second = fmaxf(second, fminf(best, v[j]));
        best = fmaxf(best, v[j]);
      }
      if (best > before) {                    // first column of this chunk that attains the new maximum (torch.argmax ties)
        int idx = 31;
#pragma unroll
        for (int j = 30; j >= 0; --j) idx = (v[j] == best) ? j : idx;
        best_idx = col0 + idx;
      }
    }
    if (row_ok) {
      const long long p = (long long)row * (2 * n_tiles) + 2 * n_blk + grp;
      ep.part_best[p] = best;
      ep.part_second[p] = second;
      ep.part_index[p] = best_idx;
    }
  } else if (EPI == EPI_F32) {
    bool keep = true;   // masked rows get exactly +0 added (masked_fill_, convolution.py:253)
    if (ep.row_range != nullptr && row_ok) {
      const int ch = row / ep.rows_per_chunk;
      const int rr = row - ch * ep.rows_per_chunk;
      const int2 rg = ep.row_range[ch];
      keep = (rr >= rg.x && rr < rg.y);
    }
    // The residual is read straight from global memory by the thread that owns the row.  Two things keep enough bytes in
    // flight for an HBM-bound K = 512 GEMM: the NEXT tile's residual rows are pulled into L2 with one bulk prefetch per
    // row as soon as this tile's epilogue starts, and inside the tile the loads of round cc+1 are issued before round cc
    // is consumed.
    if (pf_next != nullptr && pf_bytes != 0)
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pf_next), "r"(pf_bytes) : "memory");
    const bool has_res = ep.resid != nullptr && row_ok;
    const float* rs_row = has_res ? ep.resid + (long long)row * ep.ld_resid : nullptr;
    float4 xn[8];
    auto load_resid = [&](int col0) {
      const float4* rs = reinterpret_cast<const float4*>(rs_row + col0);
#pragma unroll
      for (int q = 0; q < 8; ++q) xn[q] = (col0 + 4 * q < N) ? rs[q] : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    if (has_res && gcol0 < N) load_resid(gcol0);
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {   // one 32-column fp32 sub-tile (128 B per row) per round
      const int col0 = gcol0 + cc * 32;
      if (col0 >= N) break;
      uint32_t r[32];
      tmem_ld32(taddr + cc * 32, r);
      float4 x[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) x[q] = xn[q];
      if (has_res && cc < 3 && col0 + 32 < N) load_resid(col0 + 32);
      constexpr uint32_t SLOT_MASK = gemm_staging_slots(EPI) / 2 - 1;
      uint8_t* stg = stg2 + (rc++ & SLOT_MASK) * GEMM_STAGING_BYTES;
      if (issuer) { if (SLOT_MASK) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); else tma_store_wait_read(); }
      named_bar_sync(bar_id, 128);       // staging slot free again
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col0 + 4 * q < N) b = __ldg(reinterpret_cast<const float4*>(ep.bias + col0) + q);
        float4 v;
        v.x = ep.alpha * (__uint_as_float(r[4 * q]) + b.x);
        v.y = ep.alpha * (__uint_as_float(r[4 * q + 1]) + b.y);
        v.z = ep.alpha * (__uint_as_float(r[4 * q + 2]) + b.z);
        v.w = ep.alpha * (__uint_as_float(r[4 * q + 3]) + b.w);
        if (!keep) v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has_res) { v.x += x[q].x; v.y += x[q].y; v.z += x[q].z; v.w += x[q].w; }
        stage_store16(stg, trow, q, make_uint4(__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w)));
      }
      fence_proxy_async();
      named_bar_sync(bar_id, 128);
      if (issuer) { tma_store_2d(tma_c, stg, col0, row0); tma_store_commit(); }
    }
  } else {
    // bf16 outputs: EPI_BF16 -> two 64-column sub-tiles per group; EPI_GLU -> one 64-column sub-tile (128 acc columns)
    constexpr int ROUNDS = (EPI == EPI_GLU) ? 1 : 2;
    constexpr int CH_PER_ROUND = (EPI == EPI_GLU) ? 4 : 2;
#pragma unroll 1
    for (int rd = 0; rd < ROUNDS; ++rd) {
      const int acol0 = gcol0 + rd * 64;       // accumulator column of this round (EPI_BF16)
      if (acol0 >= N) break;
      constexpr uint32_t SLOT_MASK = gemm_staging_slots(EPI) / 2 - 1;
      uint8_t* stg = stg2 + (rc++ & SLOT_MASK) * GEMM_STAGING_BYTES;
      const bool no_stage = CF_DBG(ep, 12) != 0;
      if (!no_stage) {
        if (issuer) { if (SLOT_MASK) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); else tma_store_wait_read(); }
        named_bar_sync(bar_id, 128);
      }
#pragma unroll
      for (int cc = 0; cc < CH_PER_ROUND; ++cc) {
        const int tcol = (EPI == EPI_GLU) ? cc * 32 : rd * 64 + cc * 32;   // column inside the group's 128
        const int col0 = gcol0 + tcol;
        uint32_t r[32];
        tmem_ld32(taddr + tcol, r);
        float b[32];
        if (col0 < N) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(ep.bias + col0) + q);
            b[4 * q] = t.x; b[4 * q + 1] = t.y; b[4 * q + 2] = t.z; b[4 * q + 3] = t.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) b[j] = 0.f;
        }
        tmem_ld_wait();
        if (EPI == EPI_GLU) {
          uint32_t o[8];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float a0 = __uint_as_float(r[j]) + b[j], g0 = __uint_as_float(r[j + 1]) + b[j + 1];
            const float a1 = __uint_as_float(r[j + 2]) + b[j + 2], g1 = __uint_as_float(r[j + 3]) + b[j + 3];
            o[j >> 2] = pack_bf16(a0 * sigmoid_fast(g0), a1 * sigmoid_fast(g1));
          }
          stage_store16(stg, trow, 2 * cc, make_uint4(o[0], o[1], o[2], o[3]));
          stage_store16(stg, trow, 2 * cc + 1, make_uint4(o[4], o[5], o[6], o[7]));
        } else {
          uint32_t o[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float v0 = __uint_as_float(r[j]) + b[j], v1 = __uint_as_float(r[j + 1]) + b[j + 1];
            if (ACT == ACT_RELU) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
            if (ACT == ACT_SILU) { v0 = silu_fast(v0); v1 = silu_fast(v1); }
            o[j >> 1] = pack_bf16(v0, v1);
          }
#ifdef CF_ABLATION
          if CF_DBG(ep, 8) {
            if (row0 + trow < M && col0 < N) {
              __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(ep.raw_out) + (long long)(row0 + trow) * ep.raw_ldo + col0;
#pragma unroll
              for (int q = 0; q < 2; ++q)       // 256-bit stores: one full 32-byte sector per lane and instruction
                asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 16 * q), "r"(o[8 * q]), "r"(o[8 * q + 1]),
                             "r"(o[8 * q + 2]), "r"(o[8 * q + 3]), "r"(o[8 * q + 4]), "r"(o[8 * q + 5]), "r"(o[8 * q + 6]), "r"(o[8 * q + 7])
                             : "memory");
            }
          } else if CF_DBG(ep, 4) {
#pragma unroll
            for (int q = 0; q < 16; ++q) asm volatile("" ::"r"(o[q]));
          } else
#endif
          {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              stage_store16(stg, trow, 4 * cc + q, make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]));
          }
        }
      }
      if (no_stage) continue;
      fence_proxy_async();
      named_bar_sync(bar_id, 128);
      if (issuer && !CF_DBG(ep, 32)) {
        const int ocol = (EPI == EPI_GLU) ? (gcol0 >> 1) : acol0;
        const int srow0 = CF_DBG(ep, 16) ? 0 : row0;   // debug 16: every tile lands on the first 128 rows (L2 only)
        if (ep.scatter_rows > 0) tma_store_3d(tma_c, stg, ocol, ep.scatter_row0, srow0 / ep.scatter_rows);   // tma_c is a 3-D map
        else tma_store_2d(tma_c, stg, ocol, srow0);   // (a one-group 3-D store measured 12 % slower on FFN w_1 than the 2-D form)
        tma_store_commit();
      }
    }
  }
}

template <int EPI, int ACT>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                    const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_r, int M, int N, int K,
                    GemmEpiParams ep) {
  constexpr int BN = GEMM_BN;
  constexpr uint32_t A_BYTES = GEMM_BM * 128;
  constexpr uint32_t B_BYTES = BN * 128;
  constexpr uint32_t TMEM_COLS = 2 * BN;
  constexpr int STAGES = gemm_stages(EPI);
  constexpr int SLOTS = gemm_staging_slots(EPI);

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // align with pointer arithmetic (not an integer round trip) so the compiler keeps the shared address space (STS, not ST.E)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint8_t* sStage = sB + STAGES * B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + SLOTS * GEMM_STAGING_BYTES);
  uint64_t* full_bar = bars;                     // [STAGES]
  uint64_t* empty_bar = bars + STAGES;      // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;  // [2]
  uint64_t* tempty_bar = tfull_bar + 2;          // [2]
  uint64_t* res_full = tempty_bar + 2;            // [2 groups][2 slots]  (EPI_F32 with TMA residual)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_full + 4);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // provably warp-uniform: keeps the role branches convergent
  const int lane = threadIdx.x & 31;
  const int m_tiles = (M + GEMM_BM - 1) / GEMM_BM;
  const int n_tiles = (N + BN - 1) / BN;
  const int k_blocks = (K + GEMM_BK - 1) / GEMM_BK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    if (EPI != EPI_ARGMAX) tma_prefetch_desc(&tma_c);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], GEMM_EPI_WARPS);      // one arrival per epilogue warp
    }
    for (int s = 0; s < 4; ++s) mbar_init(&res_full[s], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------ TMA producer (whole warp walks the ring, one elected lane issues)
    {
      uint32_t stage = 0, phase = 0;
      for (int t = 0;; ++t) {
        int m_blk, n_blk;
        if (!gemm_tile_of(t, blockIdx.x, gridDim.x, m_tiles, n_tiles, m_blk, n_blk)) break;
        for (int kb = 0; kb < k_blocks && !CF_DBG(ep, 2); ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&full_bar[stage], A_BYTES + B_BYTES);
            tma_load_2d(sA + stage * A_BYTES, &tma_a, &full_bar[stage], kb * GEMM_BK, m_blk * GEMM_BM);
            tma_load_2d(sB + stage * B_BYTES, &tma_b, &full_bar[stage], kb * GEMM_BK, n_blk * BN);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer: the whole warp walks the pipeline (convergent control flow,
    // warp-uniform descriptors in uniform registers), one elected lane issues the tcgen05 instructions
    {
      constexpr uint32_t idesc = make_idesc_bf16(GEMM_BM, BN);
      const uint64_t da0 = make_sw128_desc(smem_u32(sA));
      const uint64_t db0 = make_sw128_desc(smem_u32(sB));
      uint32_t stage = 0, phase = 0;
      for (int it = 0;; ++it) {
        int m_blk, n_blk;
        if (!gemm_tile_of(it, blockIdx.x, gridDim.x, m_tiles, n_tiles, m_blk, n_blk)) break;
        const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        if CF_DBG(ep, 2) {
          if (elect_one()) umma_commit(&tfull_bar[acc]);
          __syncwarp();
          continue;
        }
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t da = da0 + uint64_t((stage * A_BYTES) >> 4);
            const uint64_t db = db0 + uint64_t((stage * B_BYTES) >> 4);
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k) {
              // advance 16 K-elements = 32 bytes inside the 128B swizzle atom: +2 in the (addr >> 4) field
              umma_bf16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
            }
            umma_commit(&empty_bar[stage]);
            if (kb == k_blocks - 1) umma_commit(&tfull_bar[acc]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------ epilogue: group g = columns [128 g, 128 g + 128) of the tile
    const int ew = warp - 2;
    const int quad = warp & 3;          // TMEM lane quadrant this warp may read
    const int grp = ew >> 2;
    const bool issuer = ((ew & 3) == 0) && lane == 0;
    const int bar_id = 1 + grp;
    uint8_t* stg = sStage + grp * (SLOTS / 2) * GEMM_STAGING_BYTES;
    const int trow = quad * 32 + lane;  // row inside the tile
    const bool res_tma = (EPI == EPI_F32) && ep.resid_tma != 0;
    uint32_t rc = 0;                    // rounds of this group so far (TMA-residual epilogue)
    if (res_tma && issuer) {            // residual sub-tile of the very first round
      int m0, n0;
      if (gemm_tile_of(0, blockIdx.x, gridDim.x, m_tiles, n_tiles, m0, n0)) {
        mbar_arrive_expect_tx(&res_full[grp * 2], GEMM_STAGING_BYTES);
        tma_load_2d(stg, &tma_r, &res_full[grp * 2], n0 * BN + grp * 128, m0 * GEMM_BM);
      }
    }
    for (int it = 0;; ++it) {
      int m_blk, n_blk;
      if (!gemm_tile_of(it, blockIdx.x, gridDim.x, m_tiles, n_tiles, m_blk, n_blk)) break;
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + acc * BN + grp * 128;
      const float* pf_next = nullptr;
      uint32_t pf_bytes = 0;
      if (EPI == EPI_F32 && ep.resid != nullptr && !res_tma) {
        int m2, n2;
        if (gemm_tile_of(it + 1, blockIdx.x, gridDim.x, m_tiles, n_tiles, m2, n2)) {
          const int prow = m2 * GEMM_BM + trow, pcol = n2 * BN + grp * 128;
          if (prow < M && pcol < N) { pf_next = ep.resid + (long long)prow * ep.ld_resid + pcol; pf_bytes = uint32_t(min(128, N - pcol)) * 4u; }
        }
      }
      if CF_DBG(ep, 1) {
      } else if (res_tma) {
        int m2, n2, nrow0 = -1, ncol0 = 0;
        if (gemm_tile_of(it + 1, blockIdx.x, gridDim.x, m_tiles, n_tiles, m2, n2)) { nrow0 = m2 * GEMM_BM; ncol0 = n2 * BN + grp * 128; }
        gemm_epilogue_f32_tma(taddr, m_blk * GEMM_BM, trow, n_blk * BN + grp * 128, M, stg, bar_id, issuer, &tma_c, &tma_r, ep,
                              &res_full[grp * 2], rc, nrow0, ncol0);
      } else {
        gemm_epilogue_slab<EPI, ACT>(taddr, m_blk * GEMM_BM, trow, n_blk * BN + grp * 128, n_blk, grp, n_tiles, M, N, stg, bar_id,
                                     issuer, &tma_c, ep, rc, pf_next, pf_bytes);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
    if (issuer && EPI != EPI_ARGMAX) tma_store_wait_all();   // global writes complete before the CTA exits
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------------------
// 2-CTA variant (cta_group::2): a cluster of two CTAs on one TPC computes a 256 x 256 tile with UMMA 256x256x16.  Each CTA
// loads its own 128 rows of A and 128 of the 256 B rows (so the L2 -> shared traffic per FLOP drops by a third), the
// leader CTA issues the MMAs for the pair, each CTA keeps the accumulator rows of its own M half in its TMEM and runs the
// same epilogue on it.  Producer of the peer CTA signals the leader's full barrier (complete_tx on a cluster-mapped
// address); tcgen05.commit multicasts the "stage free" / "accumulator ready" arrivals to both CTAs; the epilogues of both
// CTAs arrive on the leader's "accumulator drained" barrier.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int GEMM2_STAGES = 6;
__host__ __device__ constexpr int gemm2_stages(int epi) { return epi == 2 ? 5 : GEMM2_STAGES; }
constexpr size_t gemm2_smem_bytes(int epi) {
  return size_t(gemm2_stages(epi)) * (GEMM_BM * 128 + 128 * 128) + gemm_staging_slots(epi) * GEMM_STAGING_BYTES + 1024 + 256;
}

CF_DEVINL uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
CF_DEVINL void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
CF_DEVINL uint32_t mapa_rank(uint32_t smem_addr, uint32_t rank) {   // shared::cta address -> shared::cluster address in `rank`
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
CF_DEVINL void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Without the cluster-scope release (which drains every outstanding store of the warp: ERRBAR): for hand-offs whose payload is
// not ordinary memory, e.g. "my tcgen05.ld reads of this accumulator have completed" (tcgen05.wait::ld + fence::before_thread_sync).
CF_DEVINL void mbar_arrive_cluster_norelease(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
CF_DEVINL void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
CF_DEVINL void umma_bf16_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
CF_DEVINL void umma_commit_2sm(uint64_t* bar) {   // arrive on `bar` (same offset) in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(uint16_t(3))
               : "memory");
}
CF_DEVINL void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
CF_DEVINL void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

template <int EPI, int ACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                     const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_r, int M, int N, int K,
                     GemmEpiParams ep) {
  constexpr int STAGES = gemm2_stages(EPI);
  constexpr int SLOTS = gemm_staging_slots(EPI);
  constexpr int BN = GEMM_BN;                 // 256 accumulator columns per CTA
  constexpr uint32_t A_BYTES = GEMM_BM * 128; // this CTA's 128 rows of A
  constexpr uint32_t B_BYTES = 128 * 128;     // this CTA's 128 of the tile's 256 B rows
  constexpr uint32_t TMEM_COLS = 2 * BN;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint8_t* sStage = sB + STAGES * B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + SLOTS * GEMM_STAGING_BYTES);
  uint64_t* full_bar = bars;                      // [STAGES]  (used in the leader CTA)
  uint64_t* empty_bar = bars + STAGES;      // [STAGES]  (one per CTA)
  uint64_t* tfull_bar = bars + 2 * STAGES;  // [2]       (one per CTA)
  uint64_t* tempty_bar = tfull_bar + 2;           // [2]       (used in the leader CTA, 2 x 256 arrivals)
  uint64_t* res_full = tempty_bar + 2;            // [2 groups][2 slots]  (EPI_F32 with TMA residual; one set per CTA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_full + 4);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int m_tiles = (M + 255) / 256;
  const int n_tiles = (N + BN - 1) / BN;
  const int k_blocks = (K + GEMM_BK - 1) / GEMM_BK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    if (EPI != EPI_ARGMAX) tma_prefetch_desc(&tma_c);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 2 * GEMM_EPI_WARPS);  // one (remote) arrival per epilogue warp of either CTA
    }
    for (int s = 0; s < 4; ++s) mbar_init(&res_full[s], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------ TMA producer (both CTAs; completion lands on the leader's barrier)
    {
      uint32_t stage = 0, phase = 0;
      for (int t = 0;; ++t) {
        int m_blk, n_blk;
        if (!gemm_tile_of(t, cluster_id, num_clusters, m_tiles, n_tiles, m_blk, n_blk)) break;
        for (int kb = 0; kb < k_blocks && !CF_DBG(ep, 2); ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * (A_BYTES + B_BYTES));
            const uint32_t fb = mapa_rank(smem_u32(&full_bar[stage]), 0);
            tma_load_2d_2sm(sA + stage * A_BYTES, &tma_a, fb, kb * GEMM_BK, m_blk * 256 + int(rank) * 128);
            tma_load_2d_2sm(sB + stage * B_BYTES, &tma_b, fb, kb * GEMM_BK, n_blk * BN + int(rank) * 128);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (leader CTA only; convergent warp, elected lane issues)
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(256, BN);
      const uint64_t da0 = make_sw128_desc(smem_u32(sA));
      const uint64_t db0 = make_sw128_desc(smem_u32(sB));
      uint32_t stage = 0, phase = 0;
      for (int it = 0;; ++it) {
        int m_blk, n_blk;
        if (!gemm_tile_of(it, cluster_id, num_clusters, m_tiles, n_tiles, m_blk, n_blk)) break;
        const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        if CF_DBG(ep, 2) {
          if (elect_one()) umma_commit_2sm(&tfull_bar[acc]);
          __syncwarp();
          continue;
        }
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t da = da0 + uint64_t((stage * A_BYTES) >> 4);
            const uint64_t db = db0 + uint64_t((stage * B_BYTES) >> 4);
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k) umma_bf16_ss_2sm(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
            umma_commit_2sm(&empty_bar[stage]);
            if (kb == k_blocks - 1) umma_commit_2sm(&tfull_bar[acc]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------ epilogue (both CTAs, each on the accumulator rows of its M half)
    const int ew = warp - 2;
    const int quad = warp & 3;
    const int grp = ew >> 2;
    const bool issuer = ((ew & 3) == 0) && lane == 0;
    const int bar_id = 1 + grp;
    uint8_t* stg = sStage + grp * (SLOTS / 2) * GEMM_STAGING_BYTES;
    const int trow = quad * 32 + lane;
    const bool res_tma = (EPI == EPI_F32) && ep.resid_tma != 0;
    uint32_t rc = 0;
    if (res_tma && issuer) {
      int m0, n0;
      if (gemm_tile_of(0, cluster_id, num_clusters, m_tiles, n_tiles, m0, n0)) {
        mbar_arrive_expect_tx(&res_full[grp * 2], GEMM_STAGING_BYTES);
        tma_load_2d(stg, &tma_r, &res_full[grp * 2], n0 * BN + grp * 128, m0 * 256 + int(rank) * 128);
      }
    }
    for (int it = 0;; ++it) {
      int m_blk, n_blk;
      if (!gemm_tile_of(it, cluster_id, num_clusters, m_tiles, n_tiles, m_blk, n_blk)) break;
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + acc * BN + grp * 128;
      const float* pf_next = nullptr;
      uint32_t pf_bytes = 0;
      if (EPI == EPI_F32 && ep.resid != nullptr && !res_tma) {
        int m2, n2;
        if (gemm_tile_of(it + 1, cluster_id, num_clusters, m_tiles, n_tiles, m2, n2)) {
          const int prow = m2 * 256 + int(rank) * 128 + trow, pcol = n2 * BN + grp * 128;
          if (prow < M && pcol < N) { pf_next = ep.resid + (long long)prow * ep.ld_resid + pcol; pf_bytes = uint32_t(min(128, N - pcol)) * 4u; }
        }
      }
      if CF_DBG(ep, 1) {
      } else if (res_tma) {
        int m2, n2, nrow0 = -1, ncol0 = 0;
        if (gemm_tile_of(it + 1, cluster_id, num_clusters, m_tiles, n_tiles, m2, n2)) { nrow0 = m2 * 256 + int(rank) * 128; ncol0 = n2 * BN + grp * 128; }
        gemm_epilogue_f32_tma(taddr, m_blk * 256 + int(rank) * 128, trow, n_blk * BN + grp * 128, M, stg, bar_id, issuer, &tma_c, &tma_r, ep,
                              &res_full[grp * 2], rc, nrow0, ncol0);
      } else {
        gemm_epilogue_slab<EPI, ACT>(taddr, m_blk * 256 + int(rank) * 128, trow, n_blk * BN + grp * 128, n_blk, grp, n_tiles, M, N, stg,
                                     bar_id, issuer, &tma_c, ep, rc, pf_next, pf_bytes);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_norelease(mapa_rank(smem_u32(&tempty_bar[acc]), 0));
    }
    if (issuer && EPI != EPI_ARGMAX) tma_store_wait_all();
  }

  tc_fence_before();
  cluster_sync_all();      // no CTA of the pair may exit (or free TMEM) while the other can still signal it
  if (warp == 1) tmem_dealloc_2sm(tmem_base, TMEM_COLS);
}

}  // namespace cf
