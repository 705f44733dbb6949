// Position-wise feed-forward module in ONE kernel (positionwise_feed_forward.py:51-60 + the residual update and LayerNorm(s)
// of encoder_layer.py:190-199 / 236-246):
//
//     x_new = x + 0.5 * (W2 SiLU(W1 y + b1) + b2),   then LN1 / LN2 of x_new as in gemm_ln.cuh (LNM_Y, LNM_XY, LNM_FINAL)
//
// The [rows, F] hidden activation (740 MB per FFN at the benchmark batch) never leaves the SM: a pair of CTAs (cluster of 2)
// owns a 128-row block and walks the hidden dimension in chunks of 256 columns.  Per chunk, CTA r
//   GEMM1 : Hacc[128 x 128] = Y[128 x d] W1[chunk cols r*128 .. +128, :]^T            (UMMA 128 x 128 x 16, accumulator in TMEM)
//   SiLU  : four epilogue warps read Hacc, add b1, SiLU, round to bf16 and write the tile as two 128B-swizzled K-major atoms
//           (128 rows x 64 hidden columns, the UMMA A-operand layout) into their own shared memory AND, with one
//           cp.async.bulk shared::cta -> shared::cluster copy per atom, into the peer's, so both CTAs hold all 256 columns
//   GEMM2 : OUT[128 x d/2] += H[128 x 256] W2[out rows r*d/2 .. +d/2, chunk cols]^T   (UMMA 128 x d/2 x 16, A straight from the atoms)
// OUT (d/2 fp32 columns) stays in TMEM for the whole row block; two Hacc buffers (128 columns each) let GEMM1 of chunk c+1 run
// while chunk c is in the SiLU epilogue.  After the last chunk four more warps run the LayerNorm epilogue of gemm_ln.cuh on
// OUT (row statistics exchanged with the peer over DSMEM).  TMEM: OUT [0, d/2) | Hacc0 | Hacc1 = 512 columns at d = 512.
//
//   warp 0 : TMA producer, one ring of 32 KB stages shared by both GEMMs (GEMM1 stage = Y k-block + W1 k-block, GEMM2 stage =
//            W2 k-block), filled in exactly the order the MMA warp consumes
//   warp 1 : TMEM allocator + MMA issuer; order per row block: G1(0), then for every chunk c: G1(c+1), G2(c)
//   warps 2..5 : SiLU epilogue (one TMEM lane quadrant each)          warps 6..9 : LayerNorm epilogue of the row block
//
// Hand-offs (mbarriers): ring full/empty; hacc_full/hacc_empty per Hacc buffer; per atom h_full (own arrival, or the peer's
// remote arrive.expect_tx + the bulk copy's complete_tx) and h_free (tcgen05.commit multicast to BOTH CTAs: an atom may be
// rewritten only when the GEMM2 k-block that read it has retired in both); out_full/out_empty for OUT.
#pragma once
#include "gemm_ln.cuh"

namespace cf {

constexpr int FFN_HC = 128;                    // hidden columns per CTA and chunk
constexpr int FFN_CHUNK = 2 * FFN_HC;          // hidden columns per chunk (both CTAs)
constexpr uint32_t FFN_STAGE_BYTES = 32768;
constexpr uint32_t FFN_ATOM_BYTES = 128 * 128; // 128 rows x 64 bf16
template <int NC> __host__ __device__ constexpr int ffn_stages() { return 3; }
template <int NC> constexpr size_t ffn_smem_bytes() {
  return size_t(ffn_stages<NC>()) * FFN_STAGE_BYTES + 4 * FFN_ATOM_BYTES + 2 * GEMM_STAGING_BYTES + 2 * 2 * 128 * sizeof(float2) + 1024 + 512;
}

CF_DEVINL void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {   // arrive on `bar` (same offset) in every CTA of the mask
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
CF_DEVINL void mbar_arrive_expect_tx_cluster(uint32_t cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(bytes) : "memory");
}
// shared memory of this CTA -> shared memory of another CTA of the cluster; completes (bytes) on an mbarrier of the destination CTA
CF_DEVINL void bulk_copy_to_cluster(uint32_t dst_cluster_addr, const void* src_smem, uint32_t bytes, uint32_t bar_cluster_addr) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster_addr),
               "r"(smem_u32(src_smem)), "r"(bytes), "r"(bar_cluster_addr)
               : "memory");
}

template <int NC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
ffn_fused_kernel(const __grid_constant__ CUtensorMap tma_y, const __grid_constant__ CUtensorMap tma_w1,
                 const __grid_constant__ CUtensorMap tma_w2, const __grid_constant__ CUtensorMap tma_x,
                 const __grid_constant__ CUtensorMap tma_r, const __grid_constant__ CUtensorMap tma_yo, int M, int F,
                 const float* __restrict__ b1, GemmLnParams ep) {
  constexpr int STAGES = ffn_stages<NC>();
  constexpr int D = 2 * NC;
  constexpr int KB1 = D / GEMM_BK;               // k-blocks of GEMM1
  constexpr uint32_t TM_OUT = 0, TM_H = NC;      // TMEM columns: OUT [0, NC), Hacc b at NC + 128 b
  constexpr uint32_t TMEM_COLS = 512;
  constexpr uint32_t G1_BYTES = GEMM_BM * 128 + FFN_HC * 128;   // Y k-block + W1 k-block
  constexpr uint32_t G2_BYTES = NC * 128;                        // W2 k-block

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sRing = smem;
  uint8_t* sH = sRing + STAGES * FFN_STAGE_BYTES;                 // 4 atoms: hidden columns [64 a, 64 a + 64) of the chunk
  uint8_t* sStage = sH + 4 * FFN_ATOM_BYTES;                      // 2 staging slots of the LayerNorm epilogue
  float2* s_stat = reinterpret_cast<float2*>(sStage + 2 * GEMM_STAGING_BYTES);   // [2 buffers][2 partials][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_stat + 2 * 2 * 128);
  uint64_t* full_bar = bars;                     // [STAGES]
  uint64_t* empty_bar = full_bar + STAGES;       // [STAGES]
  uint64_t* hacc_full = empty_bar + STAGES;      // [2]
  uint64_t* hacc_empty = hacc_full + 2;          // [2]
  uint64_t* h_full = hacc_empty + 2;             // [4]
  uint64_t* h_free = h_full + 4;                 // [4]
  uint64_t* out_full = h_free + 4;               // [1]
  uint64_t* out_empty = out_full + 1;            // [1]
  uint64_t* res_full = out_empty + 1;            // [2]
  uint64_t* stat_bar = res_full + 2;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stat_bar + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int m_tiles = (M + GEMM_BM - 1) / GEMM_BM;
  const int n_chunks = F / FFN_CHUNK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tma_y);
    tma_prefetch_desc(&tma_w1);
    tma_prefetch_desc(&tma_w2);
    tma_prefetch_desc(&tma_x);
    tma_prefetch_desc(&tma_yo);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&hacc_full[s], 1); mbar_init(&hacc_empty[s], 4); }
    for (int s = 0; s < 4; ++s) { mbar_init(&h_full[s], 1); mbar_init(&h_free[s], 2); }
    mbar_init(out_full, 1);
    mbar_init(out_empty, 4);
    for (int s = 0; s < 2; ++s) { mbar_init(&res_full[s], 1); mbar_init(&stat_bar[s], 8); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------ TMA producer: the MMA warp's consumption order
    uint32_t stage = 0, phase = 0;
    auto g1_loads = [&](int m_blk, int c) {
      for (int kb = 0; kb < KB1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* st = sRing + stage * FFN_STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], G1_BYTES);
          tma_load_2d(st, &tma_y, &full_bar[stage], kb * GEMM_BK, m_blk * GEMM_BM);
          tma_load_2d(st + GEMM_BM * 128, &tma_w1, &full_bar[stage], kb * GEMM_BK, c * FFN_CHUNK + int(rank) * FFN_HC);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    };
    auto g2_loads = [&](int c) {
      for (int a = 0; a < 4; ++a) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[stage], G2_BYTES);
          tma_load_2d(sRing + stage * FFN_STAGE_BYTES, &tma_w2, &full_bar[stage], c * FFN_CHUNK + a * GEMM_BK, int(rank) * NC);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    };
    for (int m_blk = cluster_id; m_blk < m_tiles; m_blk += num_clusters) {
      g1_loads(m_blk, 0);
      for (int c = 0; c < n_chunks; ++c) {
        if (c + 1 < n_chunks) g1_loads(m_blk, c + 1);
        g2_loads(c);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    constexpr uint32_t idesc1 = make_idesc_bf16(GEMM_BM, FFN_HC);
    constexpr uint32_t idesc2 = make_idesc_bf16(GEMM_BM, NC);
    const uint64_t dring = make_sw128_desc(smem_u32(sRing));
    const uint64_t dh = make_sw128_desc(smem_u32(sH));
    uint32_t stage = 0, phase = 0;
    auto g1 = [&](uint32_t g) {                 // GEMM1 of global chunk g into Hacc[g & 1]
      const uint32_t hb = g & 1u;
      mbar_wait(&hacc_empty[hb], ((g >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + TM_H + hb * FFN_HC;
      for (int kb = 0; kb < KB1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t da = dring + uint64_t((stage * FFN_STAGE_BYTES) >> 4);
          const uint64_t db = da + uint64_t((GEMM_BM * 128) >> 4);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) umma_bf16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc1, (kb | k) != 0);
          umma_commit(&empty_bar[stage]);
          if (kb == KB1 - 1) umma_commit(&hacc_full[hb]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    };
    auto g2 = [&](uint32_t g, int c, uint32_t it) {   // GEMM2 of chunk c (global chunk g) into OUT
      if (c == 0) {
        mbar_wait(out_empty, (it & 1u) ^ 1u);
        tc_fence_after();
      }
      const uint32_t tmem_d = tmem_base + TM_OUT;
      for (int a = 0; a < 4; ++a) {
        mbar_wait(&h_full[a], g & 1u);
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t da = dh + uint64_t((a * FFN_ATOM_BYTES) >> 4);
          const uint64_t db = dring + uint64_t((stage * FFN_STAGE_BYTES) >> 4);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) umma_bf16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc2, (c | a | k) != 0);
          umma_commit(&empty_bar[stage]);
          umma_commit_mc(&h_free[a], uint16_t(3));
          if (c == n_chunks - 1 && a == 3) umma_commit(out_full);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    };
    uint32_t it = 0;
    for (int m_blk = cluster_id; m_blk < m_tiles; m_blk += num_clusters, ++it) {
      const uint32_t g0 = it * uint32_t(n_chunks);
      g1(g0);
      for (int c = 0; c < n_chunks; ++c) {
        if (c + 1 < n_chunks) g1(g0 + c + 1);
        g2(g0 + c, c, it);
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------ SiLU epilogue: Hacc -> bf16 K-major atoms in both CTAs
    const int quad = warp & 3;
    const int trow = quad * 32 + lane;
    const bool issuer = warp == 2 && lane == 0;
    const uint32_t lane_addr = uint32_t(quad * 32) << 16;
    uint32_t it = 0;
    for (int m_blk = cluster_id; m_blk < m_tiles; m_blk += num_clusters, ++it) {
      for (int c = 0; c < n_chunks; ++c) {
        const uint32_t g = it * uint32_t(n_chunks) + uint32_t(c);
        const uint32_t hb = g & 1u;
        mbar_wait(&hacc_full[hb], (g >> 1) & 1u);
        tc_fence_after();
        const float* bias = b1 + c * FFN_CHUNK + int(rank) * FFN_HC;
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
          const int a = 2 * int(rank) + (j >> 1);
          uint8_t* atom = sH + a * FFN_ATOM_BYTES;
          uint32_t r[32];
          tmem_ld32(tmem_base + lane_addr + TM_H + hb * FFN_HC + 32 * j, r);
          float4 b[8];
#pragma unroll
          for (int qd = 0; qd < 8; ++qd) b[qd] = __ldg(reinterpret_cast<const float4*>(bias + 32 * j) + qd);
          if ((j & 1) == 0) mbar_wait(&h_free[a], (g & 1u) ^ 1u);   // GEMM2 of the previous chunk has read this atom in both CTAs
          tmem_ld_wait();
          uint32_t o[16];
#pragma unroll
          for (int qd = 0; qd < 8; ++qd) {
            o[2 * qd] = pack_bf16(silu_fast(__uint_as_float(r[4 * qd]) + b[qd].x), silu_fast(__uint_as_float(r[4 * qd + 1]) + b[qd].y));
            o[2 * qd + 1] = pack_bf16(silu_fast(__uint_as_float(r[4 * qd + 2]) + b[qd].z), silu_fast(__uint_as_float(r[4 * qd + 3]) + b[qd].w));
          }
#pragma unroll
          for (int qd = 0; qd < 4; ++qd)
            stage_store16(atom, trow, 4 * (j & 1) + qd, make_uint4(o[4 * qd], o[4 * qd + 1], o[4 * qd + 2], o[4 * qd + 3]));
          if (j & 1) {                                 // atom complete: publish it here and copy it to the peer
            fence_proxy_async();
            named_bar_sync(3, 128);
            if (issuer) {
              mbar_arrive(&h_full[a]);
              const uint32_t peer_bar = mapa_rank(smem_u32(&h_full[a]), rank ^ 1u);
              mbar_arrive_expect_tx_cluster(peer_bar, FFN_ATOM_BYTES);
              bulk_copy_to_cluster(mapa_rank(smem_u32(atom), rank ^ 1u), atom, FFN_ATOM_BYTES, peer_bar);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&hacc_empty[hb]);
      }
    }
  } else {
    // ------------------------------------------------ LayerNorm epilogue of the row block (OUT)
    const int quad = warp & 3;
    LnEpilogue<NC> le;
    le.tma_x = &tma_x; le.tma_r = &tma_r; le.tma_y = &tma_yo; le.ep = &ep;
    le.stg = sStage; le.rfull = res_full; le.stat_bar = stat_bar; le.s_stat = s_stat;
    le.n_part = 2; le.part_id = int(rank); le.rank = rank; le.bar_id = 1; le.trow = quad * 32 + lane;
    le.lane = lane; le.M = M; le.n_total = D; le.issuer = warp == 6 && lane == 0;
    const int gcol0 = int(rank) * NC;
    if (cluster_id < m_tiles) le.prefetch_first(gcol0, cluster_id * GEMM_BM, false);
    uint32_t it = 0;
    for (int m_blk = cluster_id; m_blk < m_tiles; m_blk += num_clusters, ++it) {
      mbar_wait(out_full, it & 1u);
      tc_fence_after();
      const int next_blk = m_blk + num_clusters;
      le.tile(tmem_base + (uint32_t(quad * 32) << 16) + TM_OUT, m_blk * GEMM_BM, gcol0, next_blk < m_tiles ? next_blk * GEMM_BM : -1);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(out_empty);
      if (next_blk < m_tiles) le.prefetch_first(gcol0, next_blk * GEMM_BM, true);
    }
    if (le.issuer) tma_store_wait_all();
  }

  tc_fence_before();
  cluster_sync_all();      // the peer may still copy atoms into this CTA / signal its barriers
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace cf
