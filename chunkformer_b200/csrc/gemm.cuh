// Persistent warp-specialised tcgen05 GEMM for sm_100a:  C[M,N] = A[M,K] * B[N,K]^T  (bf16 in, fp32 accumulate)
//
//   warp 0      : TMA producer  (cp.async.bulk.tensor 2-D tiles, 128B swizzle, 4-stage mbarrier ring)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x 256 x 16, accumulators in TMEM)
//   warps 2..5  : epilogue (tcgen05.ld -> registers -> fused bias / activation / GLU / residual / argmax -> global)
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
  EPI_QKV = 3,     // cols [0,d): Q -> (Q+u) at col, (Q+v) at d+col ; cols [d,3d): K,V at d+col     (bf16)
  EPI_ARGMAX = 4,  // per (row, n-tile): best logit, runner-up, index of best
};
enum GemmAct : int { ACT_NONE = 0, ACT_RELU = 1, ACT_SILU = 2 };

struct GemmEpiParams {
  const float* bias = nullptr;   // [N]
  const float* bias_u = nullptr; // EPI_QKV: pos_bias_u flattened [d]
  const float* bias_v = nullptr; // EPI_QKV: pos_bias_v flattened [d]
  void* out = nullptr;
  long long ldo = 0;             // leading dimension of out, in elements
  const float* resid = nullptr;  // EPI_F32
  long long ld_resid = 0;
  float alpha = 1.0f;
  int act = ACT_NONE;
  const int2* row_range = nullptr;  // EPI_F32: row valid iff range[row / rows_per_chunk].x <= row % rows_per_chunk < .y
  int rows_per_chunk = 1;
  int qkv_d = 0;
  float* part_best = nullptr;    // EPI_ARGMAX: [M, n_tiles]
  float* part_second = nullptr;
  int* part_index = nullptr;
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_STAGES = 4;
constexpr int GEMM_THREADS = 192;

template <int BN>
constexpr size_t gemm_smem_bytes() {
  return size_t(GEMM_STAGES) * (GEMM_BM * 128 + BN * 128) + 1024 /*align slack*/ + 256 /*barriers*/;
}

template <int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, int M, int N,
                    int K, GemmEpiParams ep) {
  static_assert(BN == 128 || BN == 256, "BN");
  constexpr uint32_t A_BYTES = GEMM_BM * 128;
  constexpr uint32_t B_BYTES = BN * 128;
  constexpr uint32_t TMEM_COLS = 2 * BN;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + GEMM_STAGES * A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + GEMM_STAGES * B_BYTES);
  uint64_t* full_bar = bars;                     // [STAGES]
  uint64_t* empty_bar = bars + GEMM_STAGES;      // [STAGES]
  uint64_t* tfull_bar = bars + 2 * GEMM_STAGES;  // [2]
  uint64_t* tempty_bar = tfull_bar + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (M + GEMM_BM - 1) / GEMM_BM;
  const int n_tiles = (N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int k_blocks = (K + GEMM_BK - 1) / GEMM_BK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int s = 0; s < GEMM_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], A_BYTES + B_BYTES);
          tma_load_2d(sA + stage * A_BYTES, &tma_a, &full_bar[stage], kb * GEMM_BK, m_blk * GEMM_BM);
          tma_load_2d(sB + stage * B_BYTES, &tma_b, &full_bar[stage], kb * GEMM_BK, n_blk * BN);
          if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(GEMM_BM, BN);
      uint32_t stage = 0, phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = make_sw128_desc(smem_u32(sA + stage * A_BYTES));
          const uint64_t db = make_sw128_desc(smem_u32(sB + stage * B_BYTES));
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            // advance 16 K-elements = 32 bytes inside the 128B swizzle atom: +2 in the (addr >> 4) field
            umma_bf16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);
          if (kb == k_blocks - 1) umma_commit(&tfull_bar[acc]);
          if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------ epilogue warps (TMEM lane quadrant = warp % 4)
    const int quad = warp & 3;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int row = m_blk * GEMM_BM + quad * 32 + lane;
      const bool row_ok = row < M;
      const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + acc * BN;

      bool keep = true;   // EPI_F32 row mask: masked rows get exactly +0 (masked_fill_, convolution.py:253)
      if (EPI == EPI_F32 && ep.row_range != nullptr && row_ok) {
        const int ch = row / ep.rows_per_chunk;
        const int rr = row - ch * ep.rows_per_chunk;
        const int2 rg = ep.row_range[ch];
        keep = (rr >= rg.x && rr < rg.y);
      }
      float best = -INFINITY, second = -INFINITY;
      int best_idx = 0;

#pragma unroll 1
      for (int cc = 0; cc < BN / 32; ++cc) {
        const int col0 = n_blk * BN + cc * 32;
        if (col0 >= N) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld32(taddr + cc * 32, r);
        tmem_ld_wait();
        if (EPI == EPI_BF16) {
          if (row_ok) {
            uint32_t o[16];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              float v0 = __uint_as_float(r[j]) + __ldg(ep.bias + col0 + j);
              float v1 = __uint_as_float(r[j + 1]) + __ldg(ep.bias + col0 + j + 1);
              if (ep.act == ACT_RELU) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
              if (ep.act == ACT_SILU) { v0 = silu(v0); v1 = silu(v1); }
              o[j >> 1] = pack_bf16(v0, v1);
            }
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.out) + (long long)row * ep.ldo + col0);
#pragma unroll
            for (int q = 0; q < 4; ++q) dst[q] = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
          }
        } else if (EPI == EPI_GLU) {
          if (row_ok) {
            uint32_t o[8];
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float a0 = __uint_as_float(r[j]) + __ldg(ep.bias + col0 + j);
              float g0 = __uint_as_float(r[j + 1]) + __ldg(ep.bias + col0 + j + 1);
              float a1 = __uint_as_float(r[j + 2]) + __ldg(ep.bias + col0 + j + 2);
              float g1 = __uint_as_float(r[j + 3]) + __ldg(ep.bias + col0 + j + 3);
              o[j >> 2] = pack_bf16(a0 * sigmoidf_(g0), a1 * sigmoidf_(g1));
            }
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.out) + (long long)row * ep.ldo + (col0 >> 1));
            dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
            dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
          }
        } else if (EPI == EPI_F32) {
          if (row_ok) {
            float* dst = reinterpret_cast<float*>(ep.out) + (long long)row * ep.ldo + col0;
            const float* rs = ep.resid ? ep.resid + (long long)row * ep.ld_resid + col0 : nullptr;
            const float sc = ep.alpha;
            if (col0 + 32 > N) {   // ragged last column block (e.g. vocab 5000): scalar, bounds-checked
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (col0 + j < N) {
                  float v = keep ? sc * (__uint_as_float(r[j]) + __ldg(ep.bias + col0 + j)) : 0.f;
                  if (rs) v += rs[j];
                  dst[j] = v;
                }
              }
            } else
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 v;
              v.x = sc * (__uint_as_float(r[j]) + __ldg(ep.bias + col0 + j));
              v.y = sc * (__uint_as_float(r[j + 1]) + __ldg(ep.bias + col0 + j + 1));
              v.z = sc * (__uint_as_float(r[j + 2]) + __ldg(ep.bias + col0 + j + 2));
              v.w = sc * (__uint_as_float(r[j + 3]) + __ldg(ep.bias + col0 + j + 3));
              if (!keep) v = make_float4(0.f, 0.f, 0.f, 0.f);
              if (rs) {
                const float4 x = *reinterpret_cast<const float4*>(rs + j);
                v.x += x.x; v.y += x.y; v.z += x.z; v.w += x.w;
              }
              *reinterpret_cast<float4*>(dst + j) = v;
            }
          }
        } else if (EPI == EPI_QKV) {
          if (row_ok) {
            __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(ep.out) + (long long)row * ep.ldo;
            if (col0 < ep.qkv_d) {
              uint32_t ou[16], ov[16];
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                const float q0 = __uint_as_float(r[j]) + __ldg(ep.bias + col0 + j);
                const float q1 = __uint_as_float(r[j + 1]) + __ldg(ep.bias + col0 + j + 1);
                ou[j >> 1] = pack_bf16(q0 + __ldg(ep.bias_u + col0 + j), q1 + __ldg(ep.bias_u + col0 + j + 1));
                ov[j >> 1] = pack_bf16(q0 + __ldg(ep.bias_v + col0 + j), q1 + __ldg(ep.bias_v + col0 + j + 1));
              }
              uint4* du = reinterpret_cast<uint4*>(orow + col0);
              uint4* dv = reinterpret_cast<uint4*>(orow + ep.qkv_d + col0);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                du[q] = make_uint4(ou[4 * q], ou[4 * q + 1], ou[4 * q + 2], ou[4 * q + 3]);
                dv[q] = make_uint4(ov[4 * q], ov[4 * q + 1], ov[4 * q + 2], ov[4 * q + 3]);
              }
            } else {
              uint32_t o[16];
#pragma unroll
              for (int j = 0; j < 32; j += 2)
                o[j >> 1] = pack_bf16(__uint_as_float(r[j]) + __ldg(ep.bias + col0 + j),
                                      __uint_as_float(r[j + 1]) + __ldg(ep.bias + col0 + j + 1));
              uint4* dst = reinterpret_cast<uint4*>(orow + ep.qkv_d + col0);
#pragma unroll
              for (int q = 0; q < 4; ++q) dst[q] = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
            }
          }
        } else if (EPI == EPI_ARGMAX) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = col0 + j;
            if (col < N) {
              const float v = __uint_as_float(r[j]) + __ldg(ep.bias + col);
              if (v > best) { second = best; best = v; best_idx = col; }
              else if (v > second) { second = v; }
            }
          }
        }
      }
      if (EPI == EPI_ARGMAX && row_ok) {
        const long long p = (long long)row * n_tiles + n_blk;
        ep.part_best[p] = best;
        ep.part_second[p] = second;
        ep.part_index[p] = best_idx;
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace cf
