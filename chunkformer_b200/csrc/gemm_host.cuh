// Host-side launcher for the tcgen05 GEMM: tensor-map encoding (driver entry point fetched at run time so the
// library links only against the static CUDA runtime and loads on a machine without a driver).
#pragma once
#include <cudaTypedefs.h>

#include <atomic>
#include <cstdlib>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "gemm.cuh"

namespace cf {

extern std::atomic<long long> g_kernel_launches;

inline PFN_cuTensorMapEncodeTiled_v12000 get_tensor_map_encoder(std::string* err) {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  static std::string init_err;
  std::call_once(once, [&]() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) {
      init_err = std::string("cuTensorMapEncodeTiled unavailable: ") + cudaGetErrorString(e);
      cudaGetLastError();
    } else {
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
  });
  if (!fn && err) *err = init_err;
  return fn;
}

// 2-D row-major tensor [rows, cols] with row pitch `ld` elements; box = [box_rows, box_cols]; 128B swizzle
// (box_cols * elem_bytes must be 128).
inline bool make_tma_2d(CUtensorMap* map, const void* ptr, bool is_f32, uint64_t rows, uint64_t cols, uint64_t ld,
                        uint32_t box_rows, uint32_t box_cols, std::string* err, bool swizzle128 = true, bool swizzle64 = false) {
  auto enc = get_tensor_map_encoder(err);
  if (!enc) return false;
  const size_t esz = is_f32 ? 4 : 2;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * esz};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : (swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(int(r));
    return false;
  }
  return true;
}
inline bool make_tma_2d_bf16(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                             uint32_t box_rows, uint32_t box_cols, std::string* err) {
  return make_tma_2d(map, ptr, false, rows, cols, ld, box_rows, box_cols, err);
}

struct GemmLaunch {
  const void* A; long long lda;  // [M, K] bf16
  const void* B; long long ldb;  // [N, K] bf16 (nn.Linear weight layout)
  int M, N, K;
  int epi, act;
  void* out; long long ldo;      // bf16 [M, N] (EPI_BF16), bf16 [M, N/2] (EPI_GLU), fp32 [M, N] (EPI_F32); unused for EPI_ARGMAX
  GemmEpiParams ep;
};

// Optional event timing of one GEMM family inside a real step (bench.py roofline: average duration of the dominant kernel
// over the timed region, on the launching stream).  timing_select(): epi * 16 + act of the family to time, or -1 = off.
struct GemmTiming {
  int select = -1;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pool;
  size_t used = 0;
};
inline GemmTiming& gemm_timing() { static GemmTiming t; return t; }

// 0 = 1-CTA kernel, 1 = 2-CTA (cta_group::2) kernel; set by launch_gemm from the problem size / CF_GEMM_2CTA.
inline int& gemm_variant_override() { static int v = -1; return v; }

template <int EPI, int ACT>
inline bool launch_gemm_epi(const GemmLaunch& g, int num_sms, cudaStream_t stream, std::string* err) {
  int use2 = gemm_variant_override();
  // Measured on B200 (tools/bench_gemm_shapes.py): the pair kernel wins when the K loop dominates (K >= 1024: FFN w_2,
  // embed.out); for K = 512 the epilogue dominates and the 1-CTA kernel is faster. Small M: a 256-row tile is mostly padding.
  if (use2 < 0) use2 = (g.M >= 2048 && g.K >= 1024) ? 1 : 0;
  CUtensorMap ta, tb, tc, tr;
  GemmEpiParams ep = g.ep;
  if (!make_tma_2d_bf16(&ta, g.A, g.M, g.K, g.lda, GEMM_BM, GEMM_BK, err)) return false;
  if (!make_tma_2d_bf16(&tb, g.B, g.N, g.K, g.ldb, use2 ? 128 : GEMM_BN, GEMM_BK, err)) return false;
  if (EPI == EPI_ARGMAX) {
    tc = ta;
  } else {
    const bool f32 = EPI == EPI_F32;
    const uint64_t ocols = EPI == EPI_GLU ? uint64_t(g.N / 2) : uint64_t(g.N);
    if (g.out == nullptr || (g.ldo * (f32 ? 4 : 2)) % 16 != 0) {
      if (err) *err = "gemm: output pointer missing or leading dimension not 16-byte aligned";
      return false;
    }
    // measured: FFN w_1 0.335 -> 0.330 ms, QKV 0.325 -> 0.325 ms (the exposed cost of the epilogue is the TMA store itself,
    // not the wait for the staging slot; profiles/README.md), so the simpler single-slot path stays the default
    static const int half_env = [] { const char* e = getenv("CF_GEMM_HALFSLOT"); return e ? atoi(e) : 0; }();
    ep.half_slot = (EPI == EPI_BF16 && !use2 && half_env) ? 1 : 0;
    if (ep.half_slot) {
      // 32-column (64-byte) store boxes: two half slots per staging tile, the store of one drains while the other is filled
      if (!make_tma_2d(&tc, g.out, false, g.M, ocols, g.ldo, GEMM_BM, 32, err, false, true)) return false;
    } else if (!make_tma_2d(&tc, g.out, f32, g.M, ocols, g.ldo, GEMM_BM, f32 ? 32 : 64, err)) return false;
  }
  tr = tc;
  ep.resid_tma = 0;
  { const char* dbg = getenv("CF_GEMM_DEBUG"); ep.debug = dbg ? atoi(dbg) : 0; }
  ep.raw_out = g.out; ep.raw_ldo = g.ldo;
  // residual through TMA when every epilogue group of every tile has four full rounds and the pitch is TMA-legal
  if (EPI == EPI_F32 && g.ep.resid != nullptr && g.N % GEMM_BN == 0 && (g.ep.ld_resid * 4) % 16 == 0 &&
      (reinterpret_cast<uintptr_t>(g.ep.resid) & 15) == 0) {
    if (!make_tma_2d(&tr, g.ep.resid, true, g.M, uint64_t(g.N), g.ep.ld_resid, GEMM_BM, 32, err)) return false;
    ep.resid_tma = 1;
  }
  if (use2) {
    auto kern2 = gemm2_tcgen05_kernel<EPI, ACT>;
    static bool attr2_set = false;
    const size_t smem2 = gemm2_smem_bytes(EPI);
    if (!attr2_set) {
      cudaError_t e = cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem2));
      if (e != cudaSuccess) {
        if (err) *err = std::string("cudaFuncSetAttribute(gemm2): ") + cudaGetErrorString(e);
        return false;
      }
      attr2_set = true;
    }
    const int tiles2 = ((g.M + 255) / 256) * ((g.N + GEMM_BN - 1) / GEMM_BN);
    if (tiles2 == 0) return true;
    int clusters = num_sms / 2;
    if (clusters > tiles2) clusters = tiles2;
    GemmTiming& tm2 = gemm_timing();
    const bool timed2 = tm2.select == EPI * 16 + ACT;
    if (timed2) {
      if (tm2.used == tm2.pool.size()) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); tm2.pool.push_back({a, b}); }
      cudaEventRecord(tm2.pool[tm2.used].first, stream);
    }
    kern2<<<2 * clusters, GEMM_THREADS, smem2, stream>>>(ta, tb, tc, tr, g.M, g.N, g.K, ep);
    if (timed2) cudaEventRecord(tm2.pool[tm2.used++].second, stream);
    ++g_kernel_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      if (err) *err = std::string("gemm2 launch: ") + cudaGetErrorString(e);
      return false;
    }
    return true;
  }
  auto kern = gemm_tcgen05_kernel<EPI, ACT>;
  static bool attr_set = false;
  const size_t smem = gemm_smem_bytes(EPI);
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) {
      if (err) *err = std::string("cudaFuncSetAttribute(gemm): ") + cudaGetErrorString(e);
      return false;
    }
    attr_set = true;
  }
  const int m_tiles = (g.M + GEMM_BM - 1) / GEMM_BM, n_tiles = (g.N + GEMM_BN - 1) / GEMM_BN;
  const int tiles = m_tiles * n_tiles;
  if (tiles == 0) return true;
  const int grid = tiles < num_sms ? tiles : num_sms;
  GemmTiming& tm = gemm_timing();
  const bool timed = tm.select == EPI * 16 + ACT;
  if (timed) {
    if (tm.used == tm.pool.size()) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); tm.pool.push_back({a, b}); }
    cudaEventRecord(tm.pool[tm.used].first, stream);
  }
  kern<<<grid, GEMM_THREADS, smem, stream>>>(ta, tb, tc, tr, g.M, g.N, g.K, ep);
  if (timed) cudaEventRecord(tm.pool[tm.used++].second, stream);
  ++g_kernel_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (err) *err = std::string("gemm launch: ") + cudaGetErrorString(e);
    return false;
  }
  return true;
}

inline bool launch_gemm(const GemmLaunch& g, int num_sms, cudaStream_t stream, std::string* err) {
  if (g.K % GEMM_BK != 0 || g.lda % 8 != 0 || g.ldb % 8 != 0) {
    if (err) *err = "gemm: K must be a multiple of 64 and leading dimensions multiples of 8";
    return false;
  }
  if (g.epi != EPI_ARGMAX && g.N % 8 != 0) {
    if (err) *err = "gemm: N must be a multiple of 8 for stored outputs";
    return false;
  }
  switch (g.epi) {
    case EPI_BF16:
      if (g.act == ACT_RELU) return launch_gemm_epi<EPI_BF16, ACT_RELU>(g, num_sms, stream, err);
      if (g.act == ACT_SILU) return launch_gemm_epi<EPI_BF16, ACT_SILU>(g, num_sms, stream, err);
      return launch_gemm_epi<EPI_BF16, ACT_NONE>(g, num_sms, stream, err);
    case EPI_GLU: return launch_gemm_epi<EPI_GLU, ACT_NONE>(g, num_sms, stream, err);
    case EPI_F32: return launch_gemm_epi<EPI_F32, ACT_NONE>(g, num_sms, stream, err);
    case EPI_ARGMAX: return launch_gemm_epi<EPI_ARGMAX, ACT_NONE>(g, num_sms, stream, err);
  }
  if (err) *err = "gemm: unknown epilogue";
  return false;
}

}  // namespace cf
