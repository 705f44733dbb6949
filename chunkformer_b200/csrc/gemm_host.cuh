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
#include "gemm_ln.cuh"
#include "ffn_fused.cuh"

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
// 3-D tiled map {cols, rows, groups} of a row-major bf16 buffer: row pitch `ld` elements, group pitch `group_ld` elements.
// groups = 1, box_groups = 1 is the 2-D case behind a 3-D instruction; otherwise a box of box_rows x box_groups rows gathers /
// scatters the same rows of consecutive groups (128B swizzle, shared-memory image = box_groups x box_rows rows of 128 bytes).
inline bool make_tma_3d_bf16(CUtensorMap* map, const void* ptr, uint64_t cols, uint64_t rows, uint64_t groups, uint64_t ld,
                             uint64_t group_ld, uint32_t box_cols, uint32_t box_rows, uint32_t box_groups, std::string* err) {
  auto enc = get_tensor_map_encoder(err);
  if (!enc) return false;
  cuuint64_t dims[3] = {cols, rows, groups};
  cuuint64_t strides[2] = {ld * 2, group_ld * 2};
  cuuint32_t box[3] = {box_cols, box_rows, box_groups};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled (3-D) failed with CUresult " + std::to_string(int(r));
    return false;
  }
  return true;
}
inline bool make_tma_2d_bf16(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                             uint32_t box_rows, uint32_t box_cols, std::string* err) {
  return make_tma_2d(map, ptr, false, rows, cols, ld, box_rows, box_cols, err);
}

struct KernelTiming;

struct GemmLaunch {
  const void* A; long long lda;  // [M, K] bf16
  const void* B; long long ldb;  // [N, K] bf16 (nn.Linear weight layout)
  int M, N, K;
  int epi, act;
  void* out; long long ldo;      // bf16 [M, N] (EPI_BF16), bf16 [M, N/2] (EPI_GLU), fp32 [M, N] (EPI_F32); unused for EPI_ARGMAX
  GemmEpiParams ep;
  int variant = -1;              // -1 = by problem size, 0 = 1-CTA kernel, 1 = 2-CTA (cta_group::2) kernel
  // bf16 / GLU output scattered by row groups (GemmEpiParams::scatter_rows): `out` = row 0 of group 0, group pitch in elements
  int scatter_rows = 0, scatter_row0 = 0, scatter_group_rows = 0; long long scatter_groups = 0;
  KernelTiming* timing = nullptr;  // optional per-handle event timing of one kernel family (bench.py roofline)
  int family = 0;                // kernel family tag recorded with the timing (cf_kernel_family in the public header)
};

// Event timing of selected kernel families inside real steps (bench.py roofline: average launch duration of the dominant
// kernel over the timed region, measured on the launching stream).  Owned by a handle; `mask` = bit set of families to time.
struct KernelTiming {
  unsigned mask = 0;
  struct Rec { cudaEvent_t a, b; int family; };
  std::vector<Rec> pool;
  size_t used = 0;
  bool begin(int family, cudaStream_t st) {
    if (!(mask & (1u << family))) return false;
    if (used == pool.size()) { Rec r{}; cudaEventCreate(&r.a); cudaEventCreate(&r.b); pool.push_back(r); }
    pool[used].family = family;
    cudaEventRecord(pool[used].a, st);
    return true;
  }
  void end(cudaStream_t st) { cudaEventRecord(pool[used++].b, st); }
  ~KernelTiming() { for (auto& r : pool) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); } }
};

// cudaFuncAttributeMaxDynamicSharedMemorySize is per (kernel, device): opt in once for each pair.  (Keyed by the kernel's
// address: all instantiations of one kernel template share a function-pointer TYPE.)
inline bool ensure_smem_optin_ptr(const void* kern, size_t bytes, std::string* err, const char* what) {
  static std::mutex mu;
  static std::vector<std::pair<const void*, unsigned long long>> done;   // kernel -> bit per device ordinal (< 64)
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  unsigned long long* mask = nullptr;
  for (auto& kv : done)
    if (kv.first == kern) { mask = &kv.second; break; }
  if (mask && dev < 64 && ((*mask >> dev) & 1ull)) return true;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
  if (e != cudaSuccess) {
    if (err) *err = std::string("cudaFuncSetAttribute(") + what + "): " + cudaGetErrorString(e);
    return false;
  }
  if (dev < 64) {
    if (!mask) { done.push_back({kern, 0ull}); mask = &done.back().second; }
    *mask |= 1ull << dev;
  }
  return true;
}
template <typename Kern>
inline bool ensure_smem_optin(Kern kern, size_t bytes, std::string* err, const char* what) {
  return ensure_smem_optin_ptr(reinterpret_cast<const void*>(kern), bytes, err, what);
}

template <int EPI, int ACT>
inline bool launch_gemm_epi(const GemmLaunch& g, int num_sms, cudaStream_t stream, std::string* err) {
  int use2 = g.variant;
  // Measured on B200 (tools/bench_gemm_shapes.py, profiles/README.md round 2): the pair kernel (half the B-operand traffic per
  // CTA) wins on every large shape - FFN w_1 0.336 -> 0.309 ms, QKV 0.333 -> 0.292, GLU 0.151 -> 0.136, CTC 0.699 -> 0.666 -
  // except the K = 512 fp32 + residual epilogue (0.180 vs 0.198), since its accumulator hand-off to the leader CTA no longer
  // carries a cluster-scope release (it used to lose at K = 512).  Small M: a 256-row tile is mostly padding.
  if (use2 < 0) use2 = (g.M >= 2048 && (g.K >= 1024 || EPI != EPI_F32)) ? 1 : 0;
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
    if (f32) {
      if (!make_tma_2d(&tc, g.out, true, g.M, ocols, g.ldo, GEMM_BM, 32, err)) return false;
    } else if (g.scatter_rows > 0) {     // 128 accumulator rows = 128 / scatter_rows groups x scatter_rows rows
      if (GEMM_BM % g.scatter_rows != 0) { if (err) *err = "gemm: scatter_rows must divide 128"; return false; }
      if (!make_tma_3d_bf16(&tc, g.out, ocols, uint64_t(g.scatter_group_rows), uint64_t(g.scatter_groups), uint64_t(g.ldo),
                            uint64_t(g.scatter_group_rows) * uint64_t(g.ldo), 64, uint32_t(g.scatter_rows),
                            uint32_t(GEMM_BM / g.scatter_rows), err)) return false;
      ep.scatter_rows = g.scatter_rows; ep.scatter_row0 = g.scatter_row0;
    } else {
      if (!make_tma_2d(&tc, g.out, false, g.M, ocols, g.ldo, GEMM_BM, 64, err)) return false;
    }
  }
  tr = tc;
  ep.resid_tma = 0;
#ifdef CF_ABLATION
  { const char* dbg = getenv("CF_GEMM_DEBUG"); ep.debug = dbg ? atoi(dbg) : 0; }
  ep.raw_out = g.out; ep.raw_ldo = g.ldo;
#endif
  // residual through TMA when every epilogue group of every tile has four full rounds and the pitch is TMA-legal
  if (EPI == EPI_F32 && g.ep.resid != nullptr && g.N % GEMM_BN == 0 && (g.ep.ld_resid * 4) % 16 == 0 &&
      (reinterpret_cast<uintptr_t>(g.ep.resid) & 15) == 0) {
    if (!make_tma_2d(&tr, g.ep.resid, true, g.M, uint64_t(g.N), g.ep.ld_resid, GEMM_BM, 32, err)) return false;
    ep.resid_tma = 1;
  }
  if (use2) {
    auto kern2 = gemm2_tcgen05_kernel<EPI, ACT>;
    const size_t smem2 = gemm2_smem_bytes(EPI);
    if (!ensure_smem_optin(kern2, smem2, err, "gemm2")) return false;
    const int tiles2 = ((g.M + 255) / 256) * ((g.N + GEMM_BN - 1) / GEMM_BN);
    if (tiles2 == 0) return true;
    int clusters = num_sms / 2;
    if (clusters > tiles2) clusters = tiles2;
    const bool timed2 = g.timing && g.timing->begin(g.family, stream);
    kern2<<<2 * clusters, GEMM_THREADS, smem2, stream>>>(ta, tb, tc, tr, g.M, g.N, g.K, ep);
    if (timed2) g.timing->end(stream);
    ++g_kernel_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      if (err) *err = std::string("gemm2 launch: ") + cudaGetErrorString(e);
      return false;
    }
    return true;
  }
  auto kern = gemm_tcgen05_kernel<EPI, ACT>;
  const size_t smem = gemm_smem_bytes(EPI);
  if (!ensure_smem_optin(kern, smem, err, "gemm")) return false;
  const int m_tiles = (g.M + GEMM_BM - 1) / GEMM_BM, n_tiles = (g.N + GEMM_BN - 1) / GEMM_BN;
  const int tiles = m_tiles * n_tiles;
  if (tiles == 0) return true;
  const int grid = tiles < num_sms ? tiles : num_sms;
  const bool timed = g.timing && g.timing->begin(g.family, stream);
  kern<<<grid, GEMM_THREADS, smem, stream>>>(ta, tb, tc, tr, g.M, g.N, g.K, ep);
  if (timed) g.timing->end(stream);
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

// ---------------------------------------------------------------------------------------------------------------------
// residual GEMM + fused LayerNorm(s) on a CTA pair (gemm_ln.cuh)
// ---------------------------------------------------------------------------------------------------------------------
struct GemmLnLaunch {
  const void* A; long long lda;       // [M, K] bf16
  const void* B; long long ldb;       // [N, K] bf16, N = 256 or 512
  int M, N, K;
  const float* bias;                  // [N]
  const float* resid; long long ld_resid;   // fp32 [M, N] or nullptr
  float alpha;
  const int2* row_range; int rows_per_chunk;
  int mode;                           // GemmLnMode
  const float *ln1_w, *ln1_b, *ln2_w, *ln2_b;
  float* x_out; long long ldx;        // LNM_Y / LNM_XY: the fp32 residual stream; LNM_FINAL: fp32 result (nullable)
  void* y_out; long long ldy;         // bf16 [M, N] (LNM_FINAL: nullable)
  const int* row_limit; int rows_per_seq;
  KernelTiming* timing = nullptr; int family = 0;
  // A gathered by row groups (GemmLnParams::gather_rows; split / quad kernels only): A = row 0 of group 0
  int gather_rows = 0, gather_row0 = 0, gather_group_rows = 0; long long gather_groups = 0;
  int variant = -1;                   // 1: gemm_ln_split_kernel (normalisation passes on their own warps); 0: gemm_ln_kernel;
                                      // 2: gemm_ln_quad_kernel (cluster of four, cta_group::2 MMAs); -1: default (1)
#ifdef CF_ABLATION
  long long* prof = nullptr;
#endif
};

template <int NC>
inline bool launch_gemm_ln_nc(const GemmLnLaunch& g, int num_sms, cudaStream_t stream, std::string* err) {
  CUtensorMap ta, tb, tx, tr, ty;
  if (!make_tma_2d_bf16(&ta, g.A, g.M, g.K, g.lda, GEMM_BM, GEMM_BK, err)) return false;
  if (!make_tma_2d_bf16(&tb, g.B, g.N, g.K, g.ldb, NC, GEMM_BK, err)) return false;
  tx = ta; tr = ta; ty = ta;
  if (g.x_out && !make_tma_2d(&tx, g.x_out, true, g.M, uint64_t(g.N), g.ldx, GEMM_BM, 32, err)) return false;
  if (g.resid && !make_tma_2d(&tr, g.resid, true, g.M, uint64_t(g.N), g.ld_resid, GEMM_BM, 32, err)) return false;
  if (g.y_out && !make_tma_2d(&ty, g.y_out, false, g.M, uint64_t(g.N), g.ldy, GEMM_BM, 64, err)) return false;
  GemmLnParams ep;
  ep.bias = g.bias; ep.has_resid = g.resid != nullptr; ep.alpha = g.alpha; ep.row_range = g.row_range;
  ep.rows_per_chunk = g.rows_per_chunk > 0 ? g.rows_per_chunk : 1; ep.mode = g.mode;
  ep.ln1_w = g.ln1_w; ep.ln1_b = g.ln1_b; ep.ln2_w = g.ln2_w; ep.ln2_b = g.ln2_b;
  ep.row_limit = g.row_limit; ep.rows_per_seq = g.rows_per_seq > 0 ? g.rows_per_seq : 1;
  ep.store_f32 = g.x_out != nullptr; ep.store_bf16 = g.y_out != nullptr;
#ifdef CF_ABLATION
  ep.prof = g.prof;
  { const char* e = getenv("CF_LN_DEBUG"); ep.debug = e ? atoi(e) : 0; }
#endif
  const int m_tiles = (g.M + GEMM_BM - 1) / GEMM_BM;
  if (m_tiles == 0) return true;
  int clusters = num_sms / 2;
  if (clusters > m_tiles) clusters = m_tiles;
  // Default: the pair kernel.  The cluster of four is 9 % / 6 % faster alone at K = 2048 (0.430 -> 0.391 ms, 0.443 -> 0.417 with
  // two LayerNorms) but leaves 16 SMs idle, and inside the power-capped step the two are equal (67.9-68.6 vs 66.8-69.0 ms).
  const int variant = g.variant < 0 ? 1 : g.variant;
  if (variant != 0) {      // split / quad kernels can gather A through a 3-D map (see GemmLnParams::gather_rows)
    if (g.gather_rows > 0) {
      if (GEMM_BM % g.gather_rows != 0) { if (err) *err = "gemm_ln: gather_rows must divide 128"; return false; }
      if (!make_tma_3d_bf16(&ta, g.A, uint64_t(g.K), uint64_t(g.gather_group_rows), uint64_t(g.gather_groups), uint64_t(g.lda),
                            uint64_t(g.gather_group_rows) * uint64_t(g.lda), GEMM_BK, uint32_t(g.gather_rows),
                            uint32_t(GEMM_BM / g.gather_rows), err)) return false;
      ep.gather_rows = g.gather_rows; ep.gather_row0 = g.gather_row0;
    }
  } else if (g.gather_rows > 0) { if (err) *err = "gemm_ln: the first-version kernel cannot gather A"; return false; }
  if (variant == 2) {
    // cluster of four (two cta_group::2 pairs per 256-row block); the number of co-resident clusters is asked of the driver
    // (33 on a B200: GPC sizes) and the kernel is persistent over that many
    auto kern = gemm_ln_quad_kernel<NC>;
    const size_t smem = gemmln_quad_smem_bytes<NC>();
    if (!ensure_smem_optin(kern, smem, err, "gemm_ln_quad")) return false;
    if (!make_tma_2d_bf16(&tb, g.B, g.N, g.K, g.ldb, NC / 2, GEMM_BK, err)) return false;
    static std::mutex mu;
    static int max_clusters[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    int quads = 0;
    {
      std::lock_guard<std::mutex> lk(mu);
      if (dev >= 0 && dev < 64 && max_clusters[dev] == 0) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(unsigned(num_sms / 4 * 4)); cfg.blockDim = dim3(gemmln_split_threads<NC>()); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = num_sms / 4 - 4; }
        max_clusters[dev] = n > 0 ? n : 1;
      }
      quads = (dev >= 0 && dev < 64) ? max_clusters[dev] : num_sms / 4 - 4;
    }
    const int m_tiles4 = (g.M + 255) / 256;
    if (quads > m_tiles4) quads = m_tiles4;
    const bool timed = g.timing && g.timing->begin(g.family, stream);
    kern<<<4 * quads, gemmln_split_threads<NC>(), smem, stream>>>(ta, tb, tx, tr, g.M, g.K, ep, g.x_out, g.ldx,
                                                                 static_cast<__nv_bfloat16*>(g.y_out), g.ldy);
    if (timed) g.timing->end(stream);
  } else if (variant == 1) {
    auto kern = gemm_ln_split_kernel<NC>;
    const size_t smem = gemmln_split_smem_bytes<NC>();
    if (!ensure_smem_optin(kern, smem, err, "gemm_ln_split")) return false;
    const bool timed = g.timing && g.timing->begin(g.family, stream);
    kern<<<2 * clusters, gemmln_split_threads<NC>(), smem, stream>>>(ta, tb, tx, tr, g.M, g.K, ep, g.x_out, g.ldx,
                                                                    static_cast<__nv_bfloat16*>(g.y_out), g.ldy);
    if (timed) g.timing->end(stream);
  } else {
    auto kern = gemm_ln_kernel<NC>;
    const size_t smem = gemmln_smem_bytes<NC>();
    if (!ensure_smem_optin(kern, smem, err, "gemm_ln")) return false;
    const bool timed = g.timing && g.timing->begin(g.family, stream);
    kern<<<2 * clusters, GEMM_THREADS, smem, stream>>>(ta, tb, tx, tr, ty, g.M, g.K, ep);
    if (timed) g.timing->end(stream);
  }
  ++g_kernel_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (err) *err = std::string("gemm_ln launch: ") + cudaGetErrorString(e);
    return false;
  }
  return true;
}

inline bool launch_gemm_ln(const GemmLnLaunch& g, int num_sms, cudaStream_t stream, std::string* err) {
  if (g.K % GEMM_BK != 0 || g.K <= 0 || g.lda % 8 != 0 || g.ldb % 8 != 0) {
    if (err) *err = "gemm_ln: K must be a positive multiple of 64 and leading dimensions multiples of 8";
    return false;
  }
  if (g.mode != LNM_Y && g.mode != LNM_XY && g.mode != LNM_FINAL) { if (err) *err = "gemm_ln: unknown mode"; return false; }
  if ((g.mode != LNM_FINAL && (!g.x_out || !g.y_out)) || (g.mode == LNM_FINAL && !g.x_out && !g.y_out) || !g.bias || !g.ln1_w || !g.ln1_b ||
      (g.mode != LNM_Y && (!g.ln2_w || !g.ln2_b))) {
    if (err) *err = "gemm_ln: missing output or LayerNorm parameters";
    return false;
  }
  if ((g.x_out && (g.ldx * 4) % 16 != 0) || (g.resid && (g.ld_resid * 4) % 16 != 0) || (g.y_out && (g.ldy * 2) % 16 != 0)) {
    if (err) *err = "gemm_ln: leading dimensions must give 16-byte aligned rows";
    return false;
  }
  if (g.N == 512) return launch_gemm_ln_nc<256>(g, num_sms, stream, err);
  if (g.N == 256) return launch_gemm_ln_nc<128>(g, num_sms, stream, err);
  if (err) *err = "gemm_ln: N (= d_model) must be 256 or 512";
  return false;
}

// ---------------------------------------------------------------------------------------------------------------------
// fused feed-forward module: w_1 -> SiLU -> w_2 -> residual -> LayerNorm(s) (ffn_fused.cuh)
// ---------------------------------------------------------------------------------------------------------------------
struct FfnLaunch {
  const void* Y; long long ldy_in;     // [M, d] bf16 (LayerNorm output feeding the module)
  const void* W1; const float* b1;     // [F, d] bf16, [F]
  const void* W2; const float* b2;     // [d, F] bf16, [d]
  int M, d, F;
  const float* resid; long long ld_resid;   // fp32 [M, d]
  float alpha;                         // 0.5 (macaron half-step residual)
  int mode;                            // GemmLnMode
  const float *ln1_w, *ln1_b, *ln2_w, *ln2_b;
  float* x_out; long long ldx;
  void* y_out; long long ldy;
  KernelTiming* timing = nullptr; int family = 0;
};

inline bool ffn_fused_supported(int d, int F) { return (d == 256 || d == 512) && F > 0 && F % FFN_CHUNK == 0; }

template <int NC>
inline bool launch_ffn_fused_nc(const FfnLaunch& g, int num_sms, cudaStream_t stream, std::string* err) {
  CUtensorMap ty, tw1, tw2, tx, tr, tyo;
  if (!make_tma_2d_bf16(&ty, g.Y, g.M, g.d, g.ldy_in, GEMM_BM, GEMM_BK, err)) return false;
  if (!make_tma_2d_bf16(&tw1, g.W1, g.F, g.d, g.d, FFN_HC, GEMM_BK, err)) return false;
  if (!make_tma_2d_bf16(&tw2, g.W2, g.d, g.F, g.F, NC, GEMM_BK, err)) return false;
  tx = ty; tr = ty; tyo = ty;
  if (g.x_out && !make_tma_2d(&tx, g.x_out, true, g.M, uint64_t(g.d), g.ldx, GEMM_BM, 32, err)) return false;
  if (g.resid && !make_tma_2d(&tr, g.resid, true, g.M, uint64_t(g.d), g.ld_resid, GEMM_BM, 32, err)) return false;
  if (g.y_out && !make_tma_2d(&tyo, g.y_out, false, g.M, uint64_t(g.d), g.ldy, GEMM_BM, 64, err)) return false;
  GemmLnParams ep;
  ep.bias = g.b2; ep.has_resid = g.resid != nullptr; ep.alpha = g.alpha; ep.mode = g.mode;
  ep.ln1_w = g.ln1_w; ep.ln1_b = g.ln1_b; ep.ln2_w = g.ln2_w; ep.ln2_b = g.ln2_b;
  ep.store_f32 = g.x_out != nullptr; ep.store_bf16 = g.y_out != nullptr;
  auto kern = ffn_fused_kernel<NC>;
  const size_t smem = ffn_smem_bytes<NC>();
  if (!ensure_smem_optin(kern, smem, err, "ffn_fused")) return false;
  const int m_tiles = (g.M + GEMM_BM - 1) / GEMM_BM;
  if (m_tiles == 0) return true;
  int clusters = num_sms / 2;
  if (clusters > m_tiles) clusters = m_tiles;
  const bool timed = g.timing && g.timing->begin(g.family, stream);
  kern<<<2 * clusters, GEMM_THREADS, smem, stream>>>(ty, tw1, tw2, tx, tr, tyo, g.M, g.F, g.b1, ep);
  if (timed) g.timing->end(stream);
  ++g_kernel_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (err) *err = std::string("ffn_fused launch: ") + cudaGetErrorString(e);
    return false;
  }
  return true;
}

inline bool launch_ffn_fused(const FfnLaunch& g, int num_sms, cudaStream_t stream, std::string* err) {
  if (!ffn_fused_supported(g.d, g.F)) { if (err) *err = "ffn_fused: d_model must be 256 or 512 and linear_units a multiple of 256"; return false; }
  if (g.mode != LNM_Y && g.mode != LNM_XY && g.mode != LNM_FINAL) { if (err) *err = "ffn_fused: unknown mode"; return false; }
  if (!g.Y || !g.W1 || !g.W2 || !g.b1 || !g.b2 || !g.ln1_w || !g.ln1_b || (g.mode != LNM_Y && (!g.ln2_w || !g.ln2_b)) ||
      (g.mode != LNM_FINAL && (!g.x_out || !g.y_out)) || (g.mode == LNM_FINAL && !g.x_out && !g.y_out)) {
    if (err) *err = "ffn_fused: missing operand, output or LayerNorm parameters";
    return false;
  }
  if (g.ldy_in % 8 != 0 || (g.x_out && (g.ldx * 4) % 16 != 0) || (g.resid && (g.ld_resid * 4) % 16 != 0) || (g.y_out && (g.ldy * 2) % 16 != 0)) {
    if (err) *err = "ffn_fused: leading dimensions must give 16-byte aligned rows";
    return false;
  }
  return g.d == 512 ? launch_ffn_fused_nc<256>(g, num_sms, stream, err) : launch_ffn_fused_nc<128>(g, num_sms, stream, err);
}

}  // namespace cf
