// C ABI of chunkformer_b200 (include/chunkformer_b200.h): handle, weight conversion, encoder driver, CTC head,
// kernel-level entry points.  One translation unit: all sm_100a kernels are header templates instantiated here.
#include "../../include/chunkformer_b200.h"

#include <atomic>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "attention_simt.cuh"
#include "attention_tc.cuh"
#include "fbank.cuh"
#include "frontend.cuh"
#include "gemm_host.cuh"
#include "misc_kernels.cuh"
#include "norm_conv.cuh"
#include "plan.h"
#include "transducer.cuh"

using namespace cf;
typedef __nv_bfloat16 bf16;

static thread_local std::string g_last_error = "";

namespace cf { std::atomic<long long> g_kernel_launches{0}; }

// --------------------------------------------------------------------------------------------------------------------
// handle
// --------------------------------------------------------------------------------------------------------------------
struct LayerW {
  bf16 *ffm_w1, *ffm_w2, *ff_w1, *ff_w2, *qkv_w, *o_w, *pos_w, *pw1_w, *pw2_w;
  float *ffm_b1, *ffm_b2, *ff_b1, *ff_b2, *qkv_b, *o_b, *pw1_b, *dw_w, *dw_b, *cn_w, *cn_b, *pw2_b;
  float *ln_ffm_w, *ln_ffm_b, *ln_mha_w, *ln_mha_b, *ln_conv_w, *ln_conv_b, *ln_ff_w, *ln_ff_b, *ln_fin_w, *ln_fin_b;
};

struct PosTable {  // projected relative-position tables, one per layer, cached per (c, l, r)
  int c, l, r, R, Rpad;
  bf16* dev = nullptr;  // [L][Rpad][d]
  unsigned long long last_use = 0;
};
constexpr size_t kMaxPosTables = 8;   // least-recently-used tables beyond this are freed (full attention makes one per length)

struct PinnedStage {   // pinned host staging for the per-call plan tables (async upload without a host synchronisation)
  void* host = nullptr; size_t bytes = 0; cudaEvent_t done = nullptr; bool in_flight = false;
};

struct cf_handle {
  cf_config cfg;
  int device = 0;
  int num_sms = 148;
  int F3 = 9;  // frequency bins after the three stride-2 convs
  std::string err;
  std::map<std::string, std::vector<float>> host_w;
  std::map<std::string, std::vector<int64_t>> host_shape;
  bool finalized = false;
  uint8_t* arena = nullptr;
  size_t arena_bytes = 0;
  std::vector<LayerW> layers;
  float *fe_wpack = nullptr, *fe_b3 = nullptr, *fe_dw2_w = nullptr, *fe_dw2_b = nullptr, *fe_b6 = nullptr, *fe_bout = nullptr;
  bf16 *fe_w3 = nullptr, *fe_w6 = nullptr, *fe_wout = nullptr;
  float *cmvn_mean = nullptr, *cmvn_istd = nullptr;
  float *after_w = nullptr, *after_b = nullptr;
  bf16* ctc_w = nullptr;
  float* ctc_b = nullptr;
  float* zeros = nullptr;  // max(N) zero floats (bias-free GEMMs)
  std::vector<PosTable> pos_tables;
  unsigned long long use_clock = 0;
  PinnedStage stage[4];
  int stage_next = 0;
  cf::KernelTiming timing;
  // per-handle options (cf_set_option)
  int opt_fused_layernorm = 1;   // LayerNorms fused into the epilogue of the residual GEMM in front of them (gemm_ln.cuh)
  int opt_ln_split = -1;         // ... 1: normalisation passes on their own warps (gemm_ln_split_kernel), 2: + cluster of four with
                                 // cta_group::2 MMAs (gemm_ln_quad_kernel), -1: default (1), 0: gemm_ln_kernel
  int opt_stream_compact = 1;    // multi-stream steps: row-wise kernels on the real chunk of every stream only
  long long last_out_rows = 0;   // rows of `out` the last cf_encode call wrote
  int opt_gemm_pair = -1;        // plain GEMMs: -1 = by shape (gemm_host.cuh), 0 = 1-CTA kernel, 1 = CTA-pair (cta_group::2) kernel
  int opt_ffn_slab_rows = 0;     // > 0: the two FFN GEMMs run slab by slab of this many rows, the hidden activation of a slab
                                 // (rows x F bf16) is produced and consumed while it is still in L2
  int opt_fused_ffn = 0;         // feed-forward modules as one kernel each, hidden activation kept on chip (ffn_fused.cuh)
  struct FbankTables { int sr = 0, bins = 0, flen = 0, fshift = 0; float* window = nullptr; float* mel_w = nullptr; int2* mel_rng = nullptr; int* mel_cnt = nullptr; };
  std::vector<FbankTables> fbank_tables;
  // feature-arrival events of the next cf_encode call (cf_encode_feature_events): rows < ev_rows[i] are present once ev[i] fires
  std::vector<int64_t> ev_rows;
  int streams_n = 0, streams_ph = 0, streams_adv = 0;   // multi-stream spec of the next cf_encode call (cf_encode_streams)
  std::vector<cudaEvent_t> ev;
};

static int fail(cf_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  g_last_error = msg;
  return code;
}
// Entry points switch to their handle's device (or, without a handle, to the device that owns the data) and restore the
// caller's current device on every exit path: a process that drives several GPUs (batch_decode(devices=...)) must not find
// its current device changed behind its back.
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (dev >= 0 && dev != prev) cudaSetDevice(dev);
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
static int device_of(const void* ptr) {       // device owning a device pointer, -1 if unknown
  cudaPointerAttributes at;
  if (ptr && cudaPointerGetAttributes(&at, ptr) == cudaSuccess && at.type == cudaMemoryTypeDevice) return at.device;
  cudaGetLastError();
  return -1;
}
#define CF_CUDA(h, call)                                                                              \
  do {                                                                                                \
    cudaError_t e__ = (call);                                                                         \
    if (e__ != cudaSuccess) {                                                                         \
      return fail(h, CF_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));               \
    }                                                                                                 \
  } while (0)

extern "C" long long cf_launch_count(void) { return cf::g_kernel_launches.load(); }
extern "C" int cf_set_option(cf_handle* h, const char* name, int value) {
  if (!h || !name) return fail(h, CF_ERR_INVALID, "cf_set_option: null argument");
  const std::string k(name);
  if (k == "fused_layernorm") { h->opt_fused_layernorm = value != 0; return CF_OK; }
  if (k == "ln_split") { h->opt_ln_split = value < 0 ? -1 : (value > 2 ? 2 : value); return CF_OK; }
  if (k == "stream_compact") { h->opt_stream_compact = value != 0; return CF_OK; }
  if (k == "gemm_pair") { h->opt_gemm_pair = value < 0 ? -1 : (value != 0); return CF_OK; }
  if (k == "fused_ffn") { h->opt_fused_ffn = value != 0; return CF_OK; }
  if (k == "ffn_slab_rows") { h->opt_ffn_slab_rows = value > 0 ? ((value + 127) / 128) * 128 : 0; return CF_OK; }
  return fail(h, CF_ERR_INVALID, "cf_set_option: unknown option " + k);
}
extern "C" int cf_kernel_timing_begin(cf_handle* h, unsigned family_mask) {
  if (!h) return fail(nullptr, CF_ERR_INVALID, "cf_kernel_timing_begin: null handle");
  h->timing.mask = family_mask;
  h->timing.used = 0;
  return CF_OK;
}
extern "C" int cf_kernel_timing_end(cf_handle* h, int family, double* total_ms, int* launches) {
  if (!h) return fail(nullptr, CF_ERR_INVALID, "cf_kernel_timing_end: null handle");
  cf::KernelTiming& t = h->timing;
  t.mask = 0;
  double tot = 0.0;
  int n = 0;
  for (size_t i = 0; i < t.used; ++i) {
    if (t.pool[i].family != family) continue;
    if (cudaEventSynchronize(t.pool[i].b) != cudaSuccess) return fail(h, CF_ERR_CUDA, "cf_kernel_timing_end: event synchronize failed");
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, t.pool[i].a, t.pool[i].b) != cudaSuccess) return fail(h, CF_ERR_CUDA, "cf_kernel_timing_end: elapsed time failed");
    tot += ms;
    ++n;
  }
  if (total_ms) *total_ms = tot;
  if (launches) *launches = n;
  return CF_OK;
}
extern "C" const char* cf_version(void) { return "chunkformer_b200 0.1.0 (sm_100a)"; }
extern "C" const char* cf_last_error(const cf_handle* h) { return h ? h->err.c_str() : g_last_error.c_str(); }

extern "C" int cf_create(const cf_config* cfg, int device, cf_handle** out) {
  if (!cfg || !out) return fail(nullptr, CF_ERR_INVALID, "cf_create: null argument");
  *out = nullptr;
  const int d = cfg->d_model;
  if (!(d == 256 || d == 512)) return fail(nullptr, CF_ERR_INVALID, "cf_create: d_model must be 256 or 512");
  if (cfg->heads <= 0 || d % cfg->heads != 0 || !((d / cfg->heads) == 64 || (d / cfg->heads) == 128))
    return fail(nullptr, CF_ERR_INVALID, "cf_create: d_model / heads must be 64 or 128");
  if (cfg->ffn <= 0 || cfg->ffn % 64 != 0) return fail(nullptr, CF_ERR_INVALID, "cf_create: ffn must be a multiple of 64");
  if (cfg->kernel != 15) return fail(nullptr, CF_ERR_INVALID, "cf_create: cnn_module_kernel must be 15");
  if (cfg->layers <= 0 || cfg->feat_dim < 15 || cfg->vocab < 0)
    return fail(nullptr, CF_ERR_INVALID, "cf_create: bad layers / feat_dim / vocab");
  if (cfg->conv_norm != 0 && cfg->conv_norm != 1) return fail(nullptr, CF_ERR_INVALID, "cf_create: conv_norm must be 0 (layer_norm) or 1 (batch_norm)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device) {
    cudaGetLastError();
    return fail(nullptr, CF_ERR_CUDA, "cf_create: no CUDA device (there is no CPU fallback)");
  }
  cudaDeviceProp prop;
  CF_CUDA(nullptr, cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(nullptr, CF_ERR_CUDA, "cf_create: device is not sm_100 (kernels are built for sm_100a only)");
  DeviceGuard guard(device);
  std::unique_ptr<cf_handle> h(new cf_handle());
  h->cfg = *cfg;
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  int f = cfg->feat_dim;
  for (int i = 0; i < 3; ++i) f = (f - 3) / 2 + 1;
  h->F3 = f;
  *out = h.release();
  return CF_OK;
}

extern "C" void cf_destroy(cf_handle* h) {
  if (!h) return;
  DeviceGuard guard(h->device);
  if (h->arena) cudaFree(h->arena);
  for (auto& t : h->pos_tables)
    if (t.dev) cudaFree(t.dev);
  for (auto& t : h->fbank_tables) { cudaFree(t.window); cudaFree(t.mel_w); cudaFree(t.mel_rng); cudaFree(t.mel_cnt); }
  for (auto& sgt : h->stage) {
    if (sgt.done) { cudaEventSynchronize(sgt.done); cudaEventDestroy(sgt.done); }
    if (sgt.host) cudaFreeHost(sgt.host);
  }
  delete h;
}

extern "C" int cf_load_tensor(cf_handle* h, const char* key, const void* data, int dtype, int ndim, const int64_t* shape) {
  if (!h || !key || !data || ndim < 0 || (ndim > 0 && !shape)) return fail(h, CF_ERR_INVALID, "cf_load_tensor: null argument");
  if (dtype != CF_F32) return fail(h, CF_ERR_INVALID, "cf_load_tensor: only fp32 host tensors are accepted");
  if (h->finalized) return fail(h, CF_ERR_STATE, "cf_load_tensor: weights already finalized");
  const std::string k(key);
  if (k.rfind("encoder.", 0) != 0 && k.rfind("ctc.ctc_lo.", 0) != 0) return CF_OK;  // strict=False
  int64_t n = 1;
  std::vector<int64_t> shp;
  for (int i = 0; i < ndim; ++i) { n *= shape[i]; shp.push_back(shape[i]); }
  const float* src = static_cast<const float*>(data);
  h->host_w[k].assign(src, src + n);
  h->host_shape[k] = shp;
  return CF_OK;
}

namespace {

struct ArenaBuilder {
  std::vector<uint8_t> host;
  std::vector<std::pair<void**, size_t>> fixups;
  size_t add_raw(const void* src, size_t bytes, void** target) {
    size_t off = (host.size() + 255) & ~size_t(255);
    host.resize(off + bytes);
    if (src) memcpy(host.data() + off, src, bytes);
    else memset(host.data() + off, 0, bytes);
    fixups.push_back({target, off});
    return off;
  }
  template <typename T> void f32(const std::vector<float>& v, T** target) {
    add_raw(v.data(), v.size() * sizeof(float), reinterpret_cast<void**>(target));
  }
  template <typename T> void b16(const std::vector<float>& v, T** target) {
    std::vector<uint16_t> t(v.size());
    for (size_t i = 0; i < v.size(); ++i) {
      __nv_bfloat16 b = __float2bfloat16_rn(v[i]);
      memcpy(&t[i], &b, 2);
    }
    add_raw(t.data(), t.size() * 2, reinterpret_cast<void**>(target));
  }
};

}  // namespace

extern "C" int cf_finalize_weights(cf_handle* h) {
  if (!h) return fail(nullptr, CF_ERR_INVALID, "cf_finalize_weights: null handle");
  if (h->finalized) return CF_OK;
  const int d = h->cfg.d_model, F = h->cfg.ffn, L = h->cfg.layers, KW = h->cfg.kernel, V = h->cfg.vocab;
  const int F3 = h->F3;
  std::string missing;
  auto get = [&](const std::string& key, int64_t expect) -> const std::vector<float>& {
    static const std::vector<float> empty;
    auto it = h->host_w.find(key);
    if (it == h->host_w.end() || int64_t(it->second.size()) != expect) {
      if (missing.empty()) missing = key + (it == h->host_w.end() ? " (absent)" : " (wrong size)");
      return empty;
    }
    return it->second;
  };
  ArenaBuilder ab;
  h->layers.assign(L, LayerW());
  // ---- front-end
  {
    const auto& w0 = get("encoder.embed.conv.0.weight", int64_t(d) * 9);
    const auto& b0 = get("encoder.embed.conv.0.bias", d);
    const auto& w2 = get("encoder.embed.conv.2.weight", int64_t(d) * 9);
    const auto& b2 = get("encoder.embed.conv.2.bias", d);
    const auto& w3 = get("encoder.embed.conv.3.weight", int64_t(d) * d);
    const auto& b3 = get("encoder.embed.conv.3.bias", d);
    const auto& w5 = get("encoder.embed.conv.5.weight", int64_t(d) * 9);
    const auto& b5 = get("encoder.embed.conv.5.bias", d);
    const auto& w6 = get("encoder.embed.conv.6.weight", int64_t(d) * d);
    const auto& b6 = get("encoder.embed.conv.6.bias", d);
    const auto& wo = get("encoder.embed.out.weight", int64_t(d) * d * F3);
    const auto& bo = get("encoder.embed.out.bias", d);
    if (!missing.empty()) return fail(h, CF_ERR_STATE, "cf_finalize_weights: missing tensor " + missing);
    std::vector<float> pack(size_t(d) * 20), dw2(size_t(9) * d), wperm(wo.size());
    for (int ch = 0; ch < d; ++ch) {
      for (int t = 0; t < 9; ++t) { pack[ch * 20 + t] = w0[ch * 9 + t]; pack[ch * 20 + 10 + t] = w2[ch * 9 + t]; }
      pack[ch * 20 + 9] = b0[ch];
      pack[ch * 20 + 19] = b2[ch];
      for (int t = 0; t < 9; ++t) dw2[t * d + ch] = w5[ch * 9 + t];
    }
    // embed.out consumes features ordered (channel, freq) (subsampling.py:163-164); our rows are (freq, channel)
    for (int o = 0; o < d; ++o)
      for (int ch = 0; ch < d; ++ch)
        for (int fq = 0; fq < F3; ++fq) wperm[size_t(o) * d * F3 + size_t(fq) * d + ch] = wo[size_t(o) * d * F3 + size_t(ch) * F3 + fq];
    ab.f32(pack, &h->fe_wpack);
    ab.b16(w3, &h->fe_w3); ab.f32(b3, &h->fe_b3);
    ab.f32(dw2, &h->fe_dw2_w); ab.f32(b5, &h->fe_dw2_b);
    ab.b16(w6, &h->fe_w6); ab.f32(b6, &h->fe_b6);
    ab.b16(wperm, &h->fe_wout); ab.f32(bo, &h->fe_bout);
    if (h->cfg.has_cmvn) {
      const auto& mean = get("encoder.global_cmvn.mean", h->cfg.feat_dim);
      const auto& istd = get("encoder.global_cmvn.istd", h->cfg.feat_dim);
      if (!missing.empty()) return fail(h, CF_ERR_STATE, "cf_finalize_weights: missing tensor " + missing);
      ab.f32(mean, &h->cmvn_mean); ab.f32(istd, &h->cmvn_istd);
    }
  }
  // ---- layers
  for (int i = 0; i < L; ++i) {
    const std::string p = "encoder.encoders." + std::to_string(i) + ".";
    LayerW& w = h->layers[i];
    auto lin = [&](const std::string& name, int64_t out_f, int64_t in_f, bf16** wt, float** bt) {
      const auto& W = get(p + name + ".weight", out_f * in_f);
      if (bt) { const auto& B = get(p + name + ".bias", out_f); if (missing.empty()) ab.f32(B, bt); }
      if (missing.empty()) ab.b16(W, wt);
    };
    auto nrm = [&](const std::string& name, float** wt, float** bt) {
      const auto& W = get(p + name + ".weight", d);
      const auto& B = get(p + name + ".bias", d);
      if (missing.empty()) { ab.f32(W, wt); ab.f32(B, bt); }
    };
    lin("feed_forward_macaron.w_1", F, d, &w.ffm_w1, &w.ffm_b1);
    lin("feed_forward_macaron.w_2", d, F, &w.ffm_w2, &w.ffm_b2);
    lin("feed_forward.w_1", F, d, &w.ff_w1, &w.ff_b1);
    lin("feed_forward.w_2", d, F, &w.ff_w2, &w.ff_b2);
    {
      const auto& wq = get(p + "self_attn.linear_q.weight", int64_t(d) * d);
      const auto& wk = get(p + "self_attn.linear_k.weight", int64_t(d) * d);
      const auto& wv = get(p + "self_attn.linear_v.weight", int64_t(d) * d);
      const auto& bq = get(p + "self_attn.linear_q.bias", d);
      const auto& bk = get(p + "self_attn.linear_k.bias", d);
      const auto& bv = get(p + "self_attn.linear_v.bias", d);
      const auto& bu = get(p + "self_attn.pos_bias_u", d);
      const auto& bvv = get(p + "self_attn.pos_bias_v", d);
      if (missing.empty()) {
        // fused projection with output columns [Q+u | Q+v | K | V]: W_q appears twice, pos_bias_u / pos_bias_v
        // (attention.py:486-488) are folded into the two Q biases.
        std::vector<float> W(wq); W.insert(W.end(), wq.begin(), wq.end());
        W.insert(W.end(), wk.begin(), wk.end()); W.insert(W.end(), wv.begin(), wv.end());
        // The two Q blocks also absorb the softmax scale in the log2 domain, (1/sqrt(d_k)) * log2(e) (attention.py:503),
        // so the attention kernel adds S_ac + S_bd and feeds exp2 directly.
        const float qs = (1.0f / sqrtf(float(d / h->cfg.heads))) * 1.4426950408889634f;
        for (size_t e = 0; e < size_t(2) * d * d; ++e) W[e] *= qs;
        std::vector<float> B(size_t(4) * d);
        for (int c = 0; c < d; ++c) { B[c] = (bq[c] + bu[c]) * qs; B[d + c] = (bq[c] + bvv[c]) * qs; B[2 * d + c] = bk[c]; B[3 * d + c] = bv[c]; }
        ab.b16(W, &w.qkv_w); ab.f32(B, &w.qkv_b);
      }
    }
    lin("self_attn.linear_out", d, d, &w.o_w, &w.o_b);
    lin("self_attn.linear_pos", d, d, &w.pos_w, nullptr);
    {
      // pointwise_conv1 (2d, d, 1): interleave value / gate rows so GLU pairs sit in adjacent accumulator columns
      const auto& W = get(p + "conv_module.pointwise_conv1.weight", int64_t(2) * d * d);
      const auto& B = get(p + "conv_module.pointwise_conv1.bias", 2 * d);
      if (missing.empty()) {
        std::vector<float> Wi(W.size()), Bi(B.size());
        for (int c = 0; c < d; ++c) {
          memcpy(&Wi[size_t(2 * c) * d], &W[size_t(c) * d], d * sizeof(float));
          memcpy(&Wi[size_t(2 * c + 1) * d], &W[size_t(d + c) * d], d * sizeof(float));
          Bi[2 * c] = B[c]; Bi[2 * c + 1] = B[d + c];
        }
        ab.b16(Wi, &w.pw1_w); ab.f32(Bi, &w.pw1_b);
      }
    }
    {
      const auto& W = get(p + "conv_module.depthwise_conv.weight", int64_t(d) * KW);
      const auto& B = get(p + "conv_module.depthwise_conv.bias", d);
      if (h->cfg.conv_norm == 1) {
        // eval-mode BatchNorm1d (convolution.py:83-89, 245-248) is a per-channel affine on the depthwise conv output:
        // a (conv + b_d) + s with a = gamma / sqrt(running_var + eps), s = beta - a * running_mean; folded into the taps and bias
        const auto& G = get(p + "conv_module.norm.weight", d);
        const auto& Bt = get(p + "conv_module.norm.bias", d);
        const auto& Mu = get(p + "conv_module.norm.running_mean", d);
        const auto& Var = get(p + "conv_module.norm.running_var", d);
        if (missing.empty()) {
          std::vector<float> Wf(W), Bf(B), ones(d, 1.0f), zeros(d, 0.0f);
          for (int ch = 0; ch < d; ++ch) {
            const float a = G[ch] / sqrtf(Var[ch] + 1e-5f);
            for (int t = 0; t < KW; ++t) Wf[size_t(ch) * KW + t] *= a;
            Bf[ch] = a * (B[ch] - Mu[ch]) + Bt[ch];
          }
          ab.f32(Wf, &w.dw_w); ab.f32(Bf, &w.dw_b); ab.f32(ones, &w.cn_w); ab.f32(zeros, &w.cn_b);
        }
      } else if (missing.empty()) { ab.f32(W, &w.dw_w); ab.f32(B, &w.dw_b); }
    }
    if (h->cfg.conv_norm != 1) nrm("conv_module.norm", &w.cn_w, &w.cn_b);
    lin("conv_module.pointwise_conv2", d, d, &w.pw2_w, &w.pw2_b);
    nrm("norm_ff_macaron", &w.ln_ffm_w, &w.ln_ffm_b);
    nrm("norm_mha", &w.ln_mha_w, &w.ln_mha_b);
    nrm("norm_conv", &w.ln_conv_w, &w.ln_conv_b);
    nrm("norm_ff", &w.ln_ff_w, &w.ln_ff_b);
    nrm("norm_final", &w.ln_fin_w, &w.ln_fin_b);
    if (!missing.empty()) return fail(h, CF_ERR_STATE, "cf_finalize_weights: missing tensor " + missing);
  }
  {
    const auto& W = get("encoder.after_norm.weight", d);
    const auto& B = get("encoder.after_norm.bias", d);
    if (!missing.empty()) return fail(h, CF_ERR_STATE, "cf_finalize_weights: missing tensor " + missing);
    ab.f32(W, &h->after_w); ab.f32(B, &h->after_b);
  }
  if (V > 0) {
    const auto& W = get("ctc.ctc_lo.weight", int64_t(V) * d);
    const auto& B = get("ctc.ctc_lo.bias", V);
    if (!missing.empty()) return fail(h, CF_ERR_STATE, "cf_finalize_weights: missing tensor " + missing);
    ab.b16(W, &h->ctc_w); ab.f32(B, &h->ctc_b);
  }
  {
    std::vector<float> z(size_t(std::max(std::max(4 * d, F), 2 * d)), 0.f);
    ab.f32(z, &h->zeros);
  }
  DeviceGuard guard(h->device);
  CF_CUDA(h, cudaMalloc(&h->arena, ab.host.size()));
  h->arena_bytes = ab.host.size();
  CF_CUDA(h, cudaMemcpy(h->arena, ab.host.data(), ab.host.size(), cudaMemcpyHostToDevice));
  for (auto& fx : ab.fixups) *fx.first = h->arena + fx.second;
  h->host_w.clear();
  h->host_shape.clear();
  h->finalized = true;
  return CF_OK;
}

// --------------------------------------------------------------------------------------------------------------------
// plan
// --------------------------------------------------------------------------------------------------------------------
extern "C" int cf_plan_create(int c, int l, int r, int kernel, int B, const int32_t* lens, const int32_t* offsets,
                              const int64_t* feat_row_offsets, cf_plan** out) {
  if (!lens || !out) return fail(nullptr, CF_ERR_INVALID, "cf_plan_create: null argument");
  std::unique_ptr<cf_plan> p(new cf_plan());
  std::string err;
  if (!cfplan::build_masked(p.get(), c, l, r, kernel, B, lens, offsets, feat_row_offsets, &err))
    return fail(nullptr, CF_ERR_INVALID, err);
  *out = p.release();
  return CF_OK;
}
extern "C" int cf_plan_create_padded(int c, int l, int r, int kernel, int B, int T, const int32_t* lens, cf_plan** out) {
  if (!lens || !out) return fail(nullptr, CF_ERR_INVALID, "cf_plan_create_padded: null argument");
  std::unique_ptr<cf_plan> p(new cf_plan());
  std::string err;
  if (!cfplan::build_padded(p.get(), c, l, r, kernel, B, T, lens, &err)) return fail(nullptr, CF_ERR_INVALID, err);
  *out = p.release();
  return CF_OK;
}
extern "C" void cf_plan_destroy(cf_plan* p) { delete p; }
extern "C" int cf_plan_num_chunks(const cf_plan* p) { return p ? p->n : 0; }
extern "C" int cf_plan_rows(const cf_plan* p) { return p ? p->n * p->c : 0; }
extern "C" int cf_plan_tables(const cf_plan* p, int32_t* n_chunks_out, int32_t* enc_lens_out) {
  if (!p) return fail(nullptr, CF_ERR_INVALID, "cf_plan_tables: null plan");
  for (int u = 0; u < p->B; ++u) {
    if (n_chunks_out) n_chunks_out[u] = p->n_chunks[u];
    if (enc_lens_out) enc_lens_out[u] = p->enc_lens[u];
  }
  return CF_OK;
}
extern "C" int cf_plan_masks(const cf_plan* p, uint8_t* att_mask, uint8_t* conv_mask) {
  if (!p) return fail(nullptr, CF_ERR_INVALID, "cf_plan_masks: null plan");
  const int W = p->l + p->c + p->r, CW = p->c + 2 * p->lorder;
  for (int g = 0; g < p->n; ++g) {
    const cf_chunk_entry& e = p->chunks[g];
    if (att_mask)
      for (int q = 0; q < W; ++q) att_mask[size_t(g) * W + q] = (q >= e.att_lo && q < e.att_hi) ? 1 : 0;
    if (conv_mask)
      for (int q = 0; q < CW; ++q) conv_mask[size_t(g) * CW + q] = (q >= e.conv_lo && q < e.conv_hi) ? 1 : 0;
  }
  return CF_OK;
}
extern "C" int cf_plan_chunk_table(const cf_plan* p, int32_t* table) {
  if (!p || !table) return fail(nullptr, CF_ERR_INVALID, "cf_plan_chunk_table: null argument");
  static_assert(sizeof(cf_chunk_entry) == 32, "chunk entry layout");
  memcpy(table, p->chunks.data(), size_t(p->n) * sizeof(cf_chunk_entry));
  return CF_OK;
}

// --------------------------------------------------------------------------------------------------------------------
// kernel dispatch helpers
// --------------------------------------------------------------------------------------------------------------------
namespace {

bool run_layernorm(int mode, int d, const LnParams& p, cudaStream_t st, std::string* err) {
  if (p.rows == 0) return true;
  const int warps = 8;
  const unsigned grid = unsigned((p.rows + warps - 1) / warps);
#define CF_LN(D, MODE) (++cf::g_kernel_launches, layernorm_kernel<D, MODE><<<grid, warps * 32, 0, st>>>(p))
  if (d == 512) { if (mode == 0) CF_LN(512, 0); else if (mode == 1) CF_LN(512, 1); else CF_LN(512, 2); }
  else if (d == 256) { if (mode == 0) CF_LN(256, 0); else if (mode == 1) CF_LN(256, 1); else CF_LN(256, 2); }
  else { *err = "layernorm: d must be 256 or 512"; return false; }
#undef CF_LN
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { *err = std::string("layernorm launch: ") + cudaGetErrorString(e); return false; }
  return true;
}

template <int D>
bool run_dwconv_d(const DwConvParams& p, cudaStream_t st, std::string* err) {
  const int c = p.c;
#define CF_DW(FG) (++cf::g_kernel_launches, dwconv_ln_silu_kernel<D, 15, FG><<<p.n_chunks * (c / FG), D / 2, 0, st>>>(p))
  if (c % 32 == 0) CF_DW(32);
  else if (c % 16 == 0) CF_DW(16);
  else if (c % 8 == 0) CF_DW(8);
  else if (c % 4 == 0) CF_DW(4);
  else if (c % 2 == 0) CF_DW(2);
  else CF_DW(1);
#undef CF_DW
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { *err = std::string("dwconv launch: ") + cudaGetErrorString(e); return false; }
  return true;
}
template <int D, int FG, int STAGES>
bool run_dwconv_tma(const DwConvParams& p, long long g_rows, int num_sms, cudaStream_t st, std::string* err) {
  CUtensorMap tm;
  if (!make_tma_2d(&tm, p.g, false, uint64_t(g_rows), uint64_t(D), uint64_t(D), FG + 14, 256, err, /*swizzle128=*/false)) return false;
  const size_t smem = dwconv_tma_smem_bytes<D, FG, STAGES>();
  if (!ensure_smem_optin(dwconv_ln_silu_tma_kernel<D, FG, STAGES>, smem, err, "dwconv")) return false;
  constexpr int CTAS = (FG == 16 && STAGES == 2) ? 3 : 2;
  const int groups = p.n_chunks * (p.c / FG);
  const int grid = groups < CTAS * num_sms ? groups : CTAS * num_sms;
  dwconv_ln_silu_tma_kernel<D, FG, STAGES><<<grid, D / 2, smem, st>>>(tm, p, groups);
  ++cf::g_kernel_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { *err = std::string("dwconv launch: ") + cudaGetErrorString(e); return false; }
  return true;
}

// g_rows: rows of the GLU buffer that may be read (>= n_chunks * c + kernel - 1); 0 = use the CUDA-core (non-TMA) kernel
bool run_dwconv(int d, int kernel, const DwConvParams& p, cudaStream_t st, std::string* err, long long g_rows = 0, int num_sms = 148) {
  if (kernel != 15) { *err = "dwconv: kernel must be 15"; return false; }
  if (p.n_chunks == 0) return true;
  // 32-frame groups, two-stage TMA ring, two CTAs per SM: the fastest of the variants measured in round 1 (16-frame groups
  // with three CTAs: 0.129 ms; three stages: 0.140 ms; this one: 0.127 ms; profiles/README.md)
  if (g_rows > 0 && p.c % 32 == 0) {
    if (d == 512) return run_dwconv_tma<512, 32, 2>(p, g_rows, num_sms, st, err);
    if (d == 256) return run_dwconv_tma<256, 32, 2>(p, g_rows, num_sms, st, err);
  }
  if (d == 512) return run_dwconv_d<512>(p, st, err);
  if (d == 256) return run_dwconv_d<256>(p, st, err);
  *err = "dwconv: d must be 256 or 512";
  return false;
}

bool run_frontend_conv(int impl, int d, const Fe1Params& f1, int num_sms, cudaStream_t st, std::string* err) {
  if (f1.n_chunks == 0) return true;
  if (f1.feat_dim > 80 || f1.F2 < 16) { *err = "frontend: feat_dim must be <= 80 with at least 16 bins after two convs"; return false; }
  const int per_chunk = f1.T2 * f1.F2;
  const int bpc = (per_chunk + 127) / 128;
  cudaError_t e = cudaSuccess;
  if (impl == 0) {
    const size_t smem = (size_t(d) * 20 + size_t(39) * f1.feat_dim) * sizeof(float) + 128 * 33 * sizeof(uint32_t);
    if (d == 512) {
      e = cudaFuncSetAttribute(frontend_conv0_dw1_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
      if (e == cudaSuccess) frontend_conv0_dw1_kernel<512><<<f1.n_chunks * bpc, 128, smem, st>>>(f1);
    } else if (d == 256) {
      e = cudaFuncSetAttribute(frontend_conv0_dw1_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
      if (e == cudaSuccess) frontend_conv0_dw1_kernel<256><<<f1.n_chunks * bpc, 128, smem, st>>>(f1);
    } else { *err = "frontend: d must be 256 or 512"; return false; }
  } else if (impl == 2) {
    // channel-major kernel: one unit per output time row, needs the 80-bin geometry (39 / 19 bins after conv0 / dw1)
    if (f1.feat_dim != 80) { *err = "frontend: the channel-major kernel needs feat_dim == 80"; return false; }
    const int units = f1.n_chunks * f1.T2;
    const int grid = units < num_sms ? units : num_sms;
    if (d == 512) {
      const size_t smem = frontend_cm_smem_bytes<512>();
      e = cudaFuncSetAttribute(frontend_conv0_dw1_cm_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
      if (e == cudaSuccess) frontend_conv0_dw1_cm_kernel<512><<<grid, fecm_threads<512>(), smem, st>>>(f1, units);
    } else if (d == 256) {
      const size_t smem = frontend_cm_smem_bytes<256>();
      e = cudaFuncSetAttribute(frontend_conv0_dw1_cm_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
      if (e == cudaSuccess) frontend_conv0_dw1_cm_kernel<256><<<grid, fecm_threads<256>(), smem, st>>>(f1, units);
    } else { *err = "frontend: d must be 256 or 512"; return false; }
  } else { *err = "frontend: unknown implementation"; return false; }
  ++cf::g_kernel_launches;
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { *err = std::string("frontend conv launch: ") + cudaGetErrorString(e); return false; }
  return true;
}

// 128-key-block kernel with a resident position-table slice: d_k = 64, tiles of 128 / c chunks, short windows
bool attention_tc_fast_supported(int c, int l, int r, int dk) {
  if (dk != 64 || !(c == 8 || c == 16 || c == 32 || c == 64) || l < 0 || r < 0 || l + r > 256) return false;
  const int W = l + c + r, nb = (l + 128 + r + 127) / 128;
  const int n_last = (W - 128 * (nb - 1) <= 65) ? 192 : 256;
  return 128 * (nb - 1) + n_last <= 448;
}
// ring kernel: d_k 64 / 128, chunk sizes dividing 128 or >= 128, any context
bool attention_tc_supported(int c, int l, int r, int dk) {
  if (!(dk == 64 || dk == 128) || l < 0 || r < 0) return false;
  return c == 8 || c == 16 || c == 32 || c == 64 || c >= 128;
}

bool run_attention(int impl, const AttnParams& p, cudaStream_t st, std::string* err) {
  if (p.n_chunks == 0) return true;
  const int dk = p.d / p.heads;
  if (impl == 1) {
    if (!attention_tc_supported(p.c, p.l, p.r, dk)) { *err = "attention: tcgen05 kernels need d_k 64 or 128 and a chunk size in {8,16,32,64} or >= 128"; return false; }
    if (attention_tc_fast_supported(p.c, p.l, p.r, dk)) return launch_attention_tc(p, st, err);
    return launch_attention_ring(p, st, err);
  }
  if (impl == 3) {      // tests / tools: force the ring kernel
    if (!attention_tc_supported(p.c, p.l, p.r, dk)) { *err = "attention: ring kernel does not cover this shape"; return false; }
    return launch_attention_ring(p, st, err);
  }
  const int W = p.l + p.c + p.r;
  const size_t smem = size_t(4) * (2 * dk + W) * sizeof(float);
  dim3 grid(p.n_chunks, p.heads);
  cudaError_t e;
  if (dk == 64) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(attention_simt_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    attention_simt_kernel<64><<<grid, 128, smem, st>>>(p);
    ++cf::g_kernel_launches;
  } else if (dk == 128) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(attention_simt_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    attention_simt_kernel<128><<<grid, 128, smem, st>>>(p);
    ++cf::g_kernel_launches;
  } else { *err = "attention: d_k must be 64 or 128"; return false; }
  e = cudaGetLastError();
  if (e != cudaSuccess) { *err = std::string("attention launch: ") + cudaGetErrorString(e); return false; }
  return true;
}

struct Carver {  // carve a caller-provided workspace into 256-byte aligned pieces
  uint8_t* base; size_t off = 0;
  explicit Carver(void* b) : base(static_cast<uint8_t*>(b)) {}
  template <typename T> T* take(size_t count) {
    off = (off + 255) & ~size_t(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * sizeof(T);
    return p;
  }
};

constexpr int FE_SLAB_CHUNKS = 256;

struct EncodeWs {
  ChunkSrc* chunk_src; int2 *att_range, *conv_range, *out_range; int* seq_limit;
  ChunkSrc* chunk_src_last; int2* out_range_last;      // per utterance: the entries of its LAST chunk (compact streaming)
  size_t table_bytes;                                  // the table region = [chunk_src, chunk_src + table_bytes)
  float* x; bf16 *y, *hbuf, *qkv, *ctx, *g, *z, *a1, *b1, *a2, *b2;
  size_t qkv_rows, g_rows;
  size_t total;
};

EncodeWs carve_encode(const cf_handle* h, const cf_plan* p, void* base) {
  const int d = h->cfg.d_model, F = h->cfg.ffn, c = p->c;
  const size_t n = size_t(p->n), Mr = n * c;
  const int T2 = 2 * c + 1, F2 = (((h->cfg.feat_dim - 3) / 2 + 1) - 3) / 2 + 1;
  const size_t S = std::min<size_t>(n, FE_SLAB_CHUNKS);
  EncodeWs w;
  Carver cv(base);
  w.chunk_src = cv.take<ChunkSrc>(n);
  w.att_range = cv.take<int2>(n + 16);
  w.conv_range = cv.take<int2>(n);
  w.out_range = cv.take<int2>(n);
  w.seq_limit = cv.take<int>(size_t(p->B));
  w.chunk_src_last = cv.take<ChunkSrc>(size_t(p->B));
  w.out_range_last = cv.take<int2>(size_t(p->B));
  w.table_bytes = (cv.off + 15) & ~size_t(15);
  w.x = cv.take<float>(Mr * d);
  w.y = cv.take<bf16>(Mr * d);
  w.hbuf = cv.take<bf16>(Mr * F);
  w.qkv_rows = size_t(p->l) + Mr + size_t(p->r) + 2 * size_t(c) + 128;
  w.qkv = cv.take<bf16>(w.qkv_rows * 4 * d);
  w.ctx = cv.take<bf16>(Mr * d);
  w.g_rows = Mr + 2 * size_t(p->lorder) + 64;
  w.g = cv.take<bf16>(w.g_rows * d);
  w.z = cv.take<bf16>(Mr * d);
  w.a1 = cv.take<bf16>(S * T2 * F2 * d);
  w.b1 = cv.take<bf16>(S * T2 * F2 * d);
  w.a2 = cv.take<bf16>(S * c * h->F3 * d);
  w.b2 = cv.take<bf16>(S * c * h->F3 * d);
  w.total = cv.off + 256;
  return w;
}

// Image of the per-call tables exactly as they lie at the start of the workspace (chunk sources, attention / conv / output
// ranges, per-sequence row limits): built on the host, then pulled into the workspace (cf_encode) or copied there once
// (cf_plan_pin).
static size_t table_image_bytes(const cf_plan* p, const EncodeWs& w) {
  (void)p;
  return w.table_bytes;                 // chunk_src is the first piece of the workspace
}
static void fill_table_image(const cf_plan* p, const EncodeWs& w, uint8_t* base, size_t need) {
  const uint8_t* dev0 = reinterpret_cast<const uint8_t*>(w.chunk_src);
  const size_t o_ar = reinterpret_cast<const uint8_t*>(w.att_range) - dev0, o_cr = reinterpret_cast<const uint8_t*>(w.conv_range) - dev0;
  const size_t o_or = reinterpret_cast<const uint8_t*>(w.out_range) - dev0, o_sl = reinterpret_cast<const uint8_t*>(w.seq_limit) - dev0;
  memset(base, 0, need);
  ChunkSrc* cs = reinterpret_cast<ChunkSrc*>(base);
  int2* ar = reinterpret_cast<int2*>(base + o_ar);
  int2* cr = reinterpret_cast<int2*>(base + o_cr);
  int2* orr = reinterpret_cast<int2*>(base + o_or);
  for (int g = 0; g < p->n; ++g) {
    cs[g].feat_row = p->chunk_feat_row[g]; cs[g].in_len = p->chunk_in_len[g]; cs[g].pad_ = 0;
    const cf_chunk_entry& e = p->chunks[g];
    ar[g] = make_int2(e.att_lo, e.att_hi); cr[g] = make_int2(e.conv_lo, e.conv_hi); orr[g] = make_int2(e.out_lo, e.out_hi);
  }
  // (the 16 phantom chunks behind the last attention tile stay empty: zeroed above)
  if (p->mode == 1) memcpy(base + o_sl, p->seq_valid_rows.data(), size_t(p->B) * sizeof(int));
  // per utterance, the entries of its last chunk: the chunk list of the row-wise kernels in compact streaming
  ChunkSrc* csl = reinterpret_cast<ChunkSrc*>(base + (reinterpret_cast<const uint8_t*>(w.chunk_src_last) - dev0));
  int2* orl = reinterpret_cast<int2*>(base + (reinterpret_cast<const uint8_t*>(w.out_range_last) - dev0));
  int g_end = 0;
  for (int u = 0; u < p->B; ++u) {
    g_end += p->n_chunks[u];
    if (p->n_chunks[u] > 0) { csl[u] = cs[g_end - 1]; orl[u] = orr[g_end - 1]; }
  }
}

// Projected relative-position tables P_l = linear_pos_l(PE) for all layers (embedding.py:119-174, attention.py:482):
// input independent, computed once per (c, l, r) and cached on the handle.
int get_pos_table(cf_handle* h, int c, int l, int r, cudaStream_t st, const PosTable** out) {
  for (auto& t : h->pos_tables)
    if (t.c == c && t.l == l && t.r == r) { t.last_use = ++h->use_clock; *out = &t; return CF_OK; }
  const int d = h->cfg.d_model, L = h->cfg.layers;
  // Bounded cache: full attention (chunk = T') makes one table per utterance length, so a service fed varied lengths would
  // otherwise grow device memory without limit.  Evict the least recently used table (work that reads it was enqueued on a
  // stream; wait for it before freeing).
  if (h->pos_tables.size() >= kMaxPosTables) {
    size_t victim = 0;
    for (size_t i = 1; i < h->pos_tables.size(); ++i)
      if (h->pos_tables[i].last_use < h->pos_tables[victim].last_use) victim = i;
    CF_CUDA(h, cudaStreamSynchronize(st));
    CF_CUDA(h, cudaDeviceSynchronize());
    cudaFree(h->pos_tables[victim].dev);
    h->pos_tables.erase(h->pos_tables.begin() + victim);
  }
  PosTable t;
  t.c = c; t.l = l; t.r = r; t.R = 2 * c + l + r - 1;
  t.Rpad = ((t.R + 127) / 128) * 128;
  std::vector<uint16_t> pe(size_t(t.Rpad) * d, 0);
  for (int p = 0; p < t.R; ++p) {
    const float rho = float(c + l - 1 - p);
    for (int m = 0; m < d / 2; ++m) {
      const float w = expf(float(2 * m) * -(logf(10000.0f) / float(d)));
      const float ang = fabsf(rho) * w;
      const float s = sinf(ang) * (rho < 0 ? -1.f : (rho > 0 ? 1.f : 0.f)), co = cosf(ang);
      __nv_bfloat16 bs = __float2bfloat16_rn(s), bc = __float2bfloat16_rn(co);
      memcpy(&pe[size_t(p) * d + 2 * m], &bs, 2);
      memcpy(&pe[size_t(p) * d + 2 * m + 1], &bc, 2);
    }
  }
  bf16* pe_dev = nullptr;
  auto bail = [&](int code, const std::string& msg) {
    if (pe_dev) cudaFree(pe_dev);
    if (t.dev) cudaFree(t.dev);
    return fail(h, code, msg);
  };
  cudaError_t e = cudaMalloc(&pe_dev, pe.size() * 2);
  if (e == cudaSuccess) e = cudaMalloc(&t.dev, size_t(L) * t.Rpad * d * 2);
  if (e == cudaSuccess) e = cudaMemcpyAsync(pe_dev, pe.data(), pe.size() * 2, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(t.dev, 0, size_t(L) * t.Rpad * d * 2, st);
  if (e != cudaSuccess) return bail(CF_ERR_CUDA, std::string("position table: ") + cudaGetErrorString(e));
  for (int i = 0; i < L; ++i) {
    GemmLaunch g{};
    g.A = pe_dev; g.lda = d; g.B = h->layers[i].pos_w; g.ldb = d; g.M = t.R; g.N = d; g.K = d; g.epi = EPI_BF16;
    g.act = ACT_NONE; g.ep.bias = h->zeros; g.out = t.dev + size_t(i) * t.Rpad * d; g.ldo = d;
    std::string err;
    if (!launch_gemm(g, h->num_sms, st, &err)) return bail(CF_ERR_CUDA, err);
  }
  e = cudaStreamSynchronize(st);               // `pe` (pageable host memory) and pe_dev go out of scope
  if (e != cudaSuccess) return bail(CF_ERR_CUDA, std::string("position table: ") + cudaGetErrorString(e));
  cudaFree(pe_dev);
  t.last_use = ++h->use_clock;
  h->pos_tables.push_back(t);
  *out = &h->pos_tables.back();
  return CF_OK;
}

}  // namespace

extern "C" size_t cf_workspace_bytes(const cf_handle* h, const cf_plan* p) {
  if (!h || !p) return 0;
  return carve_encode(h, p, nullptr).total;
}

// Copy the plan's tables into the table region of `workspace` now (synchronously) and remember it: cf_encode calls with this
// plan AND this workspace then skip the per-call table staging (a pinned ring guarded by events), issue nothing that depends
// on host memory, and can be captured in a CUDA graph and replayed.  The caller must not run another plan through that
// workspace in between; cf_plan_pin(h, p, NULL, 0, stream) undoes it.
extern "C" int cf_plan_pin(cf_handle* h, cf_plan* p, void* workspace, size_t workspace_bytes, void* stream) {
  if (!h || !p) return fail(h, CF_ERR_INVALID, "cf_plan_pin: null argument");
  if (!workspace) { p->resident_ws = nullptr; return CF_OK; }
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return fail(h, CF_ERR_INVALID, "cf_plan_pin: workspace must be 256-byte aligned");
  if (workspace_bytes < cf_workspace_bytes(h, p)) return fail(h, CF_ERR_WORKSPACE, "cf_plan_pin: workspace too small");
  if (p->n == 0) return CF_OK;
  DeviceGuard guard(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const EncodeWs w = carve_encode(h, p, workspace);
  const size_t need = table_image_bytes(p, w);
  std::vector<uint8_t> img(need);
  fill_table_image(p, w, img.data(), need);
  CF_CUDA(h, cudaMemcpyAsync(w.chunk_src, img.data(), need, cudaMemcpyHostToDevice, st));
  CF_CUDA(h, cudaStreamSynchronize(st));
  p->resident_ws = workspace;
  return CF_OK;
}

// --------------------------------------------------------------------------------------------------------------------
// encoder driver
// --------------------------------------------------------------------------------------------------------------------
// The next cf_encode call carries `n_streams` concurrent streams (frame-synchronous streaming): the plan holds one utterance
// of `placeholder_chunks` + 1 chunks per stream and the caches are (L, B, H, l, 2 d_k) / (L, B, d, lorder), updated in place.
extern "C" int cf_encode_streams(cf_handle* h, int n_streams, int placeholder_chunks, int advance) {
  if (!h || n_streams < 0 || placeholder_chunks < 0 || advance < 0) return fail(h, CF_ERR_INVALID, "cf_encode_streams: bad argument");
  h->streams_n = n_streams; h->streams_ph = placeholder_chunks; h->streams_adv = advance;
  return CF_OK;
}

extern "C" int64_t cf_encode_output_rows(const cf_handle* h) { return h ? int64_t(h->last_out_rows) : 0; }

extern "C" int cf_encode_feature_events(cf_handle* h, int n, const int64_t* rows_ready, void* const* events) {
  if (!h || n < 0 || (n > 0 && (!rows_ready || !events))) return fail(h, CF_ERR_INVALID, "cf_encode_feature_events: bad argument");
  h->ev_rows.assign(rows_ready, rows_ready + n);
  h->ev.resize(n);
  for (int i = 0; i < n; ++i) h->ev[i] = static_cast<cudaEvent_t>(events[i]);
  return CF_OK;
}

extern "C" int cf_encode(cf_handle* h, const cf_plan* p, const float* feats, void* att_cache, void* cnn_cache,
                         int trunc, void* out, int out_dtype, void* out_bf16, void* workspace, size_t workspace_bytes,
                         void* stream) {
  if (!h || !p || !feats || !out || !workspace) return fail(h, CF_ERR_INVALID, "cf_encode: null argument");
  if (!h->finalized) return fail(h, CF_ERR_STATE, "cf_encode: call cf_finalize_weights first");
  if (out_dtype != CF_F32 && out_dtype != CF_BF16) return fail(h, CF_ERR_INVALID, "cf_encode: bad out_dtype");
  if (p->kernel != h->cfg.kernel) return fail(h, CF_ERR_INVALID, "cf_encode: plan conv kernel differs from the model's");
  // per-call state armed by cf_encode_streams / cf_encode_feature_events is consumed here, so that every exit path clears it
  const int ns = h->streams_n, ph = h->streams_ph;
  // frames a stream advances per step: the plan's chunk, or less when the chunk carries right-context frames behind it
  const int adv = (h->streams_adv > 0 && ns > 0) ? h->streams_adv : p->c;
  h->streams_n = 0; h->streams_ph = 0; h->streams_adv = 0;
  std::vector<cudaEvent_t> ev;
  std::vector<int64_t> ev_rows;
  ev.swap(h->ev);
  ev_rows.swap(h->ev_rows);
  if (ns > 0) {
    if (p->mode != 0 || p->B != ns || !att_cache || !cnn_cache)
      return fail(h, CF_ERR_INVALID, "cf_encode: multi-stream call needs a masked-batch plan with one utterance per stream and both caches");
    for (int u = 0; u < p->B; ++u)
      if (p->n_chunks[u] != ph + 1) return fail(h, CF_ERR_INVALID, "cf_encode: every stream must span placeholder_chunks + 1 chunks");
    if (ph * p->c < std::max(p->l, p->lorder) || p->r != 0)
      return fail(h, CF_ERR_INVALID, "cf_encode: placeholder rows must cover the left context and the conv cache; right context must be 0");
    if (adv > p->c) return fail(h, CF_ERR_INVALID, "cf_encode: a stream cannot advance by more than the plan's chunk size");
  } else if ((att_cache || cnn_cache) && (p->mode != 0 || p->B != 1))
    return fail(h, CF_ERR_INVALID, "cf_encode: streaming caches need a masked-batch plan with one utterance");
  if ((att_cache != nullptr) != (cnn_cache != nullptr))
    return fail(h, CF_ERR_INVALID, "cf_encode: pass both caches or neither");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DeviceGuard guard(h->device);
  const EncodeWs w = carve_encode(h, p, workspace);
  if (w.total > workspace_bytes) return fail(h, CF_ERR_WORKSPACE, "cf_encode: workspace too small");
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return fail(h, CF_ERR_INVALID, "cf_encode: workspace must be 256-byte aligned");
  const int d = h->cfg.d_model, F = h->cfg.ffn, L = h->cfg.layers, H = h->cfg.heads, dk = d / H;
  const int c = p->c, l = p->l, r = p->r, lo = p->lorder;
  const int n = p->n;
  const long long Mr = (long long)n * c;
  h->last_out_rows = Mr;
  if (Mr == 0) return CF_OK;
  // Compact streaming: a stream is `ph` placeholder chunks (whose only purpose is to hold the left-context K / V and conv rows
  // in front of the real chunk in the flat QKV / GLU buffers) + its real chunk.  Every row-wise kernel (front-end, GEMMs,
  // LayerNorms) then works on the ns x c real rows only: the QKV / GLU GEMMs scatter their output into the per-stream layout and
  // the GEMMs behind attention / the conv core gather their A operand from it (3-D tensor maps); attention and the conv core still
  // walk every chunk.  `out` receives ns x c rows, stream after stream.
  const bool compact = ns > 0 && adv == c && h->opt_stream_compact != 0 && h->opt_fused_layernorm != 0 && h->opt_ln_split != 0 &&
                       h->opt_fused_ffn == 0 && (128 % c) == 0 && ev.empty();
  const long long Mrow = compact ? (long long)ns * c : Mr;      // rows of the row-wise kernels
  const int n_row = compact ? ns : n;                           // their chunks
  const int grp_rows = (ph + 1) * c;                            // compact: rows of a stream's block in the QKV / GLU / ctx / z buffers
  h->last_out_rows = Mrow;
  if (att_cache && trunc < 0) return fail(h, CF_ERR_INVALID, "cf_encode: truncated_context_size must be >= 0");
  if (!ev_rows.empty()) {
    long long need_all = 0;
    for (int g = 0; g < n; ++g) need_all = std::max<long long>(need_all, p->chunk_feat_row[g] + std::max(p->chunk_in_len[g], 0));
    if (ev_rows.back() < need_all) return fail(h, CF_ERR_INVALID, "cf_encode: the feature events cover fewer rows than the plan reads");
  }
  // kv[: trunc + l][-l:] (attention.py:466-467) and x[:, : trunc + lorder][:, -lorder:] (convolution.py:228-230) clamp at
  // the end of the buffer: a final segment shorter than the kept context hands over its last rows.
  if (trunc > Mr) trunc = int(Mr);
  std::string err;
  const PosTable* pos = nullptr;
  int rc = get_pos_table(h, c, l, r, st, &pos);
  if (rc != CF_OK) return rc;

  // ---- tables: built in a pinned staging block owned by the handle, laid out exactly like the workspace's table region, and
  // pulled into the workspace by one small kernel that reads the mapped host memory (no host synchronisation, and no
  // copy-engine operation that would queue behind the feature upload; a ring of four blocks, each guarded by an event)
  auto launch_zero = [&](void* dst, size_t bytes) {      // bytes and dst are multiples of 16 here
    if (bytes == 0) return;
    const long long n16 = (long long)(bytes / 16);
    zero16_kernel<<<unsigned(std::min<long long>((n16 + 255) / 256, 4LL * h->num_sms)), 256, 0, st>>>(static_cast<uint4*>(dst), n16);
    ++cf::g_kernel_launches;
  };
  if (p->resident_ws != workspace) {
    uint8_t* dev0 = reinterpret_cast<uint8_t*>(w.chunk_src);
    const size_t need = table_image_bytes(p, w);
    PinnedStage& sg = h->stage[h->stage_next];
    h->stage_next = (h->stage_next + 1) % 4;
    if (sg.in_flight) { CF_CUDA(h, cudaEventSynchronize(sg.done)); sg.in_flight = false; }
    if (sg.bytes < need) {
      if (sg.host) cudaFreeHost(sg.host);
      sg.host = nullptr; sg.bytes = 0;
      CF_CUDA(h, cudaHostAlloc(&sg.host, need + need / 2, cudaHostAllocMapped | cudaHostAllocPortable));
      sg.bytes = need + need / 2;
    }
    if (!sg.done) CF_CUDA(h, cudaEventCreateWithFlags(&sg.done, cudaEventDisableTiming));
    fill_table_image(p, w, static_cast<uint8_t*>(sg.host), need);
    void* mapped = nullptr;
    CF_CUDA(h, cudaHostGetDevicePointer(&mapped, sg.host, 0));
    const long long n16 = (long long)(need / 16);
    copy16_kernel<<<unsigned(std::min<long long>((n16 + 255) / 256, 2LL * h->num_sms)), 256, 0, st>>>(static_cast<const uint4*>(mapped),
                                                                                                   reinterpret_cast<uint4*>(dev0), n16);
    ++cf::g_kernel_launches;
    CF_CUDA(h, cudaGetLastError());
    CF_CUDA(h, cudaEventRecord(sg.done, st));
    sg.in_flight = true;
  }
  // zero the halo rows of the flat buffers once (cache rows are rewritten per layer when streaming); compact streaming: the
  // placeholder rows are never written by a GEMM, so the whole buffers are cleared (their Q columns and the rows outside the
  // cached context must be finite)
  if (compact) {
    launch_zero(w.qkv, w.qkv_rows * 4 * d * sizeof(bf16));
    launch_zero(w.g, w.g_rows * d * sizeof(bf16));
  } else {
    launch_zero(w.qkv, size_t(l) * 4 * d * sizeof(bf16));
    launch_zero(w.qkv + (size_t(l) + Mr) * 4 * d, (w.qkv_rows - size_t(l) - Mr) * 4 * d * sizeof(bf16));
    launch_zero(w.g, size_t(lo) * d * sizeof(bf16));
    launch_zero(w.g + (size_t(lo) + Mr) * d, (w.g_rows - size_t(lo) - Mr) * d * sizeof(bf16));
  }
  CF_CUDA(h, cudaGetLastError());

  struct EpiArgs { const float* bias = nullptr; void* out = nullptr; long long ldo = 0; int act = ACT_NONE;
                   const float* resid = nullptr; long long ld_resid = 0; float alpha = 1.0f;
                   const int2* row_range = nullptr; int rows_per_chunk = 1; int family = 0;
                   bool scatter = false;      // compact streaming: bf16 / GLU output rows go to the real chunk of every stream
                   bool gather = false; };    // compact streaming: the A operand is the real chunk of every stream
  auto gemm = [&](const void* A, long long lda, const void* B, long long ldb, long long M, int N, int K, int epi,
                  const EpiArgs& e) -> bool {
    GemmLaunch g{};
    g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.M = int(M); g.N = N; g.K = K; g.epi = epi; g.act = e.act;
    g.out = e.out; g.ldo = e.ldo;
    g.ep.bias = e.bias; g.ep.resid = e.resid; g.ep.ld_resid = e.ld_resid; g.ep.alpha = e.alpha;
    g.ep.row_range = e.row_range; g.ep.rows_per_chunk = e.rows_per_chunk;
    g.timing = &h->timing; g.family = e.family; g.variant = h->opt_gemm_pair;
    if (e.scatter) { g.scatter_rows = c; g.scatter_row0 = ph * c; g.scatter_group_rows = grp_rows; g.scatter_groups = ns; }
    return launch_gemm(g, h->num_sms, st, &err);
  };
  // residual GEMM + the LayerNorm(s) that follow it, one kernel (gemm_ln.cuh)
  struct LnArgs { int mode = LNM_Y; const float* w1 = nullptr; const float* b1 = nullptr; const float* w2 = nullptr;
                  const float* b2 = nullptr; float* x_out = nullptr; void* y_out = nullptr; bool limit = false; };
  auto gemm_ln = [&](const void* A, long long lda, const void* B, long long ldb, long long M, int K, const EpiArgs& e,
                     const LnArgs& q) -> bool {
    GemmLnLaunch g{};
    g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.M = int(M); g.N = d; g.K = K; g.bias = e.bias; g.resid = e.resid;
    g.ld_resid = e.ld_resid; g.alpha = e.alpha; g.row_range = e.row_range; g.rows_per_chunk = e.rows_per_chunk; g.mode = q.mode;
    g.ln1_w = q.w1; g.ln1_b = q.b1; g.ln2_w = q.w2; g.ln2_b = q.b2; g.x_out = q.x_out; g.ldx = d; g.y_out = q.y_out; g.ldy = d;
    g.row_limit = q.limit ? w.seq_limit : nullptr; g.rows_per_seq = q.limit ? p->rows_per_seq : 1;
    g.timing = &h->timing; g.family = e.family; g.variant = h->opt_ln_split;
    if (e.gather) { g.gather_rows = c; g.gather_row0 = ph * c; g.gather_group_rows = grp_rows; g.gather_groups = ns; }
    return launch_gemm_ln(g, h->num_sms, st, &err);
  };
  const bool fuse_ln = h->opt_fused_layernorm != 0;
  const bool fuse_ffn = fuse_ln && h->opt_fused_ffn != 0 && ffn_fused_supported(d, F);
  // whole feed-forward module: x <- x + 0.5 * (W2 SiLU(W1 y + b1) + b2) followed by the LayerNorm(s) of `q`
  auto ffn = [&](const bf16* w1, const float* b1, const bf16* w2, const float* b2, const LnArgs& q) -> bool {
    FfnLaunch g{};
    g.Y = w.y; g.ldy_in = d; g.W1 = w1; g.b1 = b1; g.W2 = w2; g.b2 = b2; g.M = int(Mrow); g.d = d; g.F = F; g.resid = w.x;
    g.ld_resid = d; g.alpha = 0.5f; g.mode = q.mode; g.ln1_w = q.w1; g.ln1_b = q.b1; g.ln2_w = q.w2; g.ln2_b = q.b2;
    g.x_out = q.x_out; g.ldx = d; g.y_out = q.y_out; g.ldy = d;
    g.timing = &h->timing; g.family = CF_FAMILY_FFN_FUSED;
    return launch_ffn_fused(g, h->num_sms, st, &err);
  };
  // feed-forward module as two GEMMs (w_1 + SiLU -> hidden activation in global memory -> w_2 + residual + LayerNorm(s)),
  // optionally slab by slab so that a slab's hidden activation never has to leave L2
  auto ffn_two = [&](const bf16* w1, const float* b1, const bf16* w2, const float* b2, const LnArgs& q0) -> bool {
    const long long slab = h->opt_ffn_slab_rows > 0 ? h->opt_ffn_slab_rows : Mrow;
    for (long long r0 = 0; r0 < Mrow; r0 += slab) {
      const long long rows = std::min(slab, Mrow - r0);
      EpiArgs e1; e1.bias = b1; e1.out = w.hbuf; e1.ldo = F; e1.act = ACT_SILU; e1.family = CF_FAMILY_FFN_W1;
      if (!gemm(w.y + r0 * d, d, w1, d, rows, F, d, EPI_BF16, e1)) return false;
      EpiArgs e2; e2.bias = b2; e2.resid = w.x + r0 * d; e2.ld_resid = d; e2.alpha = 0.5f; e2.family = CF_FAMILY_FFN_W2;
      LnArgs q = q0;
      if (q.x_out) q.x_out += r0 * d;
      if (q.y_out) q.y_out = static_cast<bf16*>(q.y_out) + r0 * d;
      if (!gemm_ln(w.hbuf, F, w2, F, rows, F, e2, q)) return false;
    }
    return true;
  };
#define CF_TRY(expr) do { if (!(expr)) return fail(h, CF_ERR_CUDA, "cf_encode: " + err); } while (0)

  // ---- front-end: slabs of chunks through conv0+dw1 -> pw1 -> dw2 -> pw2 -> out Linear
  {
    const int T2 = 2 * c + 1, F1 = (h->cfg.feat_dim - 3) / 2 + 1, F2 = (F1 - 3) / 2 + 1, F3 = h->F3;
    if (F2 < 16) return fail(h, CF_ERR_INVALID, "cf_encode: feat_dim too small for the front-end tiling");
    size_t ev_next = 0;
    const ChunkSrc* src_tab = compact ? w.chunk_src_last : w.chunk_src;
    for (int g0 = 0; g0 < n_row; g0 += FE_SLAB_CHUNKS) {
      const int S = std::min(FE_SLAB_CHUNKS, n_row - g0);
      if (!ev.empty()) {
        // features may still be arriving on another stream: wait only for the rows this slab reads
        long long need = 0;
        for (int g = g0; g < g0 + S; ++g) need = std::max<long long>(need, p->chunk_feat_row[g] + std::max(p->chunk_in_len[g], 0));
        while (ev_next < ev.size() && (ev_next == 0 || ev_rows[ev_next - 1] < need)) {
          CF_CUDA(h, cudaStreamWaitEvent(st, ev[ev_next], 0));
          ++ev_next;
        }
      }
      Fe1Params f1{};
      f1.feats = feats; f1.chunks = src_tab + g0; f1.wpack = h->fe_wpack; f1.cmvn_mean = h->cmvn_mean; f1.cmvn_istd = h->cmvn_istd;
      f1.out = w.a1; f1.n_chunks = S; f1.feat_dim = h->cfg.feat_dim; f1.T2 = T2; f1.F2 = F2; f1.in_rows = p->in_rows;
      if (!run_frontend_conv(h->cfg.feat_dim == 80 ? 2 : 0, d, f1, h->num_sms, st, &err)) return fail(h, CF_ERR_CUDA, "cf_encode: " + err);
      EpiArgs e1; e1.bias = h->fe_b3; e1.out = w.b1; e1.ldo = d; e1.act = ACT_RELU;
      CF_TRY(gemm(w.a1, d, h->fe_w3, d, (long long)S * T2 * F2, d, d, EPI_BF16, e1));
      Fe2Params f2{};
      f2.in = w.b1; f2.out = w.a2; f2.w = h->fe_dw2_w; f2.bias = h->fe_dw2_b; f2.T2 = T2; f2.F2 = F2; f2.T3 = c; f2.F3 = F3;
      f2.total = (long long)S * c * F3 * (d / 8);
      const long long rows2 = (long long)S * c;                       // output time rows of the slab
      const unsigned g2 = unsigned((rows2 * (d / 8) + 191) / 192);
      if (d == 512) frontend_dw2_rows_kernel<512><<<g2, 192, 0, st>>>(f2, rows2);
      else frontend_dw2_rows_kernel<256><<<g2, 192, 0, st>>>(f2, rows2);
      ++cf::g_kernel_launches;
      CF_CUDA(h, cudaGetLastError());
      EpiArgs e2; e2.bias = h->fe_b6; e2.out = w.b2; e2.ldo = d; e2.act = ACT_RELU;
      CF_TRY(gemm(w.a2, d, h->fe_w6, d, (long long)S * c * F3, d, d, EPI_BF16, e2));
      // (xW + b) * sqrt(d)  (subsampling.py:164, embedding.py:198)
      EpiArgs e3; e3.bias = h->fe_bout; e3.out = w.x + (size_t)g0 * c * d; e3.ldo = d; e3.alpha = sqrtf(float(d));
      if (fuse_ln) {   // + norm_ff_macaron of layer 0
        LnArgs q; q.mode = LNM_Y; q.w1 = h->layers[0].ln_ffm_w; q.b1 = h->layers[0].ln_ffm_b;
        q.x_out = w.x + (size_t)g0 * c * d; q.y_out = w.y + (size_t)g0 * c * d;
        CF_TRY(gemm_ln(w.b2, (long long)F3 * d, h->fe_wout, (long long)F3 * d, (long long)S * c, F3 * d, e3, q));
      } else
      CF_TRY(gemm(w.b2, (long long)F3 * d, h->fe_wout, (long long)F3 * d, (long long)S * c, d, F3 * d, EPI_F32, e3));
    }
  }

  for (size_t i = 0; i < ev.size(); ++i) CF_CUDA(h, cudaStreamWaitEvent(st, ev[i], 0));   // (no-op for events already waited on)

  // ---- layers (encoder_layer.py:155-248)
  auto ln = [&](int mode, const float* w1, const float* b1, const float* w2, const float* b2, float* xo, bf16* y,
                bool limit) -> bool {
    LnParams q{};
    q.x_in = w.x; q.x_out = xo; q.y = y; q.w1 = w1; q.b1 = b1; q.w2 = w2; q.b2 = b2; q.rows = Mrow;
    q.row_limit = limit ? w.seq_limit : nullptr; q.rows_per_seq = limit ? p->rows_per_seq : 1;
    return run_layernorm(mode, d, q, st, &err);
  };
  const bool use_tc = kAttentionTcReady && attention_tc_supported(c, l, r, dk);
  if (!fuse_ln) CF_TRY(ln(0, h->layers[0].ln_ffm_w, h->layers[0].ln_ffm_b, nullptr, nullptr, nullptr, w.y, false));
  for (int i = 0; i < L; ++i) {
    const LayerW& lw = h->layers[i];
    // macaron FFN: x += 0.5 * W2 SiLU(W1 LN(x) + b1) + b2
    if (fuse_ln) {
      LnArgs q; q.mode = LNM_Y; q.w1 = lw.ln_mha_w; q.b1 = lw.ln_mha_b; q.x_out = w.x; q.y_out = w.y;
      if (fuse_ffn) CF_TRY(ffn(lw.ffm_w1, lw.ffm_b1, lw.ffm_w2, lw.ffm_b2, q));
      else CF_TRY(ffn_two(lw.ffm_w1, lw.ffm_b1, lw.ffm_w2, lw.ffm_b2, q));
    } else {
    { EpiArgs e; e.bias = lw.ffm_b1; e.out = w.hbuf; e.ldo = F; e.act = ACT_SILU; e.family = CF_FAMILY_FFN_W1;
      CF_TRY(gemm(w.y, d, lw.ffm_w1, d, Mrow, F, d, EPI_BF16, e)); }
    { EpiArgs e; e.bias = lw.ffm_b2; e.out = w.x; e.ldo = d; e.resid = w.x; e.ld_resid = d; e.alpha = 0.5f; e.family = CF_FAMILY_FFN_W2;
      if (fuse_ln) {
        LnArgs q; q.mode = LNM_Y; q.w1 = lw.ln_mha_w; q.b1 = lw.ln_mha_b; q.x_out = w.x; q.y_out = w.y;
        CF_TRY(gemm_ln(w.hbuf, F, lw.ffm_w2, F, Mrow, F, e, q));
      } else CF_TRY(gemm(w.hbuf, F, lw.ffm_w2, F, Mrow, d, F, EPI_F32, e)); }
    }
    // self-attention
    if (!fuse_ln) CF_TRY(ln(0, lw.ln_mha_w, lw.ln_mha_b, nullptr, nullptr, nullptr, w.y, false));
    if (att_cache && l > 0 && ns == 0) {
      const int tot = l * H * 2 * dk;
      att_cache_import_kernel<<<(tot + 255) / 256, 256, 0, st>>>(static_cast<const float*>(att_cache) + size_t(i) * tot, w.qkv, l, H, dk, d);
      ++cf::g_kernel_launches;
    }
    { EpiArgs e; e.bias = lw.qkv_b; e.out = w.qkv + size_t(l) * 4 * d; e.ldo = 4 * d; e.scatter = compact;
      CF_TRY(gemm(w.y, d, lw.qkv_w, d, Mrow, 4 * d, d, EPI_BF16, e)); }
    if (att_cache && l > 0 && ns == 0) {
      const int tot = l * H * 2 * dk;
      att_cache_export_kernel<<<(tot + 255) / 256, 256, 0, st>>>(static_cast<float*>(att_cache) + size_t(i) * tot, w.qkv, l, H, dk, d, trunc);
      ++cf::g_kernel_launches;
    }
    if (ns > 0 && l > 0) {   // placeholder rows <- caches, then caches <- last l rows of cache + frames, for every stream
      const long long tot = (long long)ns * l * H * 2 * dk;
      float* cl = static_cast<float*>(att_cache) + size_t(i) * tot;
      const unsigned blocks8 = unsigned((tot / 8 + 255) / 256);      // thread = 8 elements (d_k is a multiple of 8)
      att_cache_streams_kernel<<<blocks8, 256, 0, st>>>(cl, w.qkv, ns, l, H, dk, d, (ph + 1) * c, ph * c, c, l, 0);
      att_cache_streams_kernel<<<blocks8, 256, 0, st>>>(cl, w.qkv, ns, l, H, dk, d, (ph + 1) * c, ph * c, adv, l, 1);
      cf::g_kernel_launches += 2;
    }
    { AttnParams a{};
      a.qkv = w.qkv; a.pos = pos->dev + size_t(i) * pos->Rpad * d; a.range = w.att_range; a.ctx = w.ctx;
      a.n_chunks = n; a.c = c; a.l = l; a.r = r; a.d = d; a.heads = H; a.scale = 1.0f / sqrtf(float(dk)); a.prescaled = 1;
      CF_TRY(run_attention(use_tc ? 1 : 0, a, st, &err)); }
    { EpiArgs e; e.bias = lw.o_b; e.out = w.x; e.ldo = d; e.resid = w.x; e.ld_resid = d; e.alpha = 1.0f; e.gather = compact;
      if (fuse_ln) {
        LnArgs q; q.mode = LNM_Y; q.w1 = lw.ln_conv_w; q.b1 = lw.ln_conv_b; q.x_out = w.x; q.y_out = w.y; q.limit = p->mode == 1;
        CF_TRY(gemm_ln(w.ctx, d, lw.o_w, d, Mrow, d, e, q));
      } else CF_TRY(gemm(w.ctx, d, lw.o_w, d, Mrow, d, d, EPI_F32, e)); }
    // convolution module
    if (!fuse_ln) CF_TRY(ln(0, lw.ln_conv_w, lw.ln_conv_b, nullptr, nullptr, nullptr, w.y, p->mode == 1));
    if (cnn_cache && ns == 0) { cnn_cache_import_kernel<<<(d * lo + 255) / 256, 256, 0, st>>>(static_cast<const float*>(cnn_cache) + size_t(i) * d * lo, w.g, d, lo); ++cf::g_kernel_launches; }
    { EpiArgs e; e.bias = lw.pw1_b; e.out = w.g + size_t(lo) * d; e.ldo = d; e.scatter = compact;
      CF_TRY(gemm(w.y, d, lw.pw1_w, d, Mrow, 2 * d, d, EPI_GLU, e)); }
    if (cnn_cache && ns == 0) { cnn_cache_export_kernel<<<(d * lo + 255) / 256, 256, 0, st>>>(static_cast<float*>(cnn_cache) + size_t(i) * d * lo, w.g, d, lo, trunc); ++cf::g_kernel_launches; }
    if (ns > 0) {
      const long long tot = (long long)ns * d * lo;
      float* cl = static_cast<float*>(cnn_cache) + size_t(i) * tot;
      cnn_cache_streams_kernel<<<unsigned((tot + 255) / 256), 256, 0, st>>>(cl, w.g, ns, d, lo, (ph + 1) * c, ph * c, c, lo, 0);
      cnn_cache_streams_kernel<<<unsigned((tot + 255) / 256), 256, 0, st>>>(cl, w.g, ns, d, lo, (ph + 1) * c, ph * c, adv, lo, 1);
      cf::g_kernel_launches += 2;
    }
    { DwConvParams q{};
      q.g = w.g; q.z = w.z; q.w = lw.dw_w; q.bias = lw.dw_b; q.ln_w = lw.cn_w; q.ln_b = lw.cn_b; q.range = w.conv_range; q.c = c; q.n_chunks = n;
      q.no_norm = h->cfg.conv_norm == 1;
      q.sub_chunk = adv < c ? adv : 0;     // streaming with right context: the conv is cut at the true chunk grid
      const bool real_only = compact && c % 32 != 0;     // (the TMA kernel for chunk sizes 32 / 64 walks every chunk)
      if (real_only) { q.n_chunks = ns; q.chunk_stride = ph + 1; q.chunk_first = ph; }
      CF_TRY(run_dwconv(d, h->cfg.kernel, q, st, &err, (q.sub_chunk > 0 || real_only) ? 0 : (long long)w.g_rows, h->num_sms)); }
    { EpiArgs e; e.bias = lw.pw2_b; e.out = w.x; e.ldo = d; e.resid = w.x; e.ld_resid = d; e.alpha = 1.0f;
      e.row_range = compact ? w.out_range_last : w.out_range; e.rows_per_chunk = c; e.gather = compact;
      if (fuse_ln) {
        LnArgs q; q.mode = LNM_Y; q.w1 = lw.ln_ff_w; q.b1 = lw.ln_ff_b; q.x_out = w.x; q.y_out = w.y;
        CF_TRY(gemm_ln(w.z, d, lw.pw2_w, d, Mrow, d, e, q));
      } else CF_TRY(gemm(w.z, d, lw.pw2_w, d, Mrow, d, d, EPI_F32, e)); }
    // FFN
    if (!fuse_ln) CF_TRY(ln(0, lw.ln_ff_w, lw.ln_ff_b, nullptr, nullptr, nullptr, w.y, false));
    if (fuse_ln) {
      LnArgs q;
      q.w1 = lw.ln_fin_w; q.b1 = lw.ln_fin_b;
      if (i + 1 < L) {
        q.mode = LNM_XY; q.w2 = h->layers[i + 1].ln_ffm_w; q.b2 = h->layers[i + 1].ln_ffm_b; q.x_out = w.x; q.y_out = w.y;
      } else {   // norm_final of the last layer + after_norm (encoder.py:670-671)
        q.mode = LNM_FINAL; q.w2 = h->after_w; q.b2 = h->after_b;
        q.x_out = out_dtype == CF_F32 ? static_cast<float*>(out) : nullptr;
        q.y_out = out_dtype == CF_BF16 ? out : out_bf16;
      }
      if (fuse_ffn) CF_TRY(ffn(lw.ff_w1, lw.ff_b1, lw.ff_w2, lw.ff_b2, q));
      else CF_TRY(ffn_two(lw.ff_w1, lw.ff_b1, lw.ff_w2, lw.ff_b2, q));
      if (i + 1 == L && out_dtype == CF_BF16 && out_bf16 && out_bf16 != out)
        CF_CUDA(h, cudaMemcpyAsync(out_bf16, out, size_t(Mrow) * d * sizeof(bf16), cudaMemcpyDeviceToDevice, st));
      continue;
    }
    { EpiArgs e; e.bias = lw.ff_b1; e.out = w.hbuf; e.ldo = F; e.act = ACT_SILU; e.family = CF_FAMILY_FFN_W1;
      CF_TRY(gemm(w.y, d, lw.ff_w1, d, Mrow, F, d, EPI_BF16, e)); }
    { EpiArgs e; e.bias = lw.ff_b2; e.out = w.x; e.ldo = d; e.resid = w.x; e.ld_resid = d; e.alpha = 0.5f; e.family = CF_FAMILY_FFN_W2;
      if (fuse_ln) {
        LnArgs q;
        q.w1 = lw.ln_fin_w; q.b1 = lw.ln_fin_b;
        if (i + 1 < L) {
          q.mode = LNM_XY; q.w2 = h->layers[i + 1].ln_ffm_w; q.b2 = h->layers[i + 1].ln_ffm_b; q.x_out = w.x; q.y_out = w.y;
        } else {   // norm_final of the last layer + after_norm (encoder.py:670-671)
          q.mode = LNM_FINAL; q.w2 = h->after_w; q.b2 = h->after_b;
          q.x_out = out_dtype == CF_F32 ? static_cast<float*>(out) : nullptr;
          q.y_out = out_dtype == CF_BF16 ? out : out_bf16;
        }
        CF_TRY(gemm_ln(w.hbuf, F, lw.ff_w2, F, Mrow, F, e, q));
        if (i + 1 == L && out_dtype == CF_BF16 && out_bf16 && out_bf16 != out)
          CF_CUDA(h, cudaMemcpyAsync(out_bf16, out, size_t(Mrow) * d * sizeof(bf16), cudaMemcpyDeviceToDevice, st));
        continue;
      }
      CF_TRY(gemm(w.hbuf, F, lw.ff_w2, F, Mrow, d, F, EPI_F32, e)); }
    if (i + 1 < L) {
      CF_TRY(ln(1, lw.ln_fin_w, lw.ln_fin_b, h->layers[i + 1].ln_ffm_w, h->layers[i + 1].ln_ffm_b, w.x, w.y, false));
    } else {
      // norm_final of the last layer + after_norm (encoder.py:670-671)
      LnParams q{};
      q.x_in = w.x; q.w1 = lw.ln_fin_w; q.b1 = lw.ln_fin_b; q.w2 = h->after_w; q.b2 = h->after_b; q.rows = Mrow;
      q.x_out = out_dtype == CF_F32 ? static_cast<float*>(out) : nullptr;
      q.y = out_dtype == CF_BF16 ? static_cast<bf16*>(out) : static_cast<bf16*>(out_bf16);
      q.rows_per_seq = 1;
      CF_TRY(run_layernorm(2, d, q, st, &err));
      if (out_dtype == CF_BF16 && out_bf16 && out_bf16 != out)
        CF_CUDA(h, cudaMemcpyAsync(out_bf16, out, size_t(Mrow) * d * sizeof(bf16), cudaMemcpyDeviceToDevice, st));
    }
  }
  CF_CUDA(h, cudaGetLastError());
  return CF_OK;
}

// --------------------------------------------------------------------------------------------------------------------
// fbank (the step before the path)
// --------------------------------------------------------------------------------------------------------------------
extern "C" int64_t cf_fbank_num_frames(int64_t n_samples, int sample_rate, int frame_length_ms, int frame_shift_ms) {
  const int64_t flen = int64_t(sample_rate) * frame_length_ms / 1000, fshift = int64_t(sample_rate) * frame_shift_ms / 1000;
  if (flen <= 0 || fshift <= 0 || n_samples < flen) return 0;
  return 1 + (n_samples - flen) / fshift;                 // snip_edges=True (kaldi.py _get_strided)
}

extern "C" int cf_fbank(cf_handle* h, const float* pcm, int64_t n_samples, int sample_rate, int num_mel_bins, int frame_length_ms,
                        int frame_shift_ms, float* out, void* stream) {
  if (!h || !pcm || !out) return fail(h, CF_ERR_INVALID, "cf_fbank: null argument");
  const int flen = sample_rate * frame_length_ms / 1000, fshift = sample_rate * frame_shift_ms / 1000;
  if (flen <= FB_PAD / 2 || flen > FB_PAD || fshift <= 0 || num_mel_bins <= 0 || num_mel_bins > FB_MAX_BINS)
    return fail(h, CF_ERR_INVALID, "cf_fbank: the frame must pad to 512 samples (e.g. 25 ms at 16 kHz) and num_mel_bins <= 96");
  const int64_t T = cf_fbank_num_frames(n_samples, sample_rate, frame_length_ms, frame_shift_ms);
  if (T == 0) return CF_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DeviceGuard guard(h->device);
  const cf_handle::FbankTables* tb = nullptr;
  for (auto& t : h->fbank_tables)
    if (t.sr == sample_rate && t.bins == num_mel_bins && t.flen == flen && t.fshift == fshift) tb = &t;
  if (!tb) {
    // povey window and mel filters exactly as torchaudio builds them (kaldi.py _feature_window_function, get_mel_banks)
    std::vector<float> win(flen);
    for (int i = 0; i < flen; ++i) win[i] = float(pow(0.5 - 0.5 * cos(2.0 * M_PI * double(i) / double(flen - 1)), 0.85));
    const int nfb = FB_PAD / 2;
    const double nyq = 0.5 * sample_rate, width = double(sample_rate) / FB_PAD;
    auto mel = [](double f) { return 1127.0 * log(1.0 + f / 700.0); };
    const double lo = mel(20.0), hi = mel(nyq), delta = (hi - lo) / (num_mel_bins + 1);
    std::vector<float> w;
    std::vector<int2> rng(num_mel_bins);
    std::vector<int> cnt(num_mel_bins);
    for (int b = 0; b < num_mel_bins; ++b) {
      const double left = lo + b * delta, center = left + delta, right = center + delta;
      int first = -1, last = -1;
      std::vector<float> row(nfb);
      for (int i = 0; i < nfb; ++i) {
        const double m = mel(width * i);
        const double v = std::max(0.0, std::min((m - left) / (center - left), (right - m) / (right - center)));
        row[i] = float(v);
        if (v > 0.0) { if (first < 0) first = i; last = i; }
      }
      if (first < 0) { first = 0; last = -1; }
      rng[b] = make_int2(first, int(w.size()));
      cnt[b] = last - first + 1;
      for (int i = first; i <= last; ++i) w.push_back(row[i]);
    }
    if (w.empty()) w.push_back(0.f);
    cf_handle::FbankTables t;
    t.sr = sample_rate; t.bins = num_mel_bins; t.flen = flen; t.fshift = fshift;
    CF_CUDA(h, cudaMalloc(&t.window, win.size() * sizeof(float)));
    CF_CUDA(h, cudaMalloc(&t.mel_w, w.size() * sizeof(float)));
    CF_CUDA(h, cudaMalloc(&t.mel_rng, rng.size() * sizeof(int2)));
    CF_CUDA(h, cudaMalloc(&t.mel_cnt, cnt.size() * sizeof(int)));
    CF_CUDA(h, cudaMemcpy(t.window, win.data(), win.size() * sizeof(float), cudaMemcpyHostToDevice));
    CF_CUDA(h, cudaMemcpy(t.mel_w, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice));
    CF_CUDA(h, cudaMemcpy(t.mel_rng, rng.data(), rng.size() * sizeof(int2), cudaMemcpyHostToDevice));
    CF_CUDA(h, cudaMemcpy(t.mel_cnt, cnt.data(), cnt.size() * sizeof(int), cudaMemcpyHostToDevice));
    h->fbank_tables.push_back(t);
    tb = &h->fbank_tables.back();
  }
  FbankParams q{};
  q.pcm = pcm; q.out = out; q.window = tb->window; q.mel_w = tb->mel_w; q.mel_rng = tb->mel_rng; q.mel_cnt = tb->mel_cnt;
  q.n_frames = T; q.frame_len = flen; q.frame_shift = fshift; q.num_bins = num_mel_bins; q.preemph = 0.97f;
  const long long blocks_needed = (T + FB_WARPS - 1) / FB_WARPS;
  const unsigned grid = unsigned(std::min<long long>(blocks_needed, 8LL * h->num_sms));
  fbank_kernel<<<grid, FB_WARPS * 32, 0, st>>>(q);
  ++cf::g_kernel_launches;
  CF_CUDA(h, cudaGetLastError());
  return CF_OK;
}

// --------------------------------------------------------------------------------------------------------------------
// CTC head
// --------------------------------------------------------------------------------------------------------------------
extern "C" size_t cf_ctc_workspace_bytes(const cf_handle* h, int64_t rows, int enc_dtype) {
  if (!h || rows <= 0) return 256;
  const size_t nt = 2 * size_t((h->cfg.vocab + 255) / 256);
  const size_t cast = enc_dtype == CF_F32 ? size_t(rows) * h->cfg.d_model * sizeof(bf16) + 256 : 0;
  return 3 * (size_t(rows) * nt * 4 + 256) + cast + 256;
}

extern "C" int cf_ctc_greedy(cf_handle* h, const void* enc, int enc_dtype, int64_t rows, int64_t* tokens_out, float* margin_out,
                             float* logp_out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!h || !enc || !tokens_out || !workspace) return fail(h, CF_ERR_INVALID, "cf_ctc_greedy: null argument");
  if (enc_dtype != CF_F32 && enc_dtype != CF_BF16) return fail(h, CF_ERR_INVALID, "cf_ctc_greedy: bad enc_dtype");
  if (!h->finalized || h->cfg.vocab <= 0 || !h->ctc_w) return fail(h, CF_ERR_STATE, "cf_ctc_greedy: no CTC head loaded");
  if (rows <= 0) return CF_OK;
  if (workspace_bytes < cf_ctc_workspace_bytes(h, rows, enc_dtype)) return fail(h, CF_ERR_WORKSPACE, "cf_ctc_greedy: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DeviceGuard guard(h->device);
  const int d = h->cfg.d_model, V = h->cfg.vocab;
  const int nt = 2 * ((V + 255) / 256);
  Carver cv(workspace);
  float* best = cv.take<float>(size_t(rows) * nt);
  float* second = cv.take<float>(size_t(rows) * nt);
  int* index = cv.take<int>(size_t(rows) * nt);
  const void* enc_bf16 = enc;
  if (enc_dtype == CF_F32) {
    bf16* tmp = cv.take<bf16>(size_t(rows) * d);
    const long long n4 = (long long)rows * d / 4;
    cast_f32_bf16_kernel<<<unsigned(std::min<long long>((n4 + 255) / 256, 16LL * h->num_sms)), 256, 0, st>>>(
        static_cast<const float*>(enc), tmp, n4);
    ++cf::g_kernel_launches;
    CF_CUDA(h, cudaGetLastError());
    enc_bf16 = tmp;
  }
  std::string err;
  GemmLaunch g{};
  g.A = enc_bf16; g.lda = d; g.B = h->ctc_w; g.ldb = d; g.M = int(rows); g.N = V; g.K = d; g.epi = EPI_ARGMAX;
  g.act = ACT_NONE; g.ep.bias = h->ctc_b; g.ep.part_best = best; g.ep.part_second = second; g.ep.part_index = index;
  if (!launch_gemm(g, h->num_sms, st, &err)) return fail(h, CF_ERR_CUDA, "cf_ctc_greedy: " + err);
  ctc_reduce_kernel<<<unsigned((rows + 255) / 256), 256, 0, st>>>(best, second, index, nt, rows,
                                                                   reinterpret_cast<long long*>(tokens_out), margin_out);
  CF_CUDA(h, cudaGetLastError());
  if (logp_out) {
    GemmLaunch q{};
    q.A = enc_bf16; q.lda = d; q.B = h->ctc_w; q.ldb = d; q.M = int(rows); q.N = V; q.K = d; q.epi = EPI_F32;
    q.act = ACT_NONE; q.ep.bias = h->ctc_b; q.out = logp_out; q.ldo = V; q.ep.alpha = 1.0f;
    if (V % 4 != 0) return fail(h, CF_ERR_INVALID, "cf_ctc_greedy: logp_out needs vocab % 4 == 0");
    if (!launch_gemm(q, h->num_sms, st, &err)) return fail(h, CF_ERR_CUDA, "cf_ctc_greedy: " + err);
    log_softmax_rows_kernel<<<unsigned((rows + 7) / 8), 256, 0, st>>>(logp_out, rows, V);
    ++cf::g_kernel_launches;
    CF_CUDA(h, cudaGetLastError());
  }
  return CF_OK;
}

extern "C" size_t cf_ctc_compact_workspace_bytes(int64_t rows) {
  const int64_t blocks = (rows + CTC_COMPACT_TILE - 1) / CTC_COMPACT_TILE;
  return size_t(blocks > 0 ? blocks : 1) * sizeof(int) + 256;
}

extern "C" int cf_ctc_compact(const int64_t* tokens, int64_t rows, const int64_t* seg_start, const int32_t* seg_len, int n_seg,
                              int mode, int64_t blank_id, int64_t* out_tokens, int32_t* out_frames, int64_t* out_offsets,
                              void* workspace, size_t workspace_bytes, void* stream) {
  if (!tokens || !seg_start || !seg_len || !out_tokens || !out_frames || !out_offsets || !workspace)
    return fail(nullptr, CF_ERR_INVALID, "cf_ctc_compact: null argument");
  if (rows < 0 || n_seg < 0 || (mode != 0 && mode != 1)) return fail(nullptr, CF_ERR_INVALID, "cf_ctc_compact: bad size or mode");
  if (workspace_bytes < cf_ctc_compact_workspace_bytes(rows)) return fail(nullptr, CF_ERR_WORKSPACE, "cf_ctc_compact: workspace too small");
  DeviceGuard guard(device_of(tokens));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (rows == 0) {   // nothing to scan: every offset is 0
    cudaError_t e = cudaMemsetAsync(out_offsets, 0, sizeof(int64_t) * (size_t(n_seg) + 1), st);
    if (e != cudaSuccess) return fail(nullptr, CF_ERR_CUDA, std::string("cf_ctc_compact: ") + cudaGetErrorString(e));
    return CF_OK;
  }
  CtcCompactParams p{};
  p.tokens = reinterpret_cast<const long long*>(tokens); p.rows = rows;
  p.seg_start = reinterpret_cast<const long long*>(seg_start); p.seg_len = seg_len; p.n_seg = n_seg; p.mode = mode;
  p.blank = blank_id; p.out_tokens = reinterpret_cast<long long*>(out_tokens); p.out_frames = out_frames;
  p.out_offsets = reinterpret_cast<long long*>(out_offsets); p.block_counts = static_cast<int*>(workspace);
  const unsigned blocks = unsigned((rows + CTC_COMPACT_TILE - 1) / CTC_COMPACT_TILE);
  ctc_compact_count_kernel<<<blocks, CTC_COMPACT_THREADS, 0, st>>>(p);
  ctc_compact_scatter_kernel<<<blocks, CTC_COMPACT_THREADS, 0, st>>>(p);
  cf::g_kernel_launches += 2;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(nullptr, CF_ERR_CUDA, std::string("cf_ctc_compact: ") + cudaGetErrorString(e));
  return CF_OK;
}

// --------------------------------------------------------------------------------------------------------------------
// transducer greedy search (transducer.cuh)
// --------------------------------------------------------------------------------------------------------------------
struct cf_rnnt {
  cf_rnnt_config cfg{};
  int device = 0;
  bool finalized = false;
  std::string err;
  std::map<std::string, std::vector<float>> host;     // checkpoint tensors until finalize
  std::map<std::string, std::vector<int64_t>> shape;
  float *embed = nullptr, *enc_w = nullptr, *enc_b = nullptr, *wc = nullptr, *bc = nullptr, *woT = nullptr, *wo = nullptr, *bo = nullptr;
  int opt_persistent = 1;          // the whole search as one persistent cooperative kernel when the model fits (cf_rnnt_set_option)
  std::vector<float*> w_ih, w_hh, b_ih, b_hh;
  std::vector<void*> owned;
  ~cf_rnnt() { for (void* q : owned) cudaFree(q); }
};
static int rfail(cf_rnnt* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  g_last_error = msg;
  return code;
}
#define CF_RCUDA(h, call)                                                                                        \
  do {                                                                                                           \
    cudaError_t e__ = (call);                                                                                    \
    if (e__ != cudaSuccess) return rfail(h, CF_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));   \
  } while (0)

extern "C" const char* cf_rnnt_last_error(const cf_rnnt* h) { return h ? h->err.c_str() : g_last_error.c_str(); }

extern "C" int cf_rnnt_create(const cf_rnnt_config* cfg, int device, cf_rnnt** out) {
  if (!cfg || !out) return rfail(nullptr, CF_ERR_INVALID, "cf_rnnt_create: null argument");
  *out = nullptr;
  const cf_rnnt_config& c = *cfg;
  if (c.vocab <= 1 || c.layers < 1 || c.layers > 8 || c.blank < 0 || c.blank >= c.vocab)
    return rfail(nullptr, CF_ERR_INVALID, "cf_rnnt_create: bad vocab / layers / blank");
  if (c.embed <= 0 || c.hidden <= 0 || c.pred_out <= 0 || c.enc_dim <= 0 || c.join_dim <= 0 || c.embed % 4 || c.hidden % 4 ||
      c.join_dim % 4 || c.join_dim > 1024)
    return rfail(nullptr, CF_ERR_INVALID, "cf_rnnt_create: embed / hidden / join_dim must be positive multiples of 4 (join_dim <= 1024)");
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n)
    return rfail(nullptr, CF_ERR_CUDA, "cf_rnnt_create: no such CUDA device (there is no CPU fallback)");
  cf_rnnt* h = new cf_rnnt();
  h->cfg = c; h->device = device;
  *out = h;
  return CF_OK;
}
extern "C" void cf_rnnt_destroy(cf_rnnt* h) {
  if (!h) return;
  DeviceGuard guard(h->device);
  delete h;
}

extern "C" int cf_rnnt_set_option(cf_rnnt* h, const char* name, int value) {
  if (!h || !name) return rfail(h, CF_ERR_INVALID, "cf_rnnt_set_option: null argument");
  if (std::string(name) == "persistent") { h->opt_persistent = value != 0; return CF_OK; }
  return rfail(h, CF_ERR_INVALID, std::string("cf_rnnt_set_option: unknown option ") + name);
}

extern "C" int cf_rnnt_load_tensor(cf_rnnt* h, const char* key, const float* host_f32, int ndim, const int64_t* shape) {
  if (!h || !key || !host_f32 || ndim < 1 || ndim > 2 || !shape) return rfail(h, CF_ERR_INVALID, "cf_rnnt_load_tensor: bad argument");
  if (h->finalized) return rfail(h, CF_ERR_STATE, "cf_rnnt_load_tensor: weights already finalized");
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) n *= size_t(shape[i]);
  h->host[key].assign(host_f32, host_f32 + n);
  h->shape[key].assign(shape, shape + ndim);
  return CF_OK;
}

extern "C" int cf_rnnt_finalize_weights(cf_rnnt* h) {
  if (!h) return rfail(nullptr, CF_ERR_INVALID, "cf_rnnt_finalize_weights: null handle");
  if (h->finalized) return CF_OK;
  const cf_rnnt_config& c = h->cfg;
  DeviceGuard guard(h->device);
  auto need = [&](const std::string& k, std::vector<int64_t> want, const std::vector<float>** out) -> bool {
    auto it = h->host.find(k);
    if (it == h->host.end()) { h->err = "cf_rnnt_finalize_weights: missing tensor " + k; return false; }
    if (h->shape[k] != want) { h->err = "cf_rnnt_finalize_weights: wrong shape for " + k; return false; }
    *out = &it->second;
    return true;
  };
  auto upload = [&](const float* src, size_t n, float** dst) -> bool {
    void* d = nullptr;
    if (cudaMalloc(&d, n * sizeof(float)) != cudaSuccess) { h->err = "cf_rnnt_finalize_weights: cudaMalloc failed"; return false; }
    h->owned.push_back(d);
    if (cudaMemcpy(d, src, n * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) { h->err = "cf_rnnt_finalize_weights: copy failed"; return false; }
    *dst = static_cast<float*>(d);
    return true;
  };
#define RN_NEED(key, shape_, var) const std::vector<float>* var; if (!need(key, shape_, &var)) { g_last_error = h->err; return CF_ERR_STATE; }
#define RN_UP(src, n, dst) if (!upload(src, n, dst)) { g_last_error = h->err; return CF_ERR_CUDA; }
  const int64_t V = c.vocab, E = c.embed, H = c.hidden, P = c.pred_out, D = c.enc_dim, J = c.join_dim;
  RN_NEED("predictor.embed.weight", (std::vector<int64_t>{V, E}), emb);
  RN_UP(emb->data(), emb->size(), &h->embed);
  h->w_ih.resize(c.layers); h->w_hh.resize(c.layers); h->b_ih.resize(c.layers); h->b_hh.resize(c.layers);
  for (int l = 0; l < c.layers; ++l) {
    const std::string sfx = "_l" + std::to_string(l);
    const int64_t in = l == 0 ? E : H;
    RN_NEED("predictor.rnn.weight_ih" + sfx, (std::vector<int64_t>{4 * H, in}), wih);
    RN_NEED("predictor.rnn.weight_hh" + sfx, (std::vector<int64_t>{4 * H, H}), whh);
    RN_NEED("predictor.rnn.bias_ih" + sfx, (std::vector<int64_t>{4 * H}), bih);
    RN_NEED("predictor.rnn.bias_hh" + sfx, (std::vector<int64_t>{4 * H}), bhh);
    RN_UP(wih->data(), wih->size(), &h->w_ih[l]); RN_UP(whh->data(), whh->size(), &h->w_hh[l]);
    RN_UP(bih->data(), bih->size(), &h->b_ih[l]); RN_UP(bhh->data(), bhh->size(), &h->b_hh[l]);
  }
  RN_NEED("predictor.projection.weight", (std::vector<int64_t>{P, H}), pw);
  RN_NEED("predictor.projection.bias", (std::vector<int64_t>{P}), pb);
  RN_NEED("joint.enc_ffn.weight", (std::vector<int64_t>{J, D}), ew);
  RN_NEED("joint.enc_ffn.bias", (std::vector<int64_t>{J}), eb);
  RN_NEED("joint.pred_ffn.weight", (std::vector<int64_t>{J, P}), fw);
  RN_NEED("joint.pred_ffn.bias", (std::vector<int64_t>{J}), fb);
  RN_NEED("joint.ffn_out.weight", (std::vector<int64_t>{V, J}), ow);
  RN_NEED("joint.ffn_out.bias", (std::vector<int64_t>{V}), ob);
  RN_UP(ew->data(), ew->size(), &h->enc_w); RN_UP(eb->data(), eb->size(), &h->enc_b); RN_UP(ob->data(), ob->size(), &h->bo);
  {  // pred_ffn o projection composed in fp64 (predictor.py:205 + joint.py:88)
    std::vector<float> wc(size_t(J) * H), bcv(J);
    std::vector<double> row(H);
    for (int64_t r = 0; r < J; ++r) {
      std::fill(row.begin(), row.end(), 0.0);
      double bacc = (*fb)[r];
      for (int64_t q = 0; q < P; ++q) {
        const double f = (*fw)[r * P + q];
        bacc += f * (*pb)[q];
        const float* pr = pw->data() + q * H;
        for (int64_t k = 0; k < H; ++k) row[k] += f * pr[k];
      }
      for (int64_t k = 0; k < H; ++k) wc[r * H + k] = float(row[k]);
      bcv[r] = float(bacc);
    }
    RN_UP(wc.data(), wc.size(), &h->wc); RN_UP(bcv.data(), bcv.size(), &h->bc);
    std::vector<float> wt(size_t(J) * V);
    for (int64_t v = 0; v < V; ++v)
      for (int64_t k = 0; k < J; ++k) wt[((k >> 2) * V + v) * 4 + (k & 3)] = (*ow)[v * J + k];   // [J / 4][V][4]
    RN_UP(wt.data(), wt.size(), &h->woT);
    RN_UP(ow->data(), ow->size(), &h->wo);
  }
#undef RN_NEED
#undef RN_UP
  h->host.clear(); h->shape.clear();
  h->finalized = true;
  return CF_OK;
}

struct RnntWs {
  RnntState s; float* E; long long* seg_start; int* seg_len; size_t bytes; size_t state_floats; int n_vtiles;
  int *list_b, *list_cur, *list_tok, *cnt2, *rem2;     // [2][B] x 3, [2], [2]: double-buffered by iteration parity (fused control)
  unsigned* barrier; unsigned* jdone; int* iterations; int n_vtiles32;   // persistent kernel: grid-barrier counter, per-utterance tile
                                                                         // arrivals, iteration count, vocabulary tiles of 32 entries
};
static RnntWs rnnt_carve(const cf_rnnt_config& c, int64_t rows, int B, void* base) {
  Carver cv(base);
  RnntWs w{};
  w.n_vtiles = (c.vocab + RNNT_JV - 1) / RNNT_JV;
  w.n_vtiles32 = (c.vocab + 31) / 32;
  w.state_floats = size_t(2) * c.layers * B * c.hidden;
  w.E = cv.take<float>(size_t(rows) * c.join_dim);
  w.seg_start = cv.take<long long>(B); w.seg_len = cv.take<int>(B);
  w.s.t = cv.take<int>(B); w.s.step = cv.take<int>(B); w.s.token = cv.take<int>(B); w.s.cur = cv.take<int>(B);
  w.s.count = cv.take<int>(B); w.s.active = cv.take<int>(B); w.s.act_cur = cv.take<int>(B); w.s.act_tok = cv.take<int>(B); w.s.need_g = cv.take<int>(B);
  w.s.n_active = cv.take<int>(1); w.s.remaining = cv.take<int>(1); w.s.overflow = cv.take<int>(1);
  w.s.h = cv.take<float>(w.state_floats); w.s.c = cv.take<float>(w.state_floats);
  w.s.g = cv.take<float>(size_t(B) * c.join_dim);
  w.s.part_val = cv.take<float>(size_t(B) * RNNT_FB * w.n_vtiles32);
  w.s.part_idx = cv.take<int>(size_t(B) * RNNT_FB * w.n_vtiles32);
  w.barrier = cv.take<unsigned>(1); w.iterations = cv.take<int>(1); w.jdone = cv.take<unsigned>(B);
  w.list_b = cv.take<int>(size_t(2) * B); w.list_cur = cv.take<int>(size_t(2) * B); w.list_tok = cv.take<int>(size_t(2) * B);
  w.cnt2 = cv.take<int>(2); w.rem2 = cv.take<int>(2);
  w.bytes = cv.off + 256;
  return w;
}
extern "C" size_t cf_rnnt_workspace_bytes(const cf_rnnt* h, int64_t rows, int n_utt) {
  if (!h || rows < 0 || n_utt < 0) return 0;
  return rnnt_carve(h->cfg, rows, n_utt > 0 ? n_utt : 1, nullptr).bytes;
}

extern "C" int cf_rnnt_greedy(cf_rnnt* h, const float* enc_f32, int64_t rows, const int64_t* seg_start, const int32_t* seg_len,
                              int n_utt, int n_steps, int capacity, int64_t* out_tokens, int32_t* out_frames,
                              int32_t* out_counts, int64_t* iterations_out, void* workspace, size_t workspace_bytes,
                              void* stream) {
  if (!h || !enc_f32 || !seg_start || !seg_len || !out_tokens || !out_frames || !out_counts || !workspace)
    return rfail(h, CF_ERR_INVALID, "cf_rnnt_greedy: null argument");
  if (!h->finalized) return rfail(h, CF_ERR_STATE, "cf_rnnt_greedy: weights not finalized");
  if (n_utt <= 0 || n_steps <= 0 || capacity <= 0 || rows < 0) return rfail(h, CF_ERR_INVALID, "cf_rnnt_greedy: bad sizes");
  for (int b = 0; b < n_utt; ++b)
    if (seg_len[b] < 0 || seg_start[b] < 0 || seg_start[b] + seg_len[b] > rows)
      return rfail(h, CF_ERR_INVALID, "cf_rnnt_greedy: utterance rows outside the encoder buffer");
  const cf_rnnt_config& c = h->cfg;
  RnntWs w = rnnt_carve(c, rows, n_utt, workspace);
  if (workspace_bytes < w.bytes) return rfail(h, CF_ERR_WORKSPACE, "cf_rnnt_greedy: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DeviceGuard guard(h->device);
  CF_RCUDA(h, cudaMemcpyAsync(w.seg_start, seg_start, sizeof(int64_t) * n_utt, cudaMemcpyHostToDevice, st));
  CF_RCUDA(h, cudaMemcpyAsync(w.seg_len, seg_len, sizeof(int32_t) * n_utt, cudaMemcpyHostToDevice, st));
  if (rows > 0) {
    dim3 grid(unsigned((c.join_dim + 63) / 64), unsigned((rows + 63) / 64));
    rnnt_linear_f32_kernel<<<grid, 256, 0, st>>>(enc_f32, h->enc_w, h->enc_b, w.E, rows, c.join_dim, c.enc_dim);
    ++cf::g_kernel_launches;
  }
  rnnt_init_kernel<<<64, 256, 0, st>>>(w.s, w.seg_len, out_counts, n_utt, c.blank, w.state_floats, w.list_b, w.list_cur, w.list_tok,
                                       w.cnt2, w.rem2);
  ++cf::g_kernel_launches;
  CF_RCUDA(h, cudaGetLastError());
  // ---- the whole search as one persistent cooperative kernel, when the model's weight slices fit one CTA per SM
  if (h->opt_persistent) {
    int dev = 0, sms = 0, coop = 0, max_smem = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const int NV = w.n_vtiles32;
    const size_t smem = sms > 0 ? rnnt_persist_smem_bytes(c.layers, c.embed, c.hidden, c.join_dim, sms) : 0;
    if (coop && sms >= NV && c.join_dim % 32 == 0 && c.layers <= 8 && n_utt <= RNNT_PMAXB && smem <= size_t(max_smem)) {
      std::string aerr;
      if (!ensure_smem_optin(rnnt_persistent_kernel, smem, &aerr, "rnnt_persistent")) return rfail(h, CF_ERR_CUDA, aerr);
      int per_sm = 0;
      CF_RCUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rnnt_persistent_kernel, 256, smem));
      if (per_sm >= 1) {
        RnntPersistParams pp{};
        pp.embed = h->embed;
        for (int l = 0; l < c.layers; ++l) { pp.w_ih[l] = h->w_ih[l]; pp.w_hh[l] = h->w_hh[l]; pp.b_ih[l] = h->b_ih[l]; pp.b_hh[l] = h->b_hh[l]; }
        pp.Wc = h->wc; pp.bc = h->bc; pp.Wo = h->wo; pp.bo = h->bo; pp.E = w.E; pp.seg_start = w.seg_start; pp.seg_len = w.seg_len;
        pp.layers = c.layers; pp.Em = c.embed; pp.H = c.hidden; pp.J = c.join_dim; pp.V = c.vocab; pp.B = n_utt; pp.n_steps = n_steps;
        pp.cap = capacity; pp.blank = c.blank; pp.NV = NV; pp.UG = sms / NV;
        pp.list_b = w.list_b; pp.list_cur = w.list_cur; pp.list_tok = w.list_tok; pp.cnt2 = w.cnt2; pp.rem2 = w.rem2;
        pp.out_tokens = reinterpret_cast<long long*>(out_tokens); pp.out_frames = out_frames; pp.out_counts = out_counts;
        pp.barrier = w.barrier; pp.iterations = w.iterations; pp.jdone = w.jdone;
        int64_t max_len = 0;
        for (int b = 0; b < n_utt; ++b) max_len = seg_len[b] > max_len ? seg_len[b] : max_len;
        pp.max_iters = max_len * (int64_t(n_steps) + 1) + 2;       // every iteration emits a symbol or moves a frame forward
        CF_RCUDA(h, cudaMemsetAsync(w.barrier, 0, sizeof(unsigned), st));
        CF_RCUDA(h, cudaMemsetAsync(w.iterations, 0, sizeof(int), st));
        CF_RCUDA(h, cudaMemsetAsync(w.jdone, 0, sizeof(unsigned) * size_t(n_utt), st));
        RnntState state = w.s;
        void* args[] = {&pp, &state};
        CF_RCUDA(h, cudaLaunchCooperativeKernel(reinterpret_cast<void*>(rnnt_persistent_kernel), dim3(unsigned(sms)), dim3(256), args, smem, st));
        ++cf::g_kernel_launches;
        int host[3] = {0, 0, 0};   // iterations, overflow, remaining of the last iteration's parity
        CF_RCUDA(h, cudaMemcpyAsync(&host[0], w.iterations, sizeof(int), cudaMemcpyDeviceToHost, st));
        CF_RCUDA(h, cudaMemcpyAsync(&host[1], w.s.overflow, sizeof(int), cudaMemcpyDeviceToHost, st));
        CF_RCUDA(h, cudaStreamSynchronize(st));
        int rem = 0;
        CF_RCUDA(h, cudaMemcpy(&rem, w.rem2 + ((host[0] - 1) & 1), sizeof(int), cudaMemcpyDeviceToHost));
        if (iterations_out) *iterations_out = host[0];
        if (host[1]) return rfail(h, CF_ERR_WORKSPACE, "cf_rnnt_greedy: an utterance emitted more than `capacity` symbols");
        if (rem != 0) return rfail(h, CF_ERR_STATE, "cf_rnnt_greedy: search did not terminate (internal error)");
        return CF_OK;
      }
    }
  }
  RnntJointParams jp{};
  jp.E = w.E; jp.WoT = h->woT; jp.bo = h->bo; jp.seg_start = w.seg_start; jp.seg_len = w.seg_len; jp.J = c.join_dim; jp.V = c.vocab;
  jp.n_vtiles = w.n_vtiles; jp.Wc = h->wc; jp.bc = h->bc; jp.H = c.hidden; jp.layers = c.layers;
  jp.out_tokens = reinterpret_cast<long long*>(out_tokens); jp.out_frames = out_frames; jp.out_counts = out_counts;
  jp.n_steps = n_steps; jp.cap = capacity; jp.blank = c.blank;
#ifdef CF_ABLATION
  { const char* e = getenv("CF_RNNT_DEBUG"); jp.debug = e ? atoi(e) : 0; }
#endif
  RnntDecideParams dp{w.seg_len, reinterpret_cast<long long*>(out_tokens), out_frames, out_counts, n_utt, w.n_vtiles, n_steps,
                      capacity, c.blank};
  const unsigned tiles = unsigned(std::min((n_utt + RNNT_BT - 1) / RNNT_BT, 8));
  const size_t jsmem = rnnt_joint_smem_bytes(c.join_dim);
  {
    std::string aerr;
    if (!ensure_smem_optin(rnnt_joint_kernel<false>, rnnt_joint_smem_bytes(1024), &aerr, "rnnt_joint") ||
        !ensure_smem_optin(rnnt_joint_kernel<true>, rnnt_joint_smem_bytes(1024), &aerr, "rnnt_joint"))
      return rfail(h, CF_ERR_CUDA, aerr);
  }
  // cluster-fused joint (projection + joint + greedy control in one launch) whenever the vocabulary makes 2 / 4 / 8 tiles
  const bool fused_proj = (w.n_vtiles == 2 || w.n_vtiles == 4 || w.n_vtiles == 8);
  cudaError_t launch_err = cudaSuccess;
  auto iteration = [&](int64_t it) {
    const int par = int(it & 1);
    for (int l = 0; l < c.layers; ++l) {
      RnntLstmParams lp{h->w_ih[l], h->w_hh[l], h->b_ih[l], h->b_hh[l], l == 0 ? h->embed : nullptr, l, l == 0 ? c.embed : c.hidden,
                        c.hidden, n_utt, c.layers,
                        fused_proj ? w.list_b + (par ^ 1) * n_utt : w.s.active, fused_proj ? w.list_cur + (par ^ 1) * n_utt : w.s.act_cur,
                        fused_proj ? w.list_tok + (par ^ 1) * n_utt : w.s.act_tok, fused_proj ? w.cnt2 + (par ^ 1) : nullptr};
      rnnt_lstm_kernel<<<dim3(unsigned((c.hidden + 3) / 4), tiles), 128, 0, st>>>(lp, w.s);
    }
    dim3 jg(unsigned(w.n_vtiles), unsigned(n_utt));
    if (fused_proj) {
      // one cluster per utterance: its vocabulary-tile CTAs compute and share the projection themselves
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = jg; cfg.blockDim = dim3(RNNT_JTHREADS); cfg.dynamicSmemBytes = jsmem; cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = unsigned(w.n_vtiles); at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      RnntJointParams jq = jp;
      jq.list_b = w.list_b + par * n_utt; jq.list_cur = w.list_cur + par * n_utt; jq.list_tok = w.list_tok + par * n_utt;
      jq.cnt = w.cnt2 + par; jq.rem = w.rem2 + par; jq.cnt_next = w.cnt2 + (par ^ 1); jq.rem_next = w.rem2 + (par ^ 1);
      cudaError_t e = cudaLaunchKernelEx(&cfg, rnnt_joint_kernel<true>, jq, w.s, n_utt);
      if (e != cudaSuccess) launch_err = e;
      cf::g_kernel_launches += c.layers + 1;
      return;                                   // the greedy control ran inside the clusters
    } else {
      rnnt_predproj_kernel<<<dim3(unsigned((c.join_dim + 3) / 4), tiles), 128, 0, st>>>(h->wc, h->bc, c.join_dim, c.hidden, n_utt,
                                                                                        c.layers, w.s);
      rnnt_joint_kernel<false><<<jg, RNNT_JTHREADS, jsmem, st>>>(jp, w.s, n_utt);
      cf::g_kernel_launches += c.layers + 3;
    }
    rnnt_decide_kernel<<<1, 256, 0, st>>>(dp, w.s);
  };
  // the host only learns how far the search is by reading `remaining`; a burst of iterations is enqueued between reads
  // (iterations after the last utterance finished are empty launches)
  int state[2] = {1, 0};   // remaining, overflow
  int64_t iters = 0, max_len = 0;
  for (int b = 0; b < n_utt; ++b) max_len = seg_len[b] > max_len ? seg_len[b] : max_len;
  // every iteration emits a symbol or moves at least one frame forward in each unfinished utterance
  const int64_t bound = max_len * (int64_t(n_steps) + 1) + 1;
  const int burst = 64;
  while (state[0] > 0) {
    if (iters > bound) return rfail(h, CF_ERR_STATE, "cf_rnnt_greedy: search did not terminate (internal error)");
    for (int i = 0; i < burst; ++i) iteration(iters + i);
    iters += burst;
    CF_RCUDA(h, launch_err);
    CF_RCUDA(h, cudaGetLastError());
    CF_RCUDA(h, cudaMemcpyAsync(&state[0], fused_proj ? w.rem2 + ((iters - 1) & 1) : w.s.remaining, sizeof(int), cudaMemcpyDeviceToHost, st));
    CF_RCUDA(h, cudaMemcpyAsync(&state[1], w.s.overflow, sizeof(int), cudaMemcpyDeviceToHost, st));
    CF_RCUDA(h, cudaStreamSynchronize(st));
  }
  if (iterations_out) *iterations_out = iters;
  if (state[1]) return rfail(h, CF_ERR_WORKSPACE, "cf_rnnt_greedy: an utterance emitted more than `capacity` symbols");
  return CF_OK;
}

// --------------------------------------------------------------------------------------------------------------------
// kernel-level entry points
// --------------------------------------------------------------------------------------------------------------------
static int current_sms() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

extern "C" int cf_op_gemm(const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K, int epi, int act,
                          const float* bias, const float* resid, int64_t ld_resid, float alpha, const int32_t* row_range,
                          int rows_per_chunk, void* out, int64_t ldo, float* part_best, float* part_second,
                          int32_t* part_index, int variant, void* stream) {
  DeviceGuard guard(device_of(A));
  if (!A || !B || !bias) return fail(nullptr, CF_ERR_INVALID, "cf_op_gemm: null argument");
  GemmLaunch g{};
  g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.M = M; g.N = N; g.K = K; g.epi = epi; g.act = act; g.out = out; g.ldo = ldo;
  g.ep.bias = bias; g.ep.resid = resid; g.ep.ld_resid = ld_resid; g.ep.alpha = alpha;
  g.ep.row_range = reinterpret_cast<const int2*>(row_range);
  g.ep.rows_per_chunk = rows_per_chunk > 0 ? rows_per_chunk : 1;
  g.ep.part_best = part_best; g.ep.part_second = part_second; g.ep.part_index = part_index;
  g.variant = variant;
  std::string err;
  if (!launch_gemm(g, current_sms(), static_cast<cudaStream_t>(stream), &err)) return fail(nullptr, CF_ERR_CUDA, err);
  return CF_OK;
}

extern "C" int cf_op_gemm_ln(const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K, const float* bias,
                             const float* resid, int64_t ld_resid, float alpha, const int32_t* row_range, int rows_per_chunk,
                             int mode, const float* ln1_w, const float* ln1_b, const float* ln2_w, const float* ln2_b,
                             float* x_out, int64_t ldx, void* y_out, int64_t ldy, const int32_t* row_limit, int rows_per_seq,
                             void* stream) {
  DeviceGuard guard(device_of(A));
  if (!A || !B) return fail(nullptr, CF_ERR_INVALID, "cf_op_gemm_ln: null argument");
  GemmLnLaunch g{};
  g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.M = M; g.N = N; g.K = K; g.bias = bias; g.resid = resid; g.ld_resid = ld_resid;
  g.alpha = alpha; g.row_range = reinterpret_cast<const int2*>(row_range); g.rows_per_chunk = rows_per_chunk;
  g.mode = mode & 15;
  g.variant = (mode & 16) ? 0 : ((mode & 32) ? 2 : ((mode & 64) ? 1 : -1));   // + 16: single-epilogue-group kernel, + 32: cluster of four, + 64: pair
  g.ln1_w = ln1_w; g.ln1_b = ln1_b; g.ln2_w = ln2_w; g.ln2_b = ln2_b; g.x_out = x_out; g.ldx = ldx; g.y_out = y_out; g.ldy = ldy;
  g.row_limit = row_limit; g.rows_per_seq = rows_per_seq;
  std::string err;
#ifdef CF_ABLATION
  long long* prof_dev = nullptr;
  const int prof_ctas = current_sms();
  if (getenv("CF_LN_PROF")) { cudaMalloc(&prof_dev, size_t(prof_ctas) * 48 * 8); cudaMemset(prof_dev, 0, size_t(prof_ctas) * 48 * 8); g.prof = prof_dev; }
#endif
  if (!launch_gemm_ln(g, current_sms(), static_cast<cudaStream_t>(stream), &err)) return fail(nullptr, CF_ERR_CUDA, err);
#ifdef CF_ABLATION
  if (prof_dev) {
    std::vector<long long> hp(size_t(prof_ctas) * 48);
    cudaMemcpy(hp.data(), prof_dev, hp.size() * 8, cudaMemcpyDeviceToHost);
    cudaFree(prof_dev);
    static const char* names[16] = {"producer total", "producer wait empty", "mma total", "mma wait tempty", "mma wait full", "P1 total",
                                    "P1 wait tfull", "P1 wait resid", "P1 wait credit", "P1 named barrier", "P2 total", "P2 wait stats",
                                    "P2 wait stats2", "P1 issuer store-read wait", "P1 tmem ld wait", "P1 barrier (warp 1)"};
    fprintf(stderr, "gemm_ln_split profile (cycles, mean over CTAs; M=%d K=%d mode=%d):\n", M, K, g.mode);
    for (int f = 0; f < 16; ++f) {
      double sum = 0; int n = 0;
      for (int c = 0; c < prof_ctas; ++c) if (hp[size_t(c) * 16 + 0] > 0) { sum += double(hp[size_t(c) * 16 + f]); ++n; }
      fprintf(stderr, "  %-22s %12.0f\n", names[f], n ? sum / n : 0.0);
    }
    static const char* segn[9] = {"issue next resid", "bias ldg + resid wait", "tmem ld wait", "add + stage", "tmem st", "stats", "fence",
                                  "named barrier", "tma store"};
    for (int wsel = 0; wsel < 2; ++wsel)
      for (int f = 0; f < 9; ++f) {
        double sum = 0;
        for (int c = 0; c < prof_ctas; ++c) sum += double(hp[(size_t(wsel + 1) * prof_ctas + c) * 16 + f]);
        fprintf(stderr, "  P1 warp %d seg %-22s %12.0f\n", wsel, segn[f], sum / prof_ctas);
      }
  }
#endif
  return CF_OK;
}

extern "C" int cf_op_ffn(const void* y_bf16, int64_t ldy_in, const void* w1, const float* b1, const void* w2, const float* b2, int M,
                         int d, int F, const float* resid, int64_t ld_resid, float alpha, int mode, const float* ln1_w,
                         const float* ln1_b, const float* ln2_w, const float* ln2_b, float* x_out, int64_t ldx, void* y_out,
                         int64_t ldy, void* stream) {
  DeviceGuard guard(device_of(y_bf16));
  FfnLaunch g{};
  g.Y = y_bf16; g.ldy_in = ldy_in; g.W1 = w1; g.b1 = b1; g.W2 = w2; g.b2 = b2; g.M = M; g.d = d; g.F = F; g.resid = resid;
  g.ld_resid = ld_resid; g.alpha = alpha; g.mode = mode; g.ln1_w = ln1_w; g.ln1_b = ln1_b; g.ln2_w = ln2_w; g.ln2_b = ln2_b;
  g.x_out = x_out; g.ldx = ldx; g.y_out = y_out; g.ldy = ldy;
  std::string err;
  if (!launch_ffn_fused(g, current_sms(), static_cast<cudaStream_t>(stream), &err)) return fail(nullptr, CF_ERR_CUDA, err);
  return CF_OK;
}

extern "C" int cf_op_layernorm(int mode, int d, const float* x_in, float* x_out, void* y_bf16, const float* w1,
                               const float* b1, const float* w2, const float* b2, int64_t rows, void* stream) {
  DeviceGuard guard(device_of(x_in));
  LnParams q{};
  q.x_in = x_in; q.x_out = x_out; q.y = static_cast<bf16*>(y_bf16); q.w1 = w1; q.b1 = b1; q.w2 = w2; q.b2 = b2;
  q.rows = rows; q.row_limit = nullptr; q.rows_per_seq = 1;
  std::string err;
  if (!run_layernorm(mode, d, q, static_cast<cudaStream_t>(stream), &err)) return fail(nullptr, CF_ERR_CUDA, err);
  return CF_OK;
}

extern "C" int cf_op_dwconv(int d, int kernel, const void* g_bf16, void* z_bf16, const float* w, const float* bias,
                            const float* ln_w, const float* ln_b, const int32_t* range, int c, int n_chunks, void* stream) {
  DeviceGuard guard(device_of(g_bf16));
  DwConvParams q{};
  q.g = static_cast<const bf16*>(g_bf16); q.z = static_cast<bf16*>(z_bf16); q.w = w; q.bias = bias; q.ln_w = ln_w;
  q.ln_b = ln_b; q.range = reinterpret_cast<const int2*>(range); q.c = c; q.n_chunks = n_chunks;
  std::string err;
  // the TMA-staged kernel is used when the caller guarantees c % 32 == 0 (g then holds n_chunks*c + 14 readable rows)
  if (!run_dwconv(d, kernel, q, static_cast<cudaStream_t>(stream), &err, (long long)n_chunks * c + 14, current_sms()))
    return fail(nullptr, CF_ERR_CUDA, err);
  return CF_OK;
}

extern "C" int cf_op_attention(int impl, const void* qkv_bf16, const void* pos_bf16, const int32_t* range, void* ctx_bf16,
                               int n_chunks, int c, int l, int r, int d, int heads, int prescaled, void* stream) {
  DeviceGuard guard(device_of(qkv_bf16));
  AttnParams a{};
  a.qkv = static_cast<const bf16*>(qkv_bf16); a.pos = static_cast<const bf16*>(pos_bf16);
  a.range = reinterpret_cast<const int2*>(range); a.ctx = static_cast<bf16*>(ctx_bf16);
  a.n_chunks = n_chunks; a.c = c; a.l = l; a.r = r; a.d = d; a.heads = heads; a.scale = 1.0f / sqrtf(float(d / heads));
  a.prescaled = prescaled;
  std::string err;
  if (!run_attention(impl, a, static_cast<cudaStream_t>(stream), &err)) return fail(nullptr, CF_ERR_CUDA, err);
  return CF_OK;
}

extern "C" int cf_op_frontend_conv(int impl, int d, const float* feats, const int64_t* chunk_feat_row, const int32_t* chunk_in_len,
                                   int n_chunks, int chunk_size, int feat_dim, const float* wpack, const float* cmvn_mean,
                                   const float* cmvn_istd, void* out_bf16, void* stream) {
  DeviceGuard guard(device_of(feats));
  if (!feats || !chunk_feat_row || !chunk_in_len || !wpack || !out_bf16) return fail(nullptr, CF_ERR_INVALID, "cf_op_frontend_conv: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  std::vector<ChunkSrc> cs(n_chunks);
  for (int g = 0; g < n_chunks; ++g) { cs[g].feat_row = chunk_feat_row[g]; cs[g].in_len = chunk_in_len[g]; cs[g].pad_ = 0; }
  ChunkSrc* dev = nullptr;
  CF_CUDA(nullptr, cudaMalloc(&dev, sizeof(ChunkSrc) * size_t(n_chunks)));
  CF_CUDA(nullptr, cudaMemcpyAsync(dev, cs.data(), sizeof(ChunkSrc) * size_t(n_chunks), cudaMemcpyHostToDevice, st));
  Fe1Params f1{};
  const int F1 = (feat_dim - 3) / 2 + 1;
  f1.feats = feats; f1.chunks = dev; f1.wpack = wpack; f1.cmvn_mean = cmvn_mean; f1.cmvn_istd = cmvn_istd;
  f1.out = static_cast<bf16*>(out_bf16); f1.n_chunks = n_chunks; f1.feat_dim = feat_dim; f1.T2 = 2 * chunk_size + 1;
  f1.F2 = (F1 - 3) / 2 + 1; f1.in_rows = 8 * (chunk_size - 1) + 15;
  std::string err;
  const bool ok = run_frontend_conv(impl, d, f1, current_sms(), st, &err);
  cudaStreamSynchronize(st);
  cudaFree(dev);
  if (!ok) return fail(nullptr, CF_ERR_CUDA, err);
  return CF_OK;
}
