/*
 * chunkformer_b200 — C ABI of the B200-native (sm_100a) ChunkFormer encoder hot path.
 *
 * The reference (ishine/chunkformer) has no FFI / plugin interface: its boundary for this path is the Python method
 * surface of ChunkFormerEncoder plus the checkpoint layout.  Each entry point below names the reference
 * interface it stands in for (file:line relative to /root/reference).  The Python mirror of that surface
 * (chunkformer_b200/encoder.py, model.py) binds these symbols with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a negative cf_status and records a
 * message retrievable with cf_last_error(); no exception crosses the ABI; the caller owns all input/output buffers;
 * the library owns converted weights; a handle belongs to one device and is not thread-safe; all device work is
 * enqueued on the caller's stream (cudaStream_t passed as void*).  There is no CPU fallback: compute entry points fail
 * with CF_ERR_CUDA when no sm_100 device is present.
 */
#ifndef CHUNKFORMER_B200_H_
#define CHUNKFORMER_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define CF_API __attribute__((visibility("default")))
#else
#define CF_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cf_handle cf_handle;
typedef struct cf_plan cf_plan;

enum cf_status {
  CF_OK = 0,
  CF_ERR_INVALID = -1,     /* bad argument / unsupported geometry */
  CF_ERR_CUDA = -2,        /* CUDA runtime / driver error, or no device */
  CF_ERR_STATE = -3,       /* call order (weights not finalized, missing tensor, ...) */
  CF_ERR_WORKSPACE = -4    /* workspace too small */
};

enum cf_dtype { CF_F32 = 0, CF_BF16 = 1 };

/* Encoder geometry = ChunkFormerEncoder.__init__ arguments that shape the inference path
 * (chunkformer/modules/encoder.py:36-193) + CTC vocabulary (chunkformer/modules/ctc.py:44). */
typedef struct cf_config {
  int32_t d_model;    /* output_size (256 | 512) */
  int32_t heads;      /* attention_heads; d_model / heads in {64, 128} */
  int32_t ffn;        /* linear_units */
  int32_t layers;     /* num_blocks */
  int32_t kernel;     /* cnn_module_kernel (15) */
  int32_t vocab;      /* CTC output size, 0 = no CTC head */
  int32_t feat_dim;   /* input_size (80) */
  int32_t has_cmvn;   /* encoder.global_cmvn.{mean,istd} present */
  int32_t conv_norm;  /* cnn_module_norm of the conv module: 0 = layer_norm, 1 = batch_norm (eval mode: running statistics,
                         folded into the depthwise conv at weight load; convolution.py:83-89) */
} cf_config;

/* ---- lifetime ---------------------------------------------------------------------------------------------------- */
/* Replaces: init_speech_model(...) building ChunkFormerEncoder + CTC (chunkformer/utils/init_model.py:61-145). */
CF_API int cf_create(const cf_config* cfg, int device, cf_handle** out);
CF_API void cf_destroy(cf_handle* h);
/* Last error message of `h` (or of the last failed cf_create / plan call when h == NULL). Never NULL. */
CF_API const char* cf_last_error(const cf_handle* h);
CF_API const char* cf_version(void);
/* Number of kernels this library has launched in the process so far (bench.py reports the per-step count). */
CF_API long long cf_launch_count(void);
/* Rows of `out` the last cf_encode call on this handle wrote: the plan's rows, except for a multi-stream step
 * (cf_encode_streams) in compact mode (option "stream_compact", default 1; chunk sizes that divide 128, right context 0), where
 * every row-wise kernel works on the real chunk of every stream only and `out` holds n_streams x chunk rows, stream after
 * stream (the placeholder rows in front of each stream's chunk exist only inside the K / V and conv buffers). */
CF_API int64_t cf_encode_output_rows(const cf_handle* h);
/* Optional, before cf_encode: the feature buffer is still being filled by copies on another stream.  events[i] (cudaEvent_t)
 * fires when feature rows < rows_ready[i] are in place (rows_ready ascending).  cf_encode makes its stream wait only for the
 * rows each front-end slab reads, so the host-to-device copy overlaps the front-end; the list is consumed by that call.
 * Replaces the reference's blocking xs.to(device) (chunkformer_model.py:395-401). */
CF_API int cf_encode_feature_events(cf_handle* h, int n, const int64_t* rows_ready, void* const* events);
/* Kaldi-compatible log-mel filterbank on the device: replaces torchaudio.compliance.kaldi.fbank(waveform, num_mel_bins,
 * frame_length, frame_shift, dither=0.0, energy_floor=0.0, sample_frequency) as called right before the path
 * (chunkformer_model.py:307-315).  pcm: n_samples device floats in 16-bit range; out: [cf_fbank_num_frames(...), num_mel_bins]
 * device floats.  The frame must pad to 512 samples (25 ms at 16 kHz). */
CF_API int64_t cf_fbank_num_frames(int64_t n_samples, int sample_rate, int frame_length_ms, int frame_shift_ms);
CF_API int cf_fbank(cf_handle* h, const float* pcm, int64_t n_samples, int sample_rate, int num_mel_bins, int frame_length_ms,
                    int frame_shift_ms, float* out, void* stream);
/* Per-handle option.  "fused_layernorm" (default 1): the LayerNorm(s) behind every residual GEMM run in that GEMM's epilogue
 * (0 = stand-alone LayerNorm kernels; kept for A/B measurement).  "fused_ffn" (default 0: measured slower; needs
 * fused_layernorm): every feed-forward module runs as one kernel that keeps the hidden activation on chip (0 = w_1 GEMM, hidden
 * activation through global memory, w_2 GEMM).  "ln_split" (default -1 = 1): which residual GEMM + LayerNorm kernel runs: 1 = CTA
 * pair with the normalisation passes on their own warps, 2 = the same on a cluster of four with cta_group::2 MMAs, 0 = the first
 * version.  "gemm_pair" (default -1 = by shape): plain GEMMs on the CTA-pair kernel (1) or the one-CTA kernel (0).
 * "ffn_slab_rows" (default 0 = off): the two feed-forward GEMMs slab by slab.  All variants give the same results to fp32
 * summation order; the knobs exist for A/B measurements (tools/ab_option.py). */
CF_API int cf_set_option(cf_handle* h, const char* name, int value);
/* Measurement hook (bench.py roofline): CUDA events around every launch of the selected kernel families made by this handle's
 * cf_encode calls, on the launching stream, from cf_kernel_timing_begin until cf_kernel_timing_end(family), which returns the
 * summed device time and the number of launches of that family.  family_mask = OR of (1 << CF_FAMILY_*).  State lives on
 * the handle (a handle is used by one thread at a time). */
enum cf_kernel_family {
  CF_FAMILY_OTHER = 0,
  CF_FAMILY_FFN_W1 = 1,      /* FFN up-projection + SiLU (positionwise_feed_forward.py:59) */
  CF_FAMILY_FFN_W2 = 2,      /* FFN down-projection + 0.5 * residual */
  CF_FAMILY_FFN_FUSED = 3    /* w_1 -> SiLU -> w_2 -> residual in one kernel */
};
CF_API int cf_kernel_timing_begin(cf_handle* h, unsigned family_mask);
CF_API int cf_kernel_timing_end(cf_handle* h, int family, double* total_ms, int* launches);

/* ---- weights ----------------------------------------------------------------------------------------------------- */
/* Replaces: load_checkpoint -> model.load_state_dict(strict=False) (chunkformer/utils/checkpoint.py:26-41).
 * One call per checkpoint tensor, any order, keyed exactly as in the reference state_dict ("encoder.embed.conv.0.weight",
 * "encoder.encoders.3.self_attn.pos_bias_u", "ctc.ctc_lo.weight", ...). `data` is host memory, fp32, C-contiguous.
 * Keys outside encoder.* / ctc.ctc_lo.* are ignored (returns CF_OK), like strict=False. */
CF_API int cf_load_tensor(cf_handle* h, const char* key, const void* data, int dtype, int ndim, const int64_t* shape);
/* Converts to device layouts (bf16 GEMM operands, fused QKV, GLU-interleaved pointwise_conv1, permuted embed.out, packed
 * stencil weights). Fails with CF_ERR_STATE and names the first missing tensor if the checkpoint is incomplete. */
CF_API int cf_finalize_weights(cf_handle* h);

/* ---- packer / plan (pure host code: usable without a GPU) -------------------------------------------------------- */
/* Replaces: the masked-batch packer and bound tables of ChunkFormerEncoder.forward_parallel_chunk
 * (chunkformer/modules/encoder.py:538-612, 627-645).  lens[B] = input frames per utterance, offsets[B] = stream offsets
 * (NULL = zeros). feat_row_offsets[B] = first row of each utterance in the flat feature buffer (NULL = utterances are
 * concatenated back to back). */
CF_API int cf_plan_create(int chunk_size, int left_context, int right_context, int conv_kernel, int B, const int32_t* lens,
                   const int32_t* offsets, const int64_t* feat_row_offsets, cf_plan** out);
/* Replaces: the padded-batch geometry of ChunkFormerEncoder.forward_encoder (encoder.py:220-274, attention.py:337-383,
 * convolution.py:125-167): xs is (B, T, feat) row-major, lens[B] valid frames. Output rows per utterance: 1+(T-15)/8. */
CF_API int cf_plan_create_padded(int chunk_size, int left_context, int right_context, int conv_kernel, int B, int T,
                          const int32_t* lens, cf_plan** out);
CF_API void cf_plan_destroy(cf_plan* p);
CF_API int cf_plan_num_chunks(const cf_plan* p);               /* n = total chunks */
CF_API int cf_plan_rows(const cf_plan* p);                     /* n * chunk_size encoder rows computed */
/* n_chunks_out[B] (encoder.py:562) and enc_lens_out[B] = calc_length(lens) (subsampling.py:270-288). */
CF_API int cf_plan_tables(const cf_plan* p, int32_t* n_chunks_out, int32_t* enc_lens_out);
/* The two boolean masks handed to the layers by the reference, for bit-exact tests:
 * att_mask[n * (l+c+r)] (encoder.py:637-645) and conv_mask[n * (c + 2*(kernel/2))] (encoder.py:627-633); 0/1 bytes. */
CF_API int cf_plan_masks(const cf_plan* p, uint8_t* att_mask, uint8_t* conv_mask);
/* Per-chunk int32 table, 8 ints per chunk: {utt, j, att_lo, att_hi, conv_lo, conv_hi, out_lo, out_hi}. */
CF_API int cf_plan_chunk_table(const cf_plan* p, int32_t* table);

/* ---- encoder ----------------------------------------------------------------------------------------------------- */
CF_API size_t cf_workspace_bytes(const cf_handle* h, const cf_plan* p);
/* Copy the plan's per-chunk tables into `workspace` once (synchronous).  cf_encode calls with this plan and this workspace then
 * issue no operation that depends on host memory (no table staging), so a steady-state call - e.g. the streaming step of
 * forward_chunk (encoder.py:310-390) once every stream's left context is filled, where the plan no longer changes - can be
 * captured in a CUDA graph and replayed.  No other plan may use the workspace in between; workspace = NULL unpins. */
CF_API int cf_plan_pin(cf_handle* h, cf_plan* p, void* workspace, size_t workspace_bytes, void* stream);
/* Replaces: ChunkFormerEncoder.forward_parallel_chunk / forward_encoder from CMVN to after_norm
 * (encoder.py:615-671; encoder_layer.py:155-248; attention.py:420-505; convolution.py:194-255; subsampling.py:120-175).
 *   feats          device, fp32, flat [rows, feat_dim] (all utterances; see feat_row_offsets of the plan)
 *   att_cache      device, fp32 (L, l, H, 2*d_k) in/out or NULL  (attention.py:459-467)
 *   cnn_cache      device, fp32 (L, d, kernel/2) in/out or NULL  (convolution.py:228-232)
 *   truncated_context_size  rows kept for the next segment's caches (chunkformer_model.py:364-365)
 *   out            device, [n * c, d_model] in out_dtype (CF_F32 | CF_BF16); rows >= enc_len of an utterance are undefined
 *   out_bf16       optional second output (bf16 copy used as the CTC GEMM operand), may be NULL
 */
CF_API int cf_encode(cf_handle* h, const cf_plan* p, const float* feats, void* att_cache, void* cnn_cache,
              int truncated_context_size, void* out, int out_dtype, void* out_bf16, void* workspace,
              size_t workspace_bytes, void* stream);

/* Frame-synchronous streaming for many concurrent streams in ONE pass.  Replaces: ChunkFormerEncoder.forward_chunk
 * (chunkformer/modules/encoder.py:310-390) at right_context_size = 0.  Call before cf_encode: the next cf_encode call then treats
 * its masked-batch plan as `n_streams` utterances of `placeholder_chunks` + 1 chunks each (placeholder_chunks * chunk_size >=
 * max(left_context, 7); the placeholder input frames may be zeros; offsets[s] = -(placeholder_chunks * chunk_size -
 * min(frames already consumed, left_context)) masks the part of the cache that is not filled yet) and its caches as device fp32
 * att_cache (L, n_streams, H, left_context, 2 d_k) / cnn_cache (L, n_streams, d, 7), the layouts forward_chunk takes and
 * returns; both are updated in place.  The rows of the last chunk of every utterance are the step's output: in compact mode
 * (the default where it applies, see cf_encode_output_rows) `out` holds exactly those rows, n_streams x chunk of them; otherwise
 * `out` has the plan's rows and the step's output is rows [placeholder_chunks * chunk, (placeholder_chunks + 1) * chunk) of
 * every utterance. */
/* advance = frames every stream moves forward per step: 0 or the plan's chunk size for right_context_size = 0; with a right
 * context r the plan's chunk is chunk + r (chunk and right context are embedded and attended as ONE chunk, encoder.py:341-347),
 * advance = chunk: the caches then end where the chunk ends (encoder.py:376-385) and the conv module is cut into sub-chunks of
 * `advance` frames with zeros to their right (convolution.py:150-167). */
CF_API int cf_encode_streams(cf_handle* h, int n_streams, int placeholder_chunks, int advance);

/* ---- CTC head ---------------------------------------------------------------------------------------------------- */
CF_API size_t cf_ctc_workspace_bytes(const cf_handle* h, int64_t rows, int enc_dtype);
/* Replaces: CTC.log_softmax + argmax (chunkformer/modules/ctc.py:73-91; chunkformer_model.py:437-438, 526-527).
 * enc: device [rows, d_model], CF_BF16 (the out_bf16 of cf_encode) or CF_F32 (rounded to bf16 by a kernel of this library
 * inside the call).  tokens_out: device int64 [rows]. margin_out: optional device fp32 [rows] = best logit minus runner-up
 * (for tolerance-aware comparisons). logp_out: optional device fp32 [rows, vocab] log-softmax. */
CF_API int cf_ctc_greedy(cf_handle* h, const void* enc, int enc_dtype, int64_t rows, int64_t* tokens_out, float* margin_out,
                  float* logp_out, void* workspace, size_t workspace_bytes, void* stream);

/* Replaces: remove_duplicates_and_blank (chunkformer/utils/model_utils.py:23-32) and the per-frame blank filter of
 * get_output_with_timestamps (:186-196), so that only the kept tokens cross PCIe instead of one id per encoder frame.
 * tokens: device int64 [rows] (cf_ctc_greedy output); utterance s owns rows [seg_start[s], seg_start[s] + seg_len[s]),
 * seg_start ascending (device int64 [n_seg], device int32 [n_seg]).  mode 0: CTC collapse (drop repeats, then blanks);
 * mode 1: drop blanks only.  out_tokens / out_frames: device [rows] capacity, kept tokens and their frame index inside the
 * utterance, in row order; out_offsets: device int64 [n_seg + 1], first output position of every utterance and the total. */
CF_API size_t cf_ctc_compact_workspace_bytes(int64_t rows);
CF_API int cf_ctc_compact(const int64_t* tokens, int64_t rows, const int64_t* seg_start, const int32_t* seg_len, int n_seg,
                   int mode, int64_t blank_id, int64_t* out_tokens, int32_t* out_frames, int64_t* out_offsets,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ---- transducer greedy search (the step after the encoder for chunkformer-rnnt-* models) ------------------------------- */
/* Replaces: optimized_search / batch_greedy_search (chunkformer/transducer/search/greedy_search.py:6-99) with
 * RNNPredictor.forward_step (transducer/predictor.py:190-207, LSTM) and TransducerJoint.forward (transducer/joint.py:69-101,
 * prejoin_linear, joint_mode add, tanh, non-HAT), as called from chunkformer_model.py:440-447 and :528-541. */
typedef struct cf_rnnt cf_rnnt;
typedef struct cf_rnnt_config {
  int32_t vocab;      /* joint output size (blank included) */
  int32_t embed;      /* predictor_conf.embed_size */
  int32_t hidden;     /* predictor_conf.hidden_size */
  int32_t layers;     /* predictor_conf.num_layers (LSTM) */
  int32_t pred_out;   /* predictor_conf.output_size */
  int32_t enc_dim;    /* joint_conf.enc_output_size = encoder output_size */
  int32_t join_dim;   /* joint_conf.join_dim */
  int32_t blank;      /* blank id (0) */
} cf_rnnt_config;
CF_API int cf_rnnt_create(const cf_rnnt_config* cfg, int device, cf_rnnt** out);
CF_API void cf_rnnt_destroy(cf_rnnt* h);
CF_API const char* cf_rnnt_last_error(const cf_rnnt* h);
/* One call per checkpoint tensor under `predictor.` / `joint.` (host fp32, any order), then finalize. */
/* Per-handle option.  "persistent" (default 1): the whole search runs as one persistent cooperative kernel (weight slices
 * resident in shared memory, grid barriers between the phases of an iteration) whenever the model fits one CTA per SM;
 * 0 = one launch per phase (kept for A/B measurement and as the path for models that do not fit). */
CF_API int cf_rnnt_set_option(cf_rnnt* h, const char* name, int value);
CF_API int cf_rnnt_load_tensor(cf_rnnt* h, const char* state_dict_key, const float* host_f32, int ndim, const int64_t* shape);
CF_API int cf_rnnt_finalize_weights(cf_rnnt* h);
CF_API size_t cf_rnnt_workspace_bytes(const cf_rnnt* h, int64_t rows, int n_utt);
/* enc: device fp32 [rows, enc_dim] (encoder output); utterance b owns rows [seg_start[b], seg_start[b] + seg_len[b]) (HOST
 * arrays).  Emits at most n_steps symbols per frame (greedy_search.py:10).  out_tokens / out_frames: device [n_utt, capacity]
 * = symbols in emission order and the frame each was emitted on; out_counts: device int32 [n_utt].  The dense
 * (T, n_steps) grid of optimized_search is (frame, k-th symbol of that frame).  Synchronises the stream (the host polls the
 * number of unfinished utterances); iterations_out (optional, host) = search iterations enqueued.
 * CF_ERR_WORKSPACE if an utterance needs more than `capacity` symbols. */
CF_API int cf_rnnt_greedy(cf_rnnt* h, const float* enc_f32, int64_t rows, const int64_t* seg_start, const int32_t* seg_len,
                   int n_utt, int n_steps, int capacity, int64_t* out_tokens, int32_t* out_frames, int32_t* out_counts,
                   int64_t* iterations_out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- kernel-level entry points (parity tests, profiling) --------------------------------------------------------- */
/* C[M,N] = A[M,K] * B[N,K]^T on the tcgen05 GEMM with a fused epilogue; epi: 0 bf16 out = act(acc+bias), 1 GLU (bf16,
 * N/2 columns, value/gate rows interleaved), 2 fp32 out = resid + rowmask*alpha*(acc+bias), 4 argmax partials
 * [M, 2*ceil(N/256)]; act: 0 none 1 relu 2 silu; variant: -1 = kernel chosen by problem size, 0 = 1-CTA kernel, 1 = 2-CTA
 * pair (cta_group::2) kernel.  All pointers device. */
CF_API int cf_op_gemm(const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K, int epi, int act,
               const float* bias, const float* resid, int64_t ld_resid, float alpha, const int32_t* row_range,
               int rows_per_chunk, void* out, int64_t ldo, float* part_best, float* part_second, int32_t* part_index,
               int variant, void* stream);
/* Residual GEMM with the following LayerNorm(s) fused into its epilogue (a CTA pair per 128-row block exchanges row statistics
 * over distributed shared memory): x_new = resid + rowmask * alpha * (A B^T + bias), N = d_model in {256, 512};
 * mode 1: x_out = x_new, y_out = LN1(x_new) (rows beyond row_limit zeroed); 2: x_out = LN1(x_new), y_out = LN2(x_out);
 * 3: x_out (fp32, nullable) / y_out (bf16, nullable) = LN2(LN1(x_new)).  Replaces every "x = x + f(norm(x))" step of
 * encoder_layer.py:190-246 together with the LayerNorm that follows it.  Default kernel: CTA pair with the normalisation passes on
 * their own warps; mode + 32 selects the cluster-of-four version (cta_group::2 MMAs), mode + 16 the first version (one set of
 * epilogue warps for all passes), mode + 64 the default explicitly: for A/B measurements, results are identical. */
CF_API int cf_op_gemm_ln(const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K, const float* bias,
                  const float* resid, int64_t ld_resid, float alpha, const int32_t* row_range, int rows_per_chunk, int mode,
                  const float* ln1_w, const float* ln1_b, const float* ln2_w, const float* ln2_b, float* x_out, int64_t ldx,
                  void* y_out, int64_t ldy, const int32_t* row_limit, int rows_per_seq, void* stream);
/* Whole feed-forward module in one kernel (positionwise_feed_forward.py:51-60 + encoder_layer.py:190-199 / 236-246):
 * x_new = resid + alpha * (W2 SiLU(W1 y + b1) + b2), then the LayerNorm modes of cf_op_gemm_ln.  The [M, F] hidden activation
 * stays on chip (a CTA pair per 128-row block exchanges SiLU tiles over distributed shared memory).  d in {256, 512},
 * F % 256 == 0; w1 [F, d], w2 [d, F] bf16 row-major (nn.Linear layout). */
CF_API int cf_op_ffn(const void* y_bf16, int64_t ldy_in, const void* w1, const float* b1, const void* w2, const float* b2, int M,
              int d, int F, const float* resid, int64_t ld_resid, float alpha, int mode, const float* ln1_w, const float* ln1_b,
              const float* ln2_w, const float* ln2_b, float* x_out, int64_t ldx, void* y_out, int64_t ldy, void* stream);
/* mode 0: y=LN1(x); 1: x<-LN1(x), y=LN2(x); 2: out=LN2(LN1(x)). */
CF_API int cf_op_layernorm(int mode, int d, const float* x_in, float* x_out, void* y_bf16, const float* w1, const float* b1,
                    const float* w2, const float* b2, int64_t rows, void* stream);
CF_API int cf_op_dwconv(int d, int kernel, const void* g_bf16, void* z_bf16, const float* w, const float* bias,
                 const float* ln_w, const float* ln_b, const int32_t* range, int c, int n_chunks, void* stream);
/* impl 0 = generic CUDA-core kernel, 1 = tcgen05 kernels (128-key-block kernel where it applies, else the ring kernel),
 * 3 = ring kernel.
 * prescaled != 0: the Q+u / Q+v columns already carry (1/sqrt(d_k)) * log2(e), as cf_encode's fused projection writes them. */
CF_API int cf_op_attention(int impl, const void* qkv_bf16, const void* pos_bf16, const int32_t* range, void* ctx_bf16,
                    int n_chunks, int c, int l, int r, int d, int heads, int prescaled, void* stream);

/* conv0 + ReLU + depthwise conv1 of the subsampling front-end (subsampling.py:70-92) for n_chunks chunks read from the flat
 * feature buffer; impl 0 = CUDA-core kernel, 1 = tcgen05 kernel. wpack: device fp32 [d][20] = {w0[9], b0, w1[9], b1};
 * chunk_feat_row / chunk_in_len: host arrays; out: device bf16 [n_chunks * (2c+1) * F2, d]. Synchronises the stream. */
CF_API int cf_op_frontend_conv(int impl, int d, const float* feats, const int64_t* chunk_feat_row, const int32_t* chunk_in_len,
                        int n_chunks, int chunk_size, int feat_dim, const float* wpack, const float* cmvn_mean,
                        const float* cmvn_istd, void* out_bf16, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CHUNKFORMER_B200_H_ */
