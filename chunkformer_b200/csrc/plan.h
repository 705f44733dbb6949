// Host-side masked-batch packer: chunk list + bound tables (pure C++, no CUDA).
// Closed forms of ChunkFormerEncoder.forward_parallel_chunk's packer (encoder.py:538-612, 627-645) and of the
// padded-batch geometry of forward_encoder (encoder.py:220-274) — see SURVEY.md 8(a) row 2 and DESIGN.md.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

struct cf_chunk_entry {
  int32_t utt, j;
  int32_t att_lo, att_hi;    // valid key-window slots [lo, hi), slot q <-> frame c*j - l + q
  int32_t conv_lo, conv_hi;  // valid conv-window slots [lo, hi), slot q <-> frame c*j - lorder + q
  int32_t out_lo, out_hi;    // rows of the chunk whose conv-module output is kept (others zeroed)
};

struct cf_plan {
  int mode = 0;  // 0 = masked batch (forward_parallel_chunk), 1 = padded batch (forward_encoder)
  int c = 0, l = 0, r = 0, kernel = 15, lorder = 7, B = 0;
  int n = 0;                     // total chunks
  int in_rows = 0;               // input frames per chunk: 8(c-1)+15
  int padded_T = 0;              // mode 1: T of the padded batch
  int rows_per_seq = 0;          // mode 1: chunks_per_seq * c
  std::vector<int32_t> lens, offsets, n_chunks, pad, valid, enc_lens;
  std::vector<int64_t> feat_row_offsets;
  std::vector<cf_chunk_entry> chunks;
  std::vector<int64_t> chunk_feat_row;  // first input row of each chunk in the flat feature buffer
  std::vector<int32_t> chunk_in_len;    // input rows present (zero padded beyond)
  std::vector<int32_t> seq_valid_rows;  // mode 1: per sequence calc_length(len) (LayerNorm zeroing limit)
  const void* resident_ws = nullptr;    // cf_plan_pin: the workspace whose table region already holds this plan's tables
};

namespace cfplan {
int calc_length(int T);
bool build_masked(cf_plan* p, int c, int l, int r, int kernel, int B, const int32_t* lens, const int32_t* offsets,
                  const int64_t* feat_row_offsets, std::string* err);
bool build_padded(cf_plan* p, int c, int l, int r, int kernel, int B, int T, const int32_t* lens, std::string* err);
}  // namespace cfplan
