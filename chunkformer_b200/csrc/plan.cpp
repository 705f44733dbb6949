#include "plan.h"

#include <algorithm>
#include <cmath>

namespace cfplan {

static inline int floordiv(int a, int b) {  // floor division, b > 0
  int q = a / b, r = a % b;
  return (r != 0 && r < 0) ? q - 1 : q;
}

// subsampling.py:270-288: three times floor((L + 0 - 3) / 2) + 1, evaluated in float by the reference.
int calc_length(int T) {
  double L = double(T);
  for (int i = 0; i < 3; ++i) L = std::floor((L - 3.0) / 2.0) + 1.0;
  return int(L);
}

static bool check_common(int c, int l, int r, int kernel, int B, std::string* err) {
  if (c <= 0 || l < 0 || r < 0 || B <= 0 || kernel <= 0 || (kernel % 2) == 0) {
    if (err) *err = "plan: need chunk_size > 0, contexts >= 0, odd conv kernel, B > 0";
    return false;
  }
  return true;
}

bool build_masked(cf_plan* p, int c, int l, int r, int kernel, int B, const int32_t* lens, const int32_t* offsets,
                  const int64_t* feat_row_offsets, std::string* err) {
  if (!check_common(c, l, r, kernel, B, err)) return false;
  const int lo = kernel / 2;
  const int size = (c - 1) * 8 + 15, step = 8 * c, W = l + c + r;
  p->mode = 0; p->c = c; p->l = l; p->r = r; p->kernel = kernel; p->lorder = lo; p->B = B;
  p->in_rows = size;
  p->lens.assign(lens, lens + B);
  p->offsets.assign(B, 0);
  if (offsets) p->offsets.assign(offsets, offsets + B);
  p->feat_row_offsets.resize(B);
  int64_t acc = 0;
  for (int u = 0; u < B; ++u) {
    if (lens[u] <= 0) { if (err) *err = "plan: utterance with no frames"; return false; }
    p->feat_row_offsets[u] = feat_row_offsets ? feat_row_offsets[u] : acc;
    acc += lens[u];
  }
  p->n_chunks.resize(B); p->pad.resize(B); p->valid.resize(B); p->enc_lens.resize(B);
  p->chunks.clear(); p->chunk_feat_row.clear(); p->chunk_in_len.clear();
  for (int u = 0; u < B; ++u) {
    const int T = lens[u];
    const int pad = T >= size ? (step - ((T - size) % step)) % step : size - T;   // encoder.py:557-560
    const int nck = (T + pad - size) / step + 1;                                    // encoder.py:562
    const int M = 1 + floordiv(T - 15, 8);                                          // encoder.py:567 (may be <= 0)
    const int o = p->offsets[u];
    p->pad[u] = pad; p->n_chunks[u] = nck; p->valid[u] = M; p->enc_lens[u] = calc_length(T);
    for (int j = 0; j < nck; ++j) {
      cf_chunk_entry e;
      e.utt = u; e.j = j;
      // key slot q <-> f = c*j - l + q valid  <=>  -o <= f < M
      e.att_lo = std::min(W, std::max(0, l - c * j - o));
      e.att_hi = std::max(e.att_lo, std::min(W, M - c * j + l));
      // conv slot q <-> f = c*j - lo + q valid  <=>  -o <= f < min(M, c*(j+1) + r)
      const int cw = c + 2 * lo;
      e.conv_lo = std::min(cw, std::max(0, lo - c * j - o));
      e.conv_hi = std::max(e.conv_lo, std::min(cw, std::min(M, c * (j + 1) + r) - c * j + lo));
      // output row i kept <=> centre slot lo + i valid (convolution.py:253)
      e.out_lo = std::min(c, std::max(0, e.conv_lo - lo));
      e.out_hi = std::max(e.out_lo, std::min(c, e.conv_hi - lo));
      p->chunks.push_back(e);
      p->chunk_feat_row.push_back(p->feat_row_offsets[u] + int64_t(step) * j);
      p->chunk_in_len.push_back(std::min(size, T - step * j));
    }
  }
  p->n = int(p->chunks.size());
  return true;
}

bool build_padded(cf_plan* p, int c, int l, int r, int kernel, int B, int T, const int32_t* lens, std::string* err) {
  if (!check_common(c, l, r, kernel, B, err)) return false;
  if (T < 15) { if (err) *err = "plan: padded batch needs T >= 15 (the reference's Conv2d fails below)"; return false; }
  const int lo = kernel / 2;
  const int W = l + c + r;
  const int Tp = (T - 15) / 8 + 1;
  const int nck = (Tp + c - 1) / c;
  p->mode = 1; p->c = c; p->l = l; p->r = r; p->kernel = kernel; p->lorder = lo; p->B = B;
  p->in_rows = (c - 1) * 8 + 15; p->padded_T = T; p->rows_per_seq = nck * c;
  p->lens.assign(lens, lens + B);
  p->offsets.assign(B, 0);
  p->n_chunks.assign(B, nck); p->pad.assign(B, 0); p->valid.resize(B); p->enc_lens.resize(B);
  p->feat_row_offsets.resize(B); p->seq_valid_rows.resize(B);
  p->chunks.clear(); p->chunk_feat_row.clear(); p->chunk_in_len.clear();
  for (int u = 0; u < B; ++u) {
    if (lens[u] > T) { if (err) *err = "plan: len > T"; return false; }
    const int M = std::max(0, calc_length(lens[u]));
    p->valid[u] = M; p->enc_lens[u] = calc_length(lens[u]); p->seq_valid_rows[u] = M;
    p->feat_row_offsets[u] = int64_t(u) * T;
    for (int j = 0; j < nck; ++j) {
      cf_chunk_entry e;
      e.utt = u; e.j = j;
      e.att_lo = std::min(W, std::max(0, l - c * j));                        // f >= 0
      e.att_hi = std::max(e.att_lo, std::min(W, M - c * j + l));             // f < len'  (attention.py:365-383)
      e.conv_lo = std::min(c + lo, std::max(0, lo - c * j));                 // left halo real, zeros before frame 0
      e.conv_hi = std::max(e.conv_lo, std::min(c + lo, Tp - c * j + lo));    // right halo zero (convolution.py:150-167)
      e.out_lo = 0;
      e.out_hi = std::max(0, std::min(c, M - c * j));                        // rows >= len' zeroed (convolution.py:189-191)
      p->chunks.push_back(e);
      p->chunk_feat_row.push_back(int64_t(u) * T + int64_t(8 * c) * j);
      p->chunk_in_len.push_back(std::min(p->in_rows, T - 8 * c * j));
    }
  }
  p->n = int(p->chunks.size());
  return true;
}

}  // namespace cfplan
