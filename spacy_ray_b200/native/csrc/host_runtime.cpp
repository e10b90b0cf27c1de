// Host-side native runtime: featurisation + batch collation.
//
// The reference has no native code; what is native upstream is
// spaCy's Cython Doc/lexeme machinery that produces the attribute arrays the
// tok2vec consumes.  This file is the equivalent for our pipeline:
//   * srb_featurize : token strings -> (n, 4) uint64 NORM/PREFIX/SUFFIX/SHAPE ids
//   * srb_collate   : gather docs of a batch out of a corpus-wide attribute store
//                     into the padded-ragged staging buffer (pinned host memory)
//                     that is then copied H2D once per step.
// Plain C ABI (loaded with ctypes), no Python headers needed.
#include <cstdint>
#include <cstring>
#include <vector>
#include <string>

namespace {
constexpr uint64_t kFnvOffset = 0xCBF29CE484222325ull;
constexpr uint64_t kFnvPrime = 0x100000001B3ull;

inline uint64_t fmix64(uint64_t h) {
  h ^= h >> 33; h *= 0xFF51AFD7ED558CCDull;
  h ^= h >> 33; h *= 0xC4CEB9FE1A85EC53ull;
  h ^= h >> 33;
  return h;
}
inline uint64_t hash_bytes(const char* p, int64_t n) {
  uint64_t h = kFnvOffset;
  for (int64_t i = 0; i < n; ++i) h = (h ^ (uint8_t)p[i]) * kFnvPrime;
  h = fmix64(h);
  return h ? h : 1;
}
inline bool is_upper(char c) { return c >= 'A' && c <= 'Z'; }
inline bool is_lower(char c) { return c >= 'a' && c <= 'z'; }
inline bool is_digit(char c) { return c >= '0' && c <= '9'; }
}  // namespace

extern "C" {

// buf: concatenated UTF-8 bytes of all words; offsets: n+1 byte offsets.
// out: n*4 uint64.  non_ascii: n bytes, set to 1 for words containing bytes >= 0x80
// (the caller recomputes those rows in Python, where unicode case/shape rules live).
void srb_featurize(const char* buf, const int64_t* offsets, int64_t n, uint64_t* out,
                   uint8_t* non_ascii) {
  char tmp[128];
  for (int64_t i = 0; i < n; ++i) {
    const char* w = buf + offsets[i];
    int64_t len = offsets[i + 1] - offsets[i];
    bool ascii = true;
    for (int64_t k = 0; k < len; ++k) if ((uint8_t)w[k] >= 0x80) { ascii = false; break; }
    non_ascii[i] = ascii ? 0 : 1;
    if (!ascii) { out[4 * i] = out[4 * i + 1] = out[4 * i + 2] = out[4 * i + 3] = 0; continue; }
    // NORM = lowercase
    if (len <= (int64_t)sizeof(tmp)) {
      for (int64_t k = 0; k < len; ++k) tmp[k] = is_upper(w[k]) ? char(w[k] + 32) : w[k];
      out[4 * i + 0] = hash_bytes(tmp, len);
    } else {
      std::string low(w, (size_t)len);
      for (auto& c : low) if (is_upper(c)) c = char(c + 32);
      out[4 * i + 0] = hash_bytes(low.data(), len);
    }
    // PREFIX = first char, SUFFIX = last three
    out[4 * i + 1] = hash_bytes(w, len < 1 ? len : 1);
    int64_t s3 = len < 3 ? len : 3;
    out[4 * i + 2] = hash_bytes(w + (len - s3), s3);
    // SHAPE
    if (len >= 100) {
      out[4 * i + 3] = hash_bytes("LONG", 4);
    } else {
      char shape[100];
      int64_t m = 0;
      char last = 0;
      int run = 0;
      for (int64_t k = 0; k < len; ++k) {
        char c = w[k];
        char cls = (is_upper(c) ? 'X' : is_lower(c) ? 'x' : is_digit(c) ? 'd' : c);
        if (cls == last && k > 0) { run += 1; } else { run = 0; last = cls; }
        if (run < 4) shape[m++] = cls;
      }
      out[4 * i + 3] = hash_bytes(shape, m);
    }
  }
}

// store: (N_total, n_attr) int64 attribute rows of the whole corpus, doc d at
// rows [doc_off[d], doc_off[d+1]).  Writes the padded layout for `ids[0..B)`:
// row 0 pad, doc, pad, doc, pad ...; rows up to `cap_rows` zero.  Returns rows used,
// or -1 if the staging buffer is too small.
int64_t srb_collate(const int64_t* store, const int64_t* doc_off, int64_t n_attr,
                    const int64_t* ids, int64_t B, int64_t* out_attrs, float* out_mask,
                    int32_t* out_starts, int32_t* out_lens, int64_t cap_rows) {
  int64_t need = 1;
  for (int64_t d = 0; d < B; ++d) need += (doc_off[ids[d] + 1] - doc_off[ids[d]]) + 1;
  if (need > cap_rows) return -1;
  std::memset(out_attrs, 0, sizeof(int64_t) * (size_t)(cap_rows * n_attr));
  std::memset(out_mask, 0, sizeof(float) * (size_t)cap_rows);
  int64_t row = 1;
  for (int64_t d = 0; d < B; ++d) {
    int64_t a = doc_off[ids[d]], n = doc_off[ids[d] + 1] - a;
    out_starts[d] = (int32_t)row;
    out_lens[d] = (int32_t)n;
    std::memcpy(out_attrs + row * n_attr, store + a * n_attr, sizeof(int64_t) * (size_t)(n * n_attr));
    for (int64_t k = 0; k < n; ++k) out_mask[row + k] = 1.0f;
    row += n + 1;
  }
  return need;
}

// Same gather for per-token int32 gold arrays (tags / BILUO actions ...): writes the
// *unpadded* doc-order concatenation plus per-doc offsets.
int64_t srb_collate_gold(const int32_t* store, const int64_t* doc_off, const int64_t* ids,
                         int64_t B, int32_t* out, int32_t* out_off, int64_t cap) {
  int64_t pos = 0;
  for (int64_t d = 0; d < B; ++d) {
    int64_t a = doc_off[ids[d]], n = doc_off[ids[d] + 1] - a;
    if (pos + n > cap) return -1;
    out_off[d] = (int32_t)pos;
    std::memcpy(out + pos, store + a, sizeof(int32_t) * (size_t)n);
    pos += n;
  }
  return pos;
}

// HashEmbed-backward grouping, done on the collate thread so the training step needs no
// device-side sort.  `gid` is the corpus-wide (n_tokens, n_cols) array of dense per-column
// vocabulary indices (built once with the ExampleStore).  For the docs `ids` of one batch this
// writes, per column c, out[c*rb ...]: the padded-layout row of every token of the batch, ordered
// so that equal ids are adjacent (counting sort by vocabulary index - two linear passes, the
// histograms stay cache resident).  The tail [n_tokens, rb) is filled with row 0 (a pad row:
// mask 0, skipped by the kernel).  `hist` is caller scratch of sum(n_groups) + n_cols ints.
int64_t srb_group_rows(const int32_t* gid, const int64_t* doc_off, const int64_t* ids, int64_t B, int32_t n_cols,
                       const int32_t* n_groups, int64_t rb, int32_t* out, int32_t* hist) {
  if (n_cols > 8) return -1;
  int32_t* h[8];
  int64_t pos = 0;
  for (int32_t c = 0; c < n_cols; ++c) { h[c] = hist + pos; pos += n_groups[c] + 1; }
  std::memset(hist, 0, sizeof(int32_t) * (size_t)pos);
  int64_t n_tok = 0;
  for (int64_t d = 0; d < B; ++d) {
    const int64_t a = doc_off[ids[d]], e = doc_off[ids[d] + 1];
    for (int64_t t = a; t < e; ++t)
      for (int32_t c = 0; c < n_cols; ++c) ++h[c][gid[t * n_cols + c] + 1];
    n_tok += e - a;
  }
  if (n_tok + B + 1 > rb) return -1;
  for (int32_t c = 0; c < n_cols; ++c)
    for (int32_t g = 0; g < n_groups[c]; ++g) h[c][g + 1] += h[c][g];      // exclusive starts in h[c][g]
  int64_t row = 1;
  for (int64_t d = 0; d < B; ++d) {
    const int64_t a = doc_off[ids[d]], e = doc_off[ids[d] + 1];
    for (int64_t t = a; t < e; ++t, ++row)
      for (int32_t c = 0; c < n_cols; ++c) out[(size_t)c * rb + h[c][gid[t * n_cols + c]]++] = (int32_t)row;
    ++row;                                                                  // the pad row between docs
  }
  for (int32_t c = 0; c < n_cols; ++c)
    std::memset(out + (size_t)c * rb + n_tok, 0, sizeof(int32_t) * (size_t)(rb - n_tok));
  return n_tok;
}

int srb_abi_version() { return 2; }
}
