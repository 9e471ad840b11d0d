// Host-side logic shared by the CUDA engine (engine.cu) and the CPU-only host
// emulator (emu.cpp): equality-table construction and the row tiling.
#pragma once
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "scan_core.cuh"

namespace sb {

#ifndef SB_THREADS
#define SB_THREADS 128
#endif
constexpr int kScanThreads = SB_THREADS;  // rows (threads) per block; each warp has its own 32-row TMA box
constexpr uint32_t kMaxRowBytes = 16384;
constexpr uint32_t kRowAlign = 128;  // rows start on 128-byte lines
constexpr int kMaxWords = 128;       // patterns up to 128 * 32 = 4096 characters
constexpr int kMaxScanWords = 32;    // the row-tiled scan kernels (one thread per text row) stop here: longer
                                     // patterns run on the warp-per-window kernels only (Engine::search)
constexpr uint64_t kCoverStride = 8192;  // end positions per window of their full scan

// Supported word counts (kernel template instantiations).
inline int round_words(int w) {
  const int opts[] = {1, 2, 3, 4, 6, 8, 16, 32, 64, 128};
  for (int o : opts)
    if (w <= o) return o;
  return -1;
}

struct ProfileParams {
  uint32_t nrows, sh0, msk0;
};

inline bool profile_params(int profile, ProfileParams& p) {
  switch (profile) {
    case kDna: p = {4, 1, 0x03030303u}; return true;
    case kIupac: p = {32, 0, 0x1F1F1F1Fu}; return true;
    case kAscii: p = {256, 0, 0xFFFFFFFFu}; return true;
    default: return false;
  }
}

// Equality table of one query: eq[row][w], bit (pad + j) set iff pattern[j]
// matches a text byte that selects `row`; the low `pad` wildcard bits are all
// ones (cf. reference src/pattern_tiling/tqueries.rs:90-114 for the
// per-pattern peq build; the wildcard padding is ours).
inline void build_eq_table(int profile, const uint8_t* p, int m, int W, uint32_t nrows, uint32_t* tab) {
  const int pad = 32 * W - m;
  memset(tab, 0, (size_t)nrows * W * sizeof(uint32_t));
  for (uint32_t row = 0; row < nrows; row++) {
    uint32_t* e = tab + (size_t)row * W;
    for (int b = 0; b < pad; b++) e[b >> 5] |= 1u << (b & 31);
    for (int j = 0; j < m; j++) {
      bool match;
      switch (profile) {
        case kDna: match = row_matches<kDna>(p[j], (int)row); break;
        case kIupac: match = row_matches<kIupac>(p[j], (int)row); break;
        default: match = row_matches<kAscii>(p[j], (int)row); break;
      }
      if (match) {
        const int b = pad + j;
        e[b >> 5] |= 1u << (b & 31);
      }
    }
  }
}

// Row tiling: rows of `ltot` bytes, one thread per row, kScanThreads rows per
// block, one block per (tile, query).  ltot is as long as possible (the
// per-row warm-up costs nwarm*128 bytes) subject to filling the GPU with a
// whole number of waves when the grid is small.  bpw = resident blocks per wave.
// nq = blocks per tile of ONE launch (the queries of a direction; half of them, rounded up, when two
// patterns share a thread): a launch then fills whole waves.
inline ScanGeom choose_geom(uint64_t n, int m, int k, uint32_t nq, int bpw) {
  ScanGeom g;
  g.nwarm = (uint32_t)((m + k + kStageBytes - 1) / kStageBytes);
  if (g.nwarm == 0) g.nwarm = 1;
  const uint64_t min_ltot = ((uint64_t)g.nwarm * kStageBytes + kRowAlign - 1) / kRowAlign * kRowAlign;
  const uint64_t per_tile_max = (uint64_t)kScanThreads * kMaxRowBytes;
  uint64_t tiles = std::max<uint64_t>(1, (n + per_tile_max - 1) / per_tile_max);
  const uint64_t fill = ((uint64_t)bpw + nq - 1) / nq;  // tiles needed to give every SM slot a block
  tiles = std::max(tiles, fill);
  uint64_t blocks = tiles * nq;
  if (blocks < 16ull * bpw) {  // few waves: whole waves (never more blocks than they hold: tiles round down)
    blocks = (blocks + bpw - 1) / bpw * bpw;
    tiles = std::max<uint64_t>(tiles, blocks / nq);
  }
  uint64_t ltot = (n + tiles * kScanThreads - 1) / (tiles * kScanThreads);
  ltot = (ltot + kRowAlign - 1) / kRowAlign * kRowAlign;
  ltot = std::max(ltot, min_ltot);
  ltot = std::min<uint64_t>(ltot, std::max<uint64_t>(kMaxRowBytes, min_ltot));
  g.ltot = (uint32_t)ltot;
  g.rows = (uint32_t)((n + ltot - 1) / ltot);
  g.nstage = g.ltot / kStageBytes;
  return g;
}

// ---------------------------------------------------------------------------
// Exact piece prefilter: layout and tables (see scan_core.cuh).
constexpr int kMaxPieces = 256;  // >= kMaxFilterWords * (32 / (1 + kFilterDelay))

struct FilterPiece {
  int off;   // first pattern position of the piece
  int len;   // characters
  int word;  // automaton word
  int bit;   // first automaton bit; bits [bit, bit+len) then kFilterDelay delay bits
};

struct FilterPlan {
  bool enabled = false;
  int WF = 0;        // automaton words (1, 2, 4 or 8)
  int L = 0;         // shortest piece
  int npieces = 0;   // k + 1
  FilterPiece piece[kMaxPieces];
  double rate = 0;   // expected piece occurrences per text position (uniform ACGT text), max over queries
  double cost = 0;   // modelled cost relative to the full scan (1.0)
  uint32_t finit[kMaxFilterWords] = {};
  uint32_t fdelay[kMaxFilterWords] = {};
};

// Probability that a uniformly random ACGT character matches pattern byte c.
inline double match_prob(int profile, uint8_t c) {
  if (profile == kIupac) {
    const uint8_t code = iupac_code(c);
    if (code == 255) return 1.0;
    int bits = 0;
    for (int i = 0; i < 4; i++) bits += (code >> i) & 1;
    return bits / 4.0;
  }
  return 0.25;  // Dna classes; Ascii: assumed, the run-time hit counter guards the assumption
}

// Lays k+1 disjoint pieces of the pattern out over WF automaton words: the pattern is cut
// into k+1 shares (lengths differ by at most 1), the shares are spread evenly over the words,
// and every piece is the prefix of its share that fits its word (32 / pieces_in_word - 3 bits).
inline bool layout_pieces(FilterPlan& f, int m, int k, int WF) {
  const int np = k + 1;
  if (np > m || np > kMaxPieces) return false;
  f.WF = WF;
  f.npieces = np;
  f.L = 1 << 30;
  for (int w = 0; w < kMaxFilterWords; w++) f.finit[w] = f.fdelay[w] = 0;
  int p = 0, off = 0;
  for (int w = 0; w < WF; w++) {
    const int cnt = np / WF + (w < np % WF ? 1 : 0);  // pieces in this word
    if (cnt == 0) continue;
    const int room = 32 / cnt - kFilterDelay;
    if (room < 1) return false;
    int bit = 0;
    for (int c = 0; c < cnt; c++, p++) {
      const int share = m / np + (p < m % np ? 1 : 0);
      FilterPiece& pc = f.piece[p];
      pc.off = off;
      pc.len = share < room ? share : room;
      pc.word = w;
      pc.bit = bit;
      f.finit[w] |= 1u << bit;
      f.fdelay[w] |= 1u << (bit + pc.len - 1);  // the piece's last bit ...
      for (int d = 0; d < kFilterDelay; d++) f.fdelay[w] |= 1u << (bit + pc.len + d);  // ... and its delay line
      bit += pc.len + kFilterDelay;
      off += share;
      if (pc.len < f.L) f.L = pc.len;
    }
  }
  return true;
}

// Chooses the automaton width with the lowest modelled cost; enabled when that is below
// `max_cost` times the cost of the full scan.  Model (per text character, in units of one
// scan word-step): scan = W; prefilter = 0.35 * WF + 0.05 (measured ratio of the two
// kernels' instruction streams); re-scan = occurrences * window * W * 3 (one thread per hit
// is about 3x less efficient than the streaming scan).
inline FilterPlan plan_filter(int profile, const uint8_t* const* queries, size_t nq, int m, int k,
                              double max_cost = 0.85) {
  FilterPlan best;
  if (k < 0 || k + 1 > m || nq == 0) return best;
  const int W = (m + 31) / 32;
  const double window = 2.0 * (m + k) + 4.0;
  const int wopts[] = {1, 2, 4, 8};
  bool have = false;
  for (int WF : wopts) {
    FilterPlan f;
    if (!layout_pieces(f, m, k, WF)) continue;
    double worst = 0;
    for (size_t q = 0; q < nq; q++) {
      double rate = 0;
      for (int p = 0; p < f.npieces; p++) {
        double pr = 1;
        for (int j = 0; j < f.piece[p].len; j++) pr *= match_prob(profile, queries[q][f.piece[p].off + j]);
        rate += pr;
      }
      if (rate > worst) worst = rate;
    }
    f.rate = worst;
    f.cost = (0.35 * WF + 0.05 + worst * window * W * 3.0) / (double)W;
    if (!have || f.cost < best.cost) best = f, have = true;
  }
  best.enabled = have && best.cost <= max_cost;
  return best;
}

// Automaton masks of the text class `row` for one query.  reversed = the query is the
// reversed partner (scanned over the reversed text by the reference) but the automaton
// scans the text FORWARD (strand-fused prefilter): every piece is matched back to front.
inline void class_masks(int profile, const FilterPlan& f, const uint8_t* pat, bool reversed, int row,
                        uint32_t* out /*[f.WF]*/) {
  for (int w = 0; w < f.WF; w++) out[w] = 0;
  for (int p = 0; p < f.npieces; p++) {
    const FilterPiece& pc = f.piece[p];
    for (int j = 0; j < pc.len; j++) {
      const uint8_t ch = pat[pc.off + (reversed ? pc.len - 1 - j : j)];
      bool match;
      switch (profile) {
        case kDna: match = row_matches<kDna>(ch, row); break;
        case kIupac: match = row_matches<kIupac>(ch, row); break;
        default: match = row_matches<kAscii>(ch, row); break;
      }
      if (match) out[pc.word] |= 1u << (pc.bit + j);
    }
    for (int d = 0; d < kFilterDelay; d++) out[pc.word] |= 1u << (pc.bit + pc.len + d);
  }
}

// Filter automaton masks: tab[byte][WF_total].  pat_rev != nullptr builds the strand-fused
// automaton: words [0, f.WF) for `pat`, words [f.WF, 2 f.WF) for the reversed partner.
inline void build_filter_table(int profile, const FilterPlan& f, const uint8_t* pat, uint32_t* tab,
                               const uint8_t* pat_rev = nullptr) {
  ProfileParams pp;
  profile_params(profile, pp);
  const int WT = pat_rev ? 2 * f.WF : f.WF;
  for (int byte = 0; byte < 256; byte++) {
    uint32_t* e = tab + (size_t)byte * WT;
    const int row = (int)(((uint32_t)byte >> pp.sh0) & (pp.msk0 & 0xFFu));
    class_masks(profile, f, pat, false, row, e);
    if (pat_rev) class_masks(profile, f, pat_rev, true, row, e + f.WF);
  }
}

// Pair table of the two-characters-per-step automaton (Dna profile only):
// tab[word][first | second << 2] = {A, B} (one 128-byte sub-table per automaton word), see
// filter16_pair / load_pair in scan_core.cuh.
inline void build_pair_table(const FilterPlan& f, const uint8_t* pat, uint32_t* tab,
                             const uint8_t* pat_rev = nullptr) {
  const int WT = pat_rev ? 2 * f.WF : f.WF;
  uint32_t cls[4][2 * kMaxFilterWords] = {};
  uint32_t init[2 * kMaxFilterWords] = {};
  for (int w = 0; w < f.WF; w++) init[w] = f.finit[w], init[f.WF + w] = f.finit[w];
  for (int c = 0; c < 4; c++) {
    class_masks(kDna, f, pat, false, c, cls[c]);
    if (pat_rev) class_masks(kDna, f, pat_rev, true, c, cls[c] + f.WF);
  }
  for (int c1 = 0; c1 < 4; c1++)
    for (int c0 = 0; c0 < 4; c0++) {
      for (int w = 0; w < WT; w++) {
        uint32_t* e = tab + (size_t)w * 32 + (size_t)(c0 | (c1 << 2)) * 2;
        e[0] = (cls[c0][w] << 1) & cls[c1][w];
        e[1] = (((init[w] & cls[c0][w]) << 1) | init[w]) & cls[c1][w];
      }
    }
}
constexpr int kPairTableWords = 16 * 2;  // per automaton word

// ---------------------------------------------------------------------------
// q-gram bitmap prefilter (scan_core.cuh: qgram16 / qgram_confirm): plan and tables.
struct QgramPlan {
  bool enabled = false;
  int q = 0;        // characters per q-gram: bitmap of 4^q bits
  int s = 0;        // sampling distance in characters: 4, 8 or 16
  int npieces = 0;  // k + 1 shares of the pattern
  int off[kMaxPieces];
  int len[kMaxPieces];
  double rate = 0;  // expected hit chunks per text character on uniform ACGT text
  size_t table_words() const { return ((size_t)1 << (2 * q)) / 32; }
};

// Shares of m / (k+1) characters (lengths differ by at most 1).  Q = min(8, shortest share - 3) so
// that the sampling distance can be one text word; used from Q >= 6 (a 4^6-bit table is hit by
// (k+1)/4096 of all positions, below that the exact confirmation would dominate).
inline QgramPlan plan_qgram(int m, int k, int strands, int min_q = 6) {
  QgramPlan f;
  const int np = k + 1;
  if (k < 0 || np > kMaxPieces || np > m) return f;
  const int lmin = m / np;
  const int q = std::min(8, lmin - 3);
  if (q < min_q) return f;
  f.q = q;
  f.s = lmin >= q + 15 ? 16 : (lmin >= q + 7 ? 8 : 4);
  f.npieces = np;
  int off = 0;
  for (int p = 0; p < np; p++) {
    const int share = m / np + (p < m % np ? 1 : 0);
    f.off[p] = off;
    f.len[p] = share;
    off += share;
  }
  f.rate = (double)np * strands / (double)((size_t)1 << (2 * q));
  f.enabled = f.rate <= 4e-3;
  return f;
}

// Forward-orientation class codes of share p of one query: the query bytes as the engine scans
// them (the reversed partner = complement(pattern), scanned over the reversed text, occurs in the
// forward text back to front).
inline void qgram_piece_classes(const QgramPlan& f, int p, const uint8_t* query, bool reversed, uint8_t* out) {
  for (int j = 0; j < f.len[p]; j++) {
    const uint8_t ch = query[f.off[p] + (reversed ? f.len[p] - 1 - j : j)];
    out[j] = (uint8_t)((ch >> 1) & 3);
  }
}

// Sets, for every share of the query, the bits of its Q-grams at offsets 0 .. S-1.
inline void add_qgram_entries(const QgramPlan& f, const uint8_t* query, bool reversed, uint32_t* bitmap) {
  std::vector<uint8_t> cls;
  for (int p = 0; p < f.npieces; p++) {
    cls.resize(f.len[p]);
    qgram_piece_classes(f, p, query, reversed, cls.data());
    for (int o = 0; o < f.s && o + f.q <= f.len[p]; o++) {
      uint32_t idx = 0;
      for (int j = 0; j < f.q; j++) idx |= (uint32_t)cls[o + j] << (2 * j);
      bitmap[idx >> 5] |= 1u << (idx & 31);
    }
  }
}

// Piece records for refine_hit (scan_core.cuh): the first min(len, 16) characters of a piece in forward
// orientation as 2-bit classes, the first start position a hit allows relative to its chunk, and where
// the piece sits in the pattern.
inline void piece_conf(const uint8_t* query, bool reversed, int off, int len, int rel0, uint32_t* out) {
  const int l = std::min(len, 16);
  uint32_t code = 0;
  for (int j = 0; j < l; j++) {
    const uint8_t ch = query[off + (reversed ? len - 1 - j : j)];
    code |= (uint32_t)((ch >> 1) & 3) << (2 * j);
  }
  out[0] = code;
  out[1] = l == 16 ? 0xFFFFFFFFu : ((1u << (2 * l)) - 1u);
  out[2] = (uint32_t)rel0;
  out[3] = (uint32_t)off;
  out[4] = (uint32_t)len;
}

inline void build_qgram_confirm(const QgramPlan& f, const uint8_t* query, bool reversed, uint32_t* out) {
  for (int p = 0; p < f.npieces; p++) piece_conf(query, reversed, f.off[p], f.len[p], 1 - f.q, out + p * kConfWords);
}

// Piece automaton: the hit chunk holds the last character of a piece in scan direction of the KERNEL
// (rtl_kernel: the right-to-left pass of reversed queries, whose hit chunk holds the first character
// of the piece in forward orientation).
inline void build_filter_confirm(const FilterPlan& f, const uint8_t* query, bool reversed, bool rtl_kernel,
                                 uint32_t* out) {
  for (int p = 0; p < f.npieces; p++)
    piece_conf(query, reversed, f.piece[p].off, f.piece[p].len, rtl_kernel ? 0 : 1 - f.piece[p].len,
               out + p * kConfWords);
}

inline size_t padded_alloc(uint64_t n) {
  // room for one extra row of any tiling plus alignment slack
  return (size_t)((n + 2ull * kMaxRowBytes + 255ull) & ~255ull);
}

}  // namespace sb
