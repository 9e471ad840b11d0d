// CPU-only emulation of the GPU pipeline for the `-m "not gpu"` tests.
//
// This is NOT a CPU fallback and is not part of libsassy_b200.so: it is a test
// harness (libsassy_b200_emu.so) that executes the very same per-thread code
// the CUDA kernels run (scan_core.cuh: process16 / select_candidate /
// trace_one) and the same host logic (host_logic.h: equality tables, row
// tiling), one "thread" after the other, so that tiling, warm-up, restart,
// candidate, minima and traceback logic can be checked against the oracle on
// a machine without a GPU.  What it cannot check is the TMA/mbarrier staging.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

#include "dna_pack.h"
#include "host_logic.h"

using namespace sb;

namespace {

template <int W, bool REV>
void scan_rows(const ScanArgs& a, const uint32_t* eq_q, uint32_t qs) {
  EqTab eqt;
  eqt.p = eq_q;
  eqt.saddr = 0;
  eqt.rowbytes = (uint32_t)W * 4u;
  const uint32_t total = a.g.nwarm + a.g.nstage;
  const uint64_t tiles = ((uint64_t)a.g.rows + kScanThreads - 1) / kScanThreads;
  for (uint64_t row = 0; row < tiles * kScanThreads; row++) {  // includes the idle threads of the last tile
    if (a.tile_list && !a.dense[row / kScanThreads]) continue;  // regional fallback: the listed tiles only
    Lane<W> s;
    lane_reset<W>(s, a.m);
    int prev_score = a.m;
    for (uint32_t it = 0; it < total; it++) {
      int64_t r;
      uint32_t col;
      bool own;
      stage_coord<REV>(a.g, it, (int64_t)row, r, col, own);
      const uint64_t stage_idx = (uint64_t)(r * (int64_t)a.g.ltot + (int64_t)col);
      const bool special = stage_is_special(a, stage_idx);
      const bool valid = r >= 0 && r < (int64_t)a.g.rows;
      for (int cc = 0; cc < kStageBytes / 16; cc++) {
        const int c = REV ? (kStageBytes / 16 - 1 - cc) : cc;
        uint32_t x[4] = {0, 0, 0, 0};
        if (valid) memcpy(x, a.text + stage_idx + 16u * c, 16);
        if (special)
          process16<W, REV, true>(s, prev_score, x, stage_idx + 16u * c, a, eqt, qs, own);
        else
          process16<W, REV, false>(s, prev_score, x, stage_idx + 16u * c, a, eqt, qs, own);
      }
    }
  }
}

// The body of scan2_kernel: two one-word patterns per "thread" through the raw-byte pair table.
template <bool REV>
void scan2_rows(const ScanArgs& a, const uint32_t* eq_all /*[queries][nrows]*/, uint32_t nqueries, uint32_t qs_base) {
  const uint32_t npairs = (nqueries + 1) / 2;
  const uint32_t total = a.g.nwarm + a.g.nstage;
  const uint64_t tiles = ((uint64_t)a.g.rows + kScanThreads - 1) / kScanThreads;
  for (uint32_t pq = 0; pq < npairs; pq++) {
    const uint32_t qa = 2 * pq;
    const bool has_b = qa + 1 < nqueries;
    const uint32_t qb = has_b ? qa + 1 : qa;
    EqPair pair[256];
    for (uint32_t c = 0; c < 256; c++) {
      const uint32_t row = (c >> a.sh0) & (a.msk0 & 0xFFu);
      pair[c].x = eq_all[(size_t)qa * a.nrows + row];
      pair[c].y = eq_all[(size_t)qb * a.nrows + row];
    }
    for (uint64_t row = 0; row < tiles * kScanThreads; row++) {
      if (a.tile_list && !a.dense[row / kScanThreads]) continue;
      Lane2 s;
      lane2_reset(s, a.m);
      int prev_a = a.m, prev_b = a.m;
      for (uint32_t it = 0; it < total; it++) {
        int64_t r;
        uint32_t col;
        bool own;
        stage_coord<REV>(a.g, it, (int64_t)row, r, col, own);
        const uint64_t stage_idx = (uint64_t)(r * (int64_t)a.g.ltot + (int64_t)col);
        const bool special = stage_is_special(a, stage_idx);
        const bool valid = r >= 0 && r < (int64_t)a.g.rows;
        for (int cc = 0; cc < kStageBytes / 16; cc++) {
          const int c = REV ? (kStageBytes / 16 - 1 - cc) : cc;
          uint32_t x[4] = {0, 0, 0, 0};
          if (valid) memcpy(x, a.text + stage_idx + 16u * c, 16);
          if (special)
            process16_2<REV, true>(s, prev_a, prev_b, x, stage_idx + 16u * c, a, pair, 0, qs_base + qa, has_b, own);
          else
            process16_2<REV, false>(s, prev_a, prev_b, x, stage_idx + 16u * c, a, pair, 0, qs_base + qa, has_b, own);
        }
      }
    }
  }
}

template <int WF, bool REV>
void filter_rows(const ScanArgs& a, const uint32_t* feq_q, uint32_t qs, bool pair) {
  EqTab eqt;
  eqt.p = feq_q;
  eqt.saddr = 0;
  eqt.rowbytes = (uint32_t)WF * 4u;
  const uint32_t total = a.g.nwarm + a.g.nstage;
  const uint64_t tiles = ((uint64_t)a.g.rows + kScanThreads - 1) / kScanThreads;
  for (uint64_t row = 0; row < tiles * kScanThreads; row++) {
    FLane<WF> s;
    flane_reset<WF>(s, a);
    for (uint32_t it = 0; it < total; it++) {
      int64_t r;
      uint32_t col;
      bool own;
      stage_coord<REV>(a.g, it, (int64_t)row, r, col, own);
      const uint64_t stage_idx = (uint64_t)(r * (int64_t)a.g.ltot + (int64_t)col);
      const bool valid = r >= 0 && r < (int64_t)a.g.rows;
      uint32_t mask = 0, mask1 = 0;
      for (int cc = 0; cc < kStageBytes / 16; cc++) {
        const int c = REV ? (kStageBytes / 16 - 1 - cc) : cc;
        uint32_t x[4] = {0, 0, 0, 0};
        if (valid) memcpy(x, a.text + stage_idx + 16u * c, 16);
        uint32_t acc[2];
        if (pair)
          filter16_pair<WF, REV>(s, x, eqt, acc);
        else
          filter16<WF, REV>(s, x, eqt, acc);
        if (acc[0]) mask |= 1u << c;
        if (acc[1]) mask1 |= 1u << c;
      }
      if (a.fused) {
        if (mask) emit_stage_hits(a, HitQueue{nullptr, nullptr}, qs, stage_idx, mask, own);
        if (mask1) emit_stage_hits(a, HitQueue{nullptr, nullptr}, qs + a.nq, stage_idx, mask1, own);
      } else if (mask | mask1) {
        emit_stage_hits(a, HitQueue{nullptr, nullptr}, qs, stage_idx, mask | mask1, own);
      }
    }
  }
}

// The body of qgram_kernel (scan_kernels.cu) on the host: one forward pass, hits for slot qs and,
// with a.fused, for the reversed partner slot qs + a.nq.
template <int Q, int S>
void qgram_rows(const ScanArgs& a, const uint32_t* bitmap, uint32_t qs) {
  const uint32_t total = a.g.nwarm + a.g.nstage;
  const uint64_t tiles = ((uint64_t)a.g.rows + kScanThreads - 1) / kScanThreads;
  for (uint64_t row = 0; row < tiles * kScanThreads; row++) {
    QLane s;
    s.w = 0, s.prev = 0;
    for (uint32_t it = 0; it < total; it++) {
      int64_t r;
      uint32_t col;
      bool own;
      stage_coord<false>(a.g, it, (int64_t)row, r, col, own);
      const uint64_t stage_idx = (uint64_t)(r * (int64_t)a.g.ltot + (int64_t)col);
      const bool valid = r >= 0 && r < (int64_t)a.g.rows;
      uint32_t mask = 0;
      for (int c = 0; c < kStageBytes / 16; c++) {
        uint32_t x[4] = {0, 0, 0, 0};
        if (valid) memcpy(x, a.text + stage_idx + 16u * c, 16);
        if (qgram16<Q, S>(s, x, bitmap) & 1u) mask |= 1u << c;
      }
      if (mask) {
        emit_stage_hits(a, HitQueue{nullptr, nullptr}, qs, stage_idx, mask, own);
        if (a.fused) emit_stage_hits(a, HitQueue{nullptr, nullptr}, qs + a.nq, stage_idx, mask, own);
      }
    }
  }
}

// The body of qgram_seq_kernel: every 64-byte block of the text is filtered on its own, the window
// rebuilt from the 16 bytes before it (zeros before the text, as the kernel's first carry).
template <int Q, int S>
void qgram_seq_blocks(const ScanArgs& a, const uint32_t* bitmap, uint32_t qs) {
  const uint64_t total_tiles = (a.n + 2047) / 2048;
  for (uint64_t b = 0; b < total_tiles * 32; b++) {
    const uint64_t stage_idx = b * 64;
    uint32_t hist[4] = {0, 0, 0, 0};
    if (b > 0) memcpy(hist, a.text + stage_idx - 16, 16);
    QLane s;
    qlane_init(s, hist);
    uint32_t mask = 0;
    for (int c = 0; c < 4; c++) {
      uint32_t x[4];
      memcpy(x, a.text + stage_idx + 16u * c, 16);
      if (qgram16<Q, S>(s, x, bitmap) & 1u) mask |= 1u << c;
    }
    if (mask) {
      emit_stage_hits(a, HitQueue{nullptr, nullptr}, qs, stage_idx, mask, true);
      if (a.fused) emit_stage_hits(a, HitQueue{nullptr, nullptr}, qs + 1, stage_idx, mask, true);
    }
  }
}

bool qgram_seq_dispatch(int Q, int S, const ScanArgs& a, const uint32_t* bitmap, uint32_t qs) {
#define EMU_Q(QQ, SS)                        \
  if (Q == QQ && S == SS) {                  \
    qgram_seq_blocks<QQ, SS>(a, bitmap, qs); \
    return true;                             \
  }
  EMU_Q(8, 4) EMU_Q(8, 8) EMU_Q(8, 16) EMU_Q(7, 4) EMU_Q(6, 4)
#undef EMU_Q
  return false;
}

bool qgram_dispatch(int Q, int S, const ScanArgs& a, const uint32_t* bitmap, uint32_t qs) {
#define EMU_Q(QQ, SS)            \
  if (Q == QQ && S == SS) {      \
    qgram_rows<QQ, SS>(a, bitmap, qs); \
    return true;                 \
  }
  EMU_Q(8, 4) EMU_Q(8, 8) EMU_Q(8, 16) EMU_Q(7, 4) EMU_Q(6, 4)
#undef EMU_Q
  return false;
}

template <bool REV>
void filter_dispatch(int WF, const ScanArgs& a, const uint32_t* feq_q, uint32_t qs, bool pair) {
  switch (WF) {
    case 1: filter_rows<1, REV>(a, feq_q, qs, pair); break;
    case 2: filter_rows<2, REV>(a, feq_q, qs, pair); break;
    case 4: filter_rows<4, REV>(a, feq_q, qs, pair); break;
    case 8: filter_rows<8, REV>(a, feq_q, qs, pair); break;
    default: abort();
  }
}

void verify_dispatch(int W, const ScanArgs& a, const uint32_t* eq_q, uint32_t qs, bool rev, uint64_t word,
                     uint32_t span = 0) {
  switch (W) {
    case 1: verify_hit<1>(a, eq_q, qs, rev, word, span); break;
    case 2: verify_hit<2>(a, eq_q, qs, rev, word, span); break;
    case 3: verify_hit<3>(a, eq_q, qs, rev, word, span); break;
    case 4: verify_hit<4>(a, eq_q, qs, rev, word, span); break;
    case 6: verify_hit<6>(a, eq_q, qs, rev, word, span); break;
    case 8: verify_hit<8>(a, eq_q, qs, rev, word, span); break;
    case 16: verify_hit<16>(a, eq_q, qs, rev, word, span); break;
    case 32: verify_hit<32>(a, eq_q, qs, rev, word, span); break;
    case 64: verify_hit<64>(a, eq_q, qs, rev, word, span); break;
    case 128: verify_hit<128>(a, eq_q, qs, rev, word, span); break;
    default: abort();
  }
}

template <bool REV>
void scan_dispatch(int W, const ScanArgs& a, const uint32_t* eq_q, uint32_t qs) {
  switch (W) {
    case 1: scan_rows<1, REV>(a, eq_q, qs); break;
    case 2: scan_rows<2, REV>(a, eq_q, qs); break;
    case 3: scan_rows<3, REV>(a, eq_q, qs); break;
    case 4: scan_rows<4, REV>(a, eq_q, qs); break;
    case 6: scan_rows<6, REV>(a, eq_q, qs); break;
    case 8: scan_rows<8, REV>(a, eq_q, qs); break;
    case 16: scan_rows<16, REV>(a, eq_q, qs); break;
    case 32: scan_rows<32, REV>(a, eq_q, qs); break;
    default: abort();
  }
}

}  // namespace

// refine_kernel + verify on the refined list (Engine::search, Dna): every hit chunk becomes zero or
// more nominal end positions, each re-scanned over 2k + 1 end positions.
void refine_and_verify(int W, ScanArgs& a, const std::vector<uint32_t>& eq, const ProfileParams& pp, const uint8_t* rev,
                       const std::vector<uint64_t>& hits, unsigned long long nhits) {
  std::vector<uint64_t> exact;
  std::vector<uint32_t> spans;
  for (unsigned long long h = 0; h < nhits; h++) {
    const uint32_t qs = key_qs(hits[h]);
    if (hit_in_dense_tile(a, key_pos(hits[h]) * kHitChars)) continue;
    int64_t lo, hi;
    if (refine_hit(a, qs, rev[qs] != 0, key_pos(hits[h]) * kHitChars, lo, hi))
      exact.push_back(cand_key(qs, (uint64_t)lo)), spans.push_back((uint32_t)(hi - lo));
  }
  a.hit_exact = 1;
  for (size_t i = 0; i < exact.size(); i++) {
    const uint32_t qs = key_qs(exact[i]);
    verify_dispatch(W, a, &eq[(size_t)qs * pp.nrows * W], qs, rev[qs] != 0, key_pos(exact[i]), spans[i]);
  }
  a.hit_exact = 0;
}

struct EmuResult {
  std::vector<GpuMatch> m;
  std::vector<uint32_t> ops;
  uint32_t ops_words;
  uint64_t candidates;
  uint64_t hits;
  int filter_words;  // 0 = full scan
  int filter_len;
  uint32_t dense_tiles = 0;
  ScanGeom g;
};

// The two end zones the overhang changes, per query: the body of overhang_edges_kernel
// (scan_kernels.cu) on the host.
template <int W>
void edges_rows(const ScanArgs& a, const uint32_t* eq, uint32_t slot, bool rev, const uint8_t* text, uint64_t n,
                const uint32_t* init_pv, int left_total, uint32_t steps, float alpha) {
  if (n + steps == 0) return;
  const uint64_t span = (uint64_t)a.m + (uint64_t)a.k;
  Lane<W> init;
  for (int w = 0; w < W; w++) init.pv[w] = init_pv[w], init.mv[w] = 0;
  uint32_t wild[W];
  for (int w = 0; w < W; w++) wild[w] = 0xFFFFFFFFu;
  {  // left: end positions 0 .. min(n, m+k)
    if (left_total <= a.k) emit_candidate(a, slot, 0, left_total);
    Lane<W> s = init;
    const uint64_t lim = n < span ? n : span;
    for (uint64_t i = 0; i < lim; i++) {
      const uint8_t c = text_at_dir(text, n, rev, i);
      myers_step<W>(s, eq + (((uint32_t)c >> a.sh0) & (a.msk0 & 0xFFu)) * W);
      const int sc = lane_score<W>(s);
      if (sc <= a.k) emit_candidate(a, slot, i + 1, sc);
    }
  }
  if (steps > 0) {  // right: end positions n+1 .. n+steps
    const uint64_t w0 = n > span ? n - span : 0;
    Lane<W> s;
    if (w0 == 0)
      s = init;
    else
      lane_reset<W>(s, a.m);
    for (uint64_t i = w0; i < n; i++) {
      const uint8_t c = text_at_dir(text, n, rev, i);
      myers_step<W>(s, eq + (((uint32_t)c >> a.sh0) & (a.msk0 & 0xFFu)) * W);
    }
    for (uint32_t o = 1; o <= steps; o++) {
      myers_step<W>(s, wild);
      const int sc = lane_score<W>(s) + overhang_overshoot_cost(alpha, o);
      if (sc <= a.k) emit_candidate(a, slot, n + o, sc);
    }
  }
}

void edges_dispatch(int W, const ScanArgs& a, const uint32_t* eq, uint32_t slot, bool rev, const uint8_t* text,
                    uint64_t n, const uint32_t* init_pv, int left_total, uint32_t steps, float alpha) {
  switch (W) {
    case 1: edges_rows<1>(a, eq, slot, rev, text, n, init_pv, left_total, steps, alpha); break;
    case 2: edges_rows<2>(a, eq, slot, rev, text, n, init_pv, left_total, steps, alpha); break;
    case 3: edges_rows<3>(a, eq, slot, rev, text, n, init_pv, left_total, steps, alpha); break;
    case 4: edges_rows<4>(a, eq, slot, rev, text, n, init_pv, left_total, steps, alpha); break;
    case 6: edges_rows<6>(a, eq, slot, rev, text, n, init_pv, left_total, steps, alpha); break;
    case 8: edges_rows<8>(a, eq, slot, rev, text, n, init_pv, left_total, steps, alpha); break;
    case 16: edges_rows<16>(a, eq, slot, rev, text, n, init_pv, left_total, steps, alpha); break;
    case 32: edges_rows<32>(a, eq, slot, rev, text, n, init_pv, left_total, steps, alpha); break;
    default: abort();
  }
}

extern "C" {

// queries: nq * m bytes; rev[q] = 1 scans the reversed text.  ltot_override > 0
// forces the row length (to exercise row boundaries on tiny texts); bpw plays
// the part of "resident blocks per wave" in the tiling heuristic.  use_filter: 0 = full
// scan, 1 = prefilter + verification whenever a piece layout exists, -1 = the engine's rule.
// Options of the post-processing (SearchOpts in engine.h); the same per-thread functions as the
// kernels run here on the host: end_filter_pass, n_fraction_ok, trace_one_ov, and the edge
// computation of overhang_edges_kernel (scan_kernels.cu) restated below.
struct EmuOpts {
  int without_trace, only_best, n_endpoint;
  float max_n_frac;      // < 0: off
  const uint8_t* pam;
  int pam_len;
  float alpha;           // < 0: off
  int max_overhang;      // < 0: unlimited
};

EmuResult* emu_search_opts(int profile, const uint8_t* queries, const uint8_t* rev, uint32_t nq, int m,
                           const uint8_t* text, uint64_t n, int k, int all_minima, int include_pos0,
                           uint32_t ltot_override, int bpw, int use_filter, const EmuOpts* eo);

EmuResult* emu_search(int profile, const uint8_t* queries, const uint8_t* rev, uint32_t nq, int m,
                      const uint8_t* text, uint64_t n, int k, int all_minima, int include_pos0,
                      uint32_t ltot_override, int bpw, int use_filter) {
  return emu_search_opts(profile, queries, rev, nq, m, text, n, k, all_minima, include_pos0, ltot_override, bpw,
                         use_filter, nullptr);
}

EmuResult* emu_search_opts(int profile, const uint8_t* queries, const uint8_t* rev, uint32_t nq, int m,
                           const uint8_t* text, uint64_t n, int k, int all_minima, int include_pos0,
                           uint32_t ltot_override, int bpw, int use_filter, const EmuOpts* eo) {
  const bool ov = eo && eo->alpha >= 0.f;
  if (ov) use_filter = 0;  // as Engine::search: no prefilter with overhang
  ProfileParams pp;
  if (!profile_params(profile, pp) || m <= 0) return nullptr;
  const int W = round_words((m + 31) / 32);
  if (W < 0) return nullptr;
  EmuResult* res = new EmuResult;
  res->ops_words = (uint32_t)((m + k + 1 + 15) / 16);
  res->candidates = 0;

  std::vector<uint32_t> eq((size_t)nq * pp.nrows * W);
  for (uint32_t q = 0; q < nq; q++) build_eq_table(profile, queries + (size_t)q * m, m, W, pp.nrows, &eq[(size_t)q * pp.nrows * W]);

  ScanGeom g = choose_geom(n, m, k, nq, bpw > 0 ? bpw : 444);
  if (ltot_override) {
    g.ltot = std::max<uint32_t>(ltot_override / kStageBytes * kStageBytes, g.nwarm * kStageBytes);  // any multiple of the stage size
    g.rows = (uint32_t)((n + g.ltot - 1) / g.ltot);
    g.nstage = g.ltot / kStageBytes;
  }
  res->g = g;
  std::vector<uint8_t> padded(std::max<size_t>(padded_alloc(n), (size_t)g.rows * g.ltot + 256), 0);
  if (n) memcpy(padded.data(), text, n);

  // with the prefilter every hit word may report up to 4+m+k positions (overlapping windows)
  // (with overhang: up to m more end positions beyond the text per query)
  const uint64_t cap = (uint64_t)nq * (n / 4 + 2) * (uint64_t)(use_filter ? 8 + (m + k) / 4 : 4) + 64 +
                       (ov ? (uint64_t)nq * (uint64_t)(m + k + 8) : 0);
  std::vector<uint64_t> keys(cap);
  std::vector<uint32_t> cost(cap);
  unsigned long long count = 0;
  if (include_pos0 && m <= k && n > 0 && !ov)
    for (uint32_t q = 0; q < nq; q++) {
      keys[count] = cand_key(q, 0);
      cost[count] = (uint32_t)m;
      count++;
    }

  ScanArgs a;
  memset(&a, 0, sizeof a);
  a.text = padded.data();
  a.n = n;
  a.g = g;
  a.sh0 = pp.sh0;
  a.msk0 = pp.msk0;
  a.nrows = pp.nrows;
  a.rowbytes = (uint32_t)W * 4u;
  a.m = m;
  a.k = k;
  a.cand_keys = keys.data();
  a.cand_cost = cost.data();
  a.cand_count = &count;
  a.cand_cap = cap;
  if (ov) a.emit_min = std::min<uint64_t>(n, (uint64_t)m + (uint64_t)k);
  std::vector<const uint8_t*> qptr(nq);
  for (uint32_t q = 0; q < nq; q++) qptr[q] = queries + (size_t)q * m;
  // Regional fallback of Engine::search: tiles of the scan geometry with more hits than a re-scan is
  // worth are marked dense (their hits are skipped) and scanned whole.
  const bool force_regional = getenv("SASSY_EMU_FORCE_REGIONAL") != nullptr;  // tests: small texts never reach heavy_hits
  std::vector<uint8_t> dense_flags;
  std::vector<uint32_t> dense_list(1, 0);
  uint32_t dense_count = 0;
  auto mark_dense = [&](const std::vector<uint64_t>& hits, unsigned long long nhits, bool qgram) {
    const uint32_t ntiles = (g.rows + kScanThreads - 1) / kScanThreads;
    const uint64_t tile_bytes = (uint64_t)kScanThreads * g.ltot;
    const double hit_cost = qgram ? 32.0 + 0.25 * (2.0 * (m + k) + kHitChars) : 2.0 * (m + k) + kHitChars;
    const double max_hits = 0.5 * (double)tile_bytes * nq / hit_cost;
    dense_flags.assign(ntiles + 1, 0);
    a.tile_bytes = tile_bytes;
    a.dense = dense_flags.data();
    // Engine::search: the regional pass runs only above `heavy_hits` (and only up to 32 words)
    if (W > kMaxScanWords) return;
    if ((double)nhits <= std::max(1024.0, 0.25 * (double)n * nq / hit_cost) && !force_regional) return;
    std::vector<uint32_t> counts(ntiles + 1, 0);
    for (unsigned long long h = 0; h < nhits; h++) counts[(key_pos(hits[h]) * kHitChars) / tile_bytes]++;
    for (uint32_t t = 0; t < ntiles; t++)
      if ((double)counts[t] > (double)(uint32_t)std::min(max_hits, 4.0e9)) dense_flags[t] = 1, dense_count++;
  };
  auto scan_dense = [&](ScanArgs& aa) {
    res->dense_tiles = dense_count;
    if (!dense_count) return;
    aa.tile_list = dense_list.data();  // marker: scan_rows visits the rows of dense tiles only
    for (uint32_t q = 0; q < nq; q++) {
      const uint32_t* eq_q = &eq[(size_t)q * pp.nrows * W];
      if (rev[q]) {
        aa.reset_idx = n - 1;
        scan_dispatch<true>(W, aa, eq_q, q);
      } else {
        aa.reset_idx = 0;
        scan_dispatch<false>(W, aa, eq_q, q);
      }
    }
    aa.tile_list = nullptr;
  };
  FilterPlan fp;
  // use_filter: 0 off, < 0 automatic, 1 force the piece automaton (4: and refine its hits), 2 / 3 as Engine::search with the
  // filter forced (q-gram bitmap when it can be planned -- 2: contiguous tiles, 3: row tiles --
  // else the piece automaton)
  if (use_filter != 0) fp = plan_filter(profile, qptr.data(), nq, m, k, use_filter > 0 ? 1e30 : 0.85);
  res->hits = 0;
  QgramPlan qp;
  {
    uint32_t nf = 0;
    while (nf < nq && !rev[nf]) nf++;
    if (n > 0 && (use_filter == 2 || use_filter == 3 || use_filter < 0) && !ov && profile == kDna && nf == 1 && nq <= 2)
      qp = plan_qgram(m, k, (int)nq);
  }
  if (qp.enabled) fp.enabled = false;
  res->filter_words = fp.enabled ? fp.WF : (qp.enabled ? 1 : 0);  // per strand
  res->filter_len = fp.enabled ? fp.L : (qp.enabled ? qp.q : 0);
  if (qp.enabled) {
    std::vector<uint32_t> bitmap(qp.table_words(), 0);
    std::vector<uint32_t> conf((size_t)nq * qp.npieces * kConfWords);
    for (uint32_t q = 0; q < nq; q++) {
      add_qgram_entries(qp, qptr[q], rev[q] != 0, bitmap.data());
      build_qgram_confirm(qp, qptr[q], rev[q] != 0, &conf[(size_t)q * qp.npieces * kConfWords]);
    }
    std::vector<uint64_t> hits((size_t)nq * (n / kHitChars + 2) + 16);
    unsigned long long nhits = 0;
    ScanArgs f = a;
    f.g.nwarm = 1;
    f.fused = nq == 2 ? 1 : 0;
    f.nq = 1;
    f.hit_keys = hits.data();
    f.hit_count = &nhits;
    f.hit_cap = hits.size();
    // 3: the row-tiled kernel (LDG data path of the product), else the contiguous-tile kernel
    if (!(use_filter == 3 ? qgram_dispatch(qp.q, qp.s, f, bitmap.data(), 0)
                          : qgram_seq_dispatch(qp.q, qp.s, f, bitmap.data(), 0)))
      abort();
    a.qconf = conf.data();
    a.qnp = (uint32_t)qp.npieces;
    res->hits = nhits;
    mark_dense(hits, nhits, true);
    refine_and_verify(W, a, eq, pp, rev, hits, nhits);
    a.qconf = nullptr, a.qnp = 0;
    scan_dense(a);
  } else if (n > 0 && fp.enabled) {
    uint32_t nfwd = 0;
    while (nfwd < nq && !rev[nfwd]) nfwd++;
    const bool fused = nfwd > 0 && nq == 2 * nfwd && fp.WF <= 2;  // as Engine::search
    const int WT = fused ? 2 * fp.WF : fp.WF;
    const size_t ntab = fused ? nfwd : nq;
    const bool pair = profile == kDna && WT <= 4;
    const size_t tab_words = pair ? (size_t)kPairTableWords * WT : (size_t)256 * WT;
    std::vector<uint32_t> feq(ntab * tab_words + 4);
    for (uint32_t q = 0; q < ntab; q++) {
      const uint8_t* partner = fused ? qptr[q + nfwd] : nullptr;
      if (pair)
        build_pair_table(fp, qptr[q], &feq[q * tab_words], partner);
      else
        build_filter_table(profile, fp, qptr[q], &feq[q * tab_words], partner);
    }
    std::vector<uint64_t> hits((size_t)nq * (n / kHitChars + 2) + 16);
    unsigned long long nhits = 0;
    ScanArgs f = a;
    f.g.nwarm = 1;
    for (int w = 0; w < kMaxFilterWords; w++) f.finit[w] = fp.finit[w], f.fdelay[w] = fp.fdelay[w];
    if (fused)
      for (int w = 0; w < fp.WF; w++) f.finit[fp.WF + w] = fp.finit[w], f.fdelay[fp.WF + w] = fp.fdelay[w];
    f.fused = fused ? 1 : 0;
    f.nq = fused ? nfwd : nq;
    f.hit_keys = hits.data();
    f.hit_count = &nhits;
    f.hit_cap = hits.size();
    for (uint32_t q = 0; q < ntab; q++) {
      const uint32_t* feq_q = &feq[q * tab_words];
      if (rev[q])
        filter_dispatch<true>(WT, f, feq_q, q, pair);
      else
        filter_dispatch<false>(WT, f, feq_q, q, pair);
    }
    if (fused)
      for (int p = 0; p < fp.npieces; p++) a.rev_lead = std::max<uint32_t>(a.rev_lead, (uint32_t)fp.piece[p].len);
    res->hits = nhits;
    mark_dense(hits, nhits, false);
    if (profile == kDna && use_filter == 4) {  // SASSY_B200_REFINE=2: piece-automaton hits are refined too
      std::vector<uint32_t> conf((size_t)nq * fp.npieces * kConfWords);
      for (uint32_t q = 0; q < nq; q++)
        build_filter_confirm(fp, qptr[q], rev[q] != 0, rev[q] != 0 && !fused, &conf[(size_t)q * fp.npieces * kConfWords]);
      a.qconf = conf.data();
      a.qnp = (uint32_t)fp.npieces;
      a.rev_lead = 0;
      refine_and_verify(W, a, eq, pp, rev, hits, nhits);
      a.qconf = nullptr, a.qnp = 0;
    } else {
      for (unsigned long long h = 0; h < nhits; h++) {
        const uint32_t qs = key_qs(hits[h]);
        if (hit_in_dense_tile(a, key_pos(hits[h]) * kHitChars)) continue;
        verify_dispatch(W, a, &eq[(size_t)qs * pp.nrows * W], qs, rev[qs] != 0, key_pos(hits[h]));
      }
    }
    scan_dense(a);
  } else if (n > 0 && W > kMaxScanWords) {
    // as Engine::search: windows of kCoverStride end positions cover the text (cover_kernel)
    a.hit_exact = 1;
    const uint64_t per = (n + kCoverStride - 1) / kCoverStride;
    for (uint32_t q = 0; q < nq; q++)
      for (uint64_t j = 0; j < per; j++)
        verify_dispatch(W, a, &eq[(size_t)q * pp.nrows * W], q, rev[q] != 0, j * kCoverStride + (uint64_t)k + 1,
                        (uint32_t)(kCoverStride - 1 - 2 * (uint64_t)k));
    a.hit_exact = 0;
  } else if (n > 0) {
    uint32_t nf = 0;
    while (nf < nq && !rev[nf]) nf++;
    const bool no_scan2 = getenv("SASSY_EMU_NO_SCAN2") != nullptr;
    // as Engine::search: batches of one-word patterns take two patterns per thread
    if (W == 1 && nf >= 2 && !no_scan2) {
      a.reset_idx = 0;
      scan2_rows<false>(a, &eq[0], nf, 0);
    }
    if (W == 1 && nq - nf >= 2 && !no_scan2) {
      a.reset_idx = n - 1;
      scan2_rows<true>(a, &eq[(size_t)nf * pp.nrows], nq - nf, nf);
    }
    for (uint32_t q = 0; q < nq; q++) {
      const bool paired = W == 1 && !no_scan2 && (rev[q] ? nq - nf >= 2 : nf >= 2);
      if (paired) continue;
      const uint32_t* eq_q = &eq[(size_t)q * pp.nrows * W];
      if (rev[q]) {
        a.reset_idx = n - 1;
        scan_dispatch<true>(W, a, eq_q, q);
      } else {
        a.reset_idx = 0;
        scan_dispatch<false>(W, a, eq_q, q);
      }
    }
  }
  if (ov) {  // the two edges per query (overhang_edges_kernel)
    std::vector<uint32_t> init_pv(W, 0);
    const int pad = 32 * W - m;
    for (int j = 0; j < m; j++)
      if (overhang_left_cost(j + 1, eo->alpha, eo->max_overhang) - overhang_left_cost(j, eo->alpha, eo->max_overhang))
        init_pv[(pad + j) >> 5] |= 1u << ((pad + j) & 31);
    const int left_total = overhang_left_cost(m, eo->alpha, eo->max_overhang);
    const float r = ceilf(((float)k + eo->alpha) / eo->alpha);
    uint64_t steps = std::isnan(r) ? 0 : (r >= 1e18f ? ~0ull : (uint64_t)r);
    steps = std::min<uint64_t>(steps, (uint64_t)m);
    if (eo->max_overhang >= 0) steps = std::min<uint64_t>(steps, (uint64_t)eo->max_overhang);
    for (uint32_t q = 0; q < nq; q++)
      edges_dispatch(W, a, &eq[(size_t)q * pp.nrows * W], q, rev[q] != 0, padded.data(), n, init_pv.data(),
                     left_total, (uint32_t)steps, eo->alpha);
  }
  res->candidates = count;
  if (count > cap) {  // the emulator's buffer is sized for the worst case; never read beyond it
    delete res;
    return nullptr;
  }

  // sort by key (the GPU path uses a device radix sort)
  std::vector<uint64_t> order(count);
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](uint64_t x, uint64_t y) { return keys[x] < keys[y]; });
  std::vector<uint64_t> skeys(count);
  std::vector<uint32_t> scost(count);
  for (uint64_t i = 0; i < count; i++) skeys[i] = keys[order[i]], scost[i] = cost[order[i]];
  EndFilter ef;
  memset(&ef, 0, sizeof ef);
  const bool end_filter = eo && (eo->pam_len > 0 || (eo->n_endpoint && eo->max_n_frac >= 0.f));
  if (end_filter) {
    ef.text = TextRef{padded.data(), n, nullptr, nullptr, nq};
    ef.rev_flags = rev;
    ef.profile = profile;
    ef.m = m, ef.k = k;
    ef.pam_len = eo->pam_len;
    for (int i = 0; i < eo->pam_len; i++) ef.pam[0][i] = eo->pam[i], ef.pam[1][i] = complement_byte(profile, eo->pam[i]);
    ef.n_endpoint = (eo->n_endpoint && eo->max_n_frac >= 0.f) ? 1 : 0;
    ef.max_n_frac = eo->max_n_frac;
  }
  std::vector<uint64_t> sel;
  std::vector<uint32_t> sel_cost;
  for (uint64_t i = 0; i < count; i++)
    if (select_candidate(skeys.data(), scost.data(), i, count, all_minima != 0) &&
        (!end_filter || end_filter_pass(ef, skeys[i])))
      sel.push_back(skeys[i]), sel_cost.push_back(scost[i]);
  if (eo && eo->only_best) {  // launch_best: minimal cost, rightmost end, per query slot
    std::vector<uint64_t> bk;
    std::vector<uint32_t> bc;
    for (uint32_t q = 0; q < nq; q++) {
      int best = -1;
      for (size_t i = 0; i < sel.size(); i++)
        if (key_qs(sel[i]) == q && (best < 0 || sel_cost[i] < sel_cost[best] ||
                                    (sel_cost[i] == sel_cost[best] && key_pos(sel[i]) > key_pos(sel[best]))))
          best = (int)i;
      if (best >= 0) bk.push_back(sel[best]), bc.push_back(sel_cost[best]);
    }
    sel = bk, sel_cost = bc;
  }

  std::vector<uint32_t> scratch((size_t)trace_words_per_match(m, k, W));
  for (size_t i = 0; i < sel.size(); i++) {
    const uint32_t qs = key_qs(sel[i]);
    const uint64_t end = key_pos(sel[i]);
    ColStore cs;
    cs.base = scratch.data();
    cs.stride = 1;
    TraceOut out;
    const uint8_t* pat = queries + (size_t)qs * m;
    const uint32_t* eq_q = &eq[(size_t)qs * pp.nrows * W];
    std::vector<uint32_t> ops(res->ops_words, 0);
    const bool rv = rev[qs] != 0;
    GpuMatch gm;
    gm.qs = qs;
    if (eo && eo->without_trace) {  // trace_kernel, costs branch
      gm.text_start = ~0ull;
      gm.text_end = end < n ? end : n;
      gm.cost = (int32_t)sel_cost[i];
      gm.nops = 0;
      gm.failed = pack_overhang(0, end > n ? (uint32_t)(end - n) : 0u);
    } else {
#define EMU_TRACE(P)                                                                                              \
  if (ov)                                                                                                         \
    trace_one_ov<P>(padded.data(), n, rv, pat, m, k, eq_q, W, pp.sh0, pp.msk0, end, eo->alpha, eo->max_overhang,  \
                    cs, ops.data(), res->ops_words, out);                                                        \
  else                                                                                                            \
    trace_one<P>(padded.data(), n, rv, pat, m, k, eq_q, W, pp.sh0, pp.msk0, end, cs, ops.data(), res->ops_words, out);
      switch (profile) {
        case kDna: EMU_TRACE(kDna) break;
        case kIupac: EMU_TRACE(kIupac) break;
        default: EMU_TRACE(kAscii) break;
      }
#undef EMU_TRACE
      gm.text_start = out.text_start;
      gm.text_end = out.text_end;
      gm.cost = out.cost;
      gm.nops = out.nops;
      gm.failed = out.failed;
      if (eo && eo->max_n_frac >= 0.f &&
          !n_fraction_ok(padded.data(), n, rv, out.text_start, out.text_end < n ? out.text_end : n, eo->max_n_frac, 0))
        continue;  // traced N filter: dropped
    }
    res->m.push_back(gm);
    res->ops.insert(res->ops.end(), ops.begin(), ops.end());
  }
  return res;
}

// The prefilter plan for inspection in tests: out = {enabled, WF, npieces, L, then (off, len, word, bit) per piece};
// returns the number of ints written (<= cap).
int emu_plan_filter(int profile, const uint8_t* queries, uint32_t nq, int m, int k, double max_cost, int* out, int cap) {
  std::vector<const uint8_t*> qptr(nq);
  for (uint32_t q = 0; q < nq; q++) qptr[q] = queries + (size_t)q * m;
  const FilterPlan fp = plan_filter(profile, qptr.data(), nq, m, k, max_cost);
  int n = 0;
  auto put = [&](int v) {
    if (n < cap) out[n] = v;
    n++;
  };
  put(fp.enabled ? 1 : 0), put(fp.WF), put(fp.npieces), put(fp.L);
  for (int p = 0; p < fp.npieces; p++) put(fp.piece[p].off), put(fp.piece[p].len), put(fp.piece[p].word), put(fp.piece[p].bit);
  return n;
}

// The q-gram plan for inspection in tests: out = {enabled, q, s, npieces, then (off, len) per share}.
int emu_plan_qgram(int m, int k, int strands, int* out, int cap) {
  const QgramPlan f = plan_qgram(m, k, strands);
  int n = 0;
  auto put = [&](int v) {
    if (n < cap) out[n] = v;
    n++;
  };
  put(f.enabled ? 1 : 0), put(f.q), put(f.s), put(f.npieces);
  for (int p = 0; p < f.npieces; p++) put(f.off[p]), put(f.len[p]);
  return n;
}

// Transport encoding (dna_pack.h): every host packer against the device-side decoder.
int emu_dna_pack_best() { return dna_pack_best_level(); }
int emu_dna_pack(int level, const uint8_t* src, size_t n, uint8_t* dst) {
  if (level > dna_pack_best_level()) return -1;
  return dna_pack(src, dst, n, level) ? 1 : 0;
}
void emu_dna_unpack(const uint8_t* packed, size_t n, uint8_t* out) {
  for (size_t i = 0; i < n; i++) {
    const size_t g = i / 8;
    const uint32_t planes = (uint32_t)packed[2 * g] | ((uint32_t)packed[2 * g + 1] << 8);
    out[i] = (uint8_t)(dna_unpack4(planes, (int)((i % 8) / 4)) >> (8 * (i % 4)));
  }
}

// Many texts scanned as one concatenation (Engine::search_texts): candidate position -> (text, position in the text).
int emu_concat_locate(uint64_t pos, int rev, uint64_t total, const uint64_t* offs, const uint64_t* lens, uint32_t ntexts,
                      uint32_t* ti, uint64_t* local) {
  return concat_locate(pos, rev != 0, total, offs, lens, ntexts, *ti, *local) ? 1 : 0;
}

size_t emu_len(const EmuResult* r) { return r->m.size(); }
const GpuMatch* emu_matches(const EmuResult* r) { return r->m.data(); }
const uint32_t* emu_ops(const EmuResult* r) { return r->ops.data(); }
uint32_t emu_ops_words(const EmuResult* r) { return r->ops_words; }
uint64_t emu_candidates(const EmuResult* r) { return r->candidates; }
uint64_t emu_hits(const EmuResult* r) { return r->hits; }
int emu_filter_words(const EmuResult* r) { return r->filter_words; }
int emu_filter_len(const EmuResult* r) { return r->filter_len; }
uint32_t emu_dense_tiles(const EmuResult* r) { return r->dense_tiles; }
uint32_t emu_ltot(const EmuResult* r) { return r->g.ltot; }
uint32_t emu_rows(const EmuResult* r) { return r->g.rows; }
void emu_free(EmuResult* r) { delete r; }

}  // extern "C"
