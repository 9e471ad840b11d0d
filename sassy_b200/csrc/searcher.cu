#include "searcher.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "../../include/sassy_gpu.h"
#include "shard_merge.h"

namespace sb {

int parse_alphabet(const std::string& alphabet) {
  std::string a = alphabet;
  std::transform(a.begin(), a.end(), a.begin(), [](unsigned char c) { return (char)tolower(c); });
  if (a == "dna") return kDna;
  if (a == "iupac") return kIupac;
  if (a == "ascii") return kAscii;
  return -1;
}

std::string Match::cigar() const {
  std::string out;
  size_t i = 0;
  while (i < ops.size()) {
    size_t j = i;
    while (j < ops.size() && ops[j] == ops[i]) j++;
    out += std::to_string(j - i);
    out += ops[i];
    i = j;
  }
  return out;
}

Searcher::Searcher(const std::string& alphabet, bool rc, float alpha, int device) : rc_(rc) {
  profile_ = parse_alphabet(alphabet);
  if (profile_ < 0) throw std::invalid_argument("Unsupported alphabet: " + alphabet);
  if (!isnan(alpha)) {
    // reference Searcher::_overhang_check (src/search.rs:373-383)
    if (profile_ != kIupac) throw std::invalid_argument("Overhang is not supported for this profile (use iupac)");
    if (!(alpha >= 0.f && alpha <= 1.f)) throw std::invalid_argument("Alpha must be in range 0.0 <= alpha <= 1.0");
    alpha_ = alpha;
  }
  engine_.reset(new Engine(profile_, device));
}

// Reference: Iupac patterns must be valid IUPAC (src/profiles/iupac.rs:19-24 panics);
// Dna and Ascii accept every byte (src/profiles/dna.rs:19-23, ascii.rs:70-72).
void Searcher::validate_pattern(const uint8_t* p, size_t m) const {
  if (profile_ == kIupac && !iupac_valid(p, m)) throw InvalidPattern("Pattern is not valid IUPAC");
}

static const char kOpChars[4] = {'=', 'X', 'I', 'D'};

static void unpack_ops(const MatchSet& ms, size_t i, std::string& out) {
  const GpuMatch& g = ms.m[i];
  const uint32_t* w = &ms.ops[i * ms.ops_words];
  out.resize(g.nops);
  for (uint32_t a = 0; a < g.nops; a++) out[a] = kOpChars[(w[a >> 4] >> ((a & 15) * 2)) & 3u];
}

SearchOpts Searcher::v1_opts(bool all_minima) const {
  SearchOpts o;
  o.all_minima = all_minima;
  o.include_pos0 = true;
  o.without_trace = without_trace_;
  o.only_best = only_best_;
  o.max_n_frac = max_n_frac_;
  o.n_endpoint = true;  // v1 filters end points before the traceback as well (src/search.rs:907-919)
  o.alpha = alpha_;
  o.max_overhang = max_overhang_;
  return o;
}

// v1 strands (reference src/search.rs:787-881): queries [0, n_pat) = the patterns on the text,
// queries [n_pat, 2 n_pat) = complement(pattern) on the reversed text; reversed-text
// coordinates are mapped back (:859-877), the CIGAR is kept as produced (:874-876).  Without
// trace only the end position is known (:1464-1475): text_start / pattern_start are
// usize::MAX, and for the reverse strand the known end becomes text_start (:866-872).
// Slot = text index * n_queries + query.
template <class LenFn>
std::vector<Match> Searcher::convert_v1(const MatchSet& ms, size_t n_pat, size_t m, LenFn text_len) const {
  const size_t nq = rc_ ? 2 * n_pat : n_pat;
  std::vector<Match> out(ms.m.size());
  for (size_t i = 0; i < ms.m.size(); i++) {
    const GpuMatch& g = ms.m[i];
    if (g.failed & 1u) throw std::runtime_error("Trace failed (text contains bytes outside the profile's alphabet?)");
    Match& mm = out[i];
    const size_t ti = g.qs / nq, q = g.qs % nq;
    const uint64_t n = text_len(ti);
    mm.pattern_idx = q % n_pat;
    mm.text_idx = ti;
    mm.cost = g.cost;
    // overhang: rows of the pattern hanging over the text ends (scan_core.cuh pack_overhang)
    mm.pattern_start = without_trace_ ? ~0ull : (uint64_t)((g.failed >> 8) & 0xFFFu);
    mm.pattern_end = m - (uint64_t)(g.failed >> 20);
    if (q < n_pat) {
      mm.strand = kFwd;
      mm.text_start = g.text_start;
      mm.text_end = g.text_end;
    } else {
      mm.strand = kRc;
      mm.text_start = n - g.text_end;
      mm.text_end = without_trace_ ? ~0ull : n - g.text_start;
    }
    if (!without_trace_) unpack_ops(ms, i, mm.ops);
  }
  return out;
}

std::vector<Match> Searcher::search(const uint8_t* pattern, size_t m, const DeviceText& text, size_t k,
                                    bool all_minima) {
  return search_with_pam(pattern, m, text, k, all_minima, nullptr, 0);
}

std::vector<Match> Searcher::search_with_pam(const uint8_t* pattern, size_t m, const DeviceText& text, size_t k,
                                             bool all_minima, const uint8_t* pam, size_t pam_len) {
  if (m == 0) throw std::invalid_argument("empty pattern");
  validate_pattern(pattern, m);
  if (rc_ && profile_ == kAscii)
    throw std::invalid_argument("reverse complement is not implemented for the Ascii profile");
  if (pam_len > (size_t)kMaxPam) throw std::invalid_argument("PAM longer than 16 characters");
  std::vector<uint8_t> comp;
  std::vector<Query> qs;
  qs.push_back(Query{pattern, false});
  if (rc_) {
    comp.resize(m);
    for (size_t i = 0; i < m; i++) comp[i] = complement_byte(profile_, pattern[i]);
    qs.push_back(Query{comp.data(), true});
  }
  const int kk = (int)std::min<size_t>(k, 1u << 20);
  SearchOpts o = v1_opts(all_minima);
  o.pam = pam;
  o.pam_len = (int)pam_len;
  engine_->search(text, qs, (int)m, kk, o, ms_);
  if (raw_only_) return {};  // the caller reads ms_ / the gathered records itself
  const uint64_t n = text.n;
  return convert_v1(ms_, 1, m, [n](size_t) { return n; });
}

std::vector<Match> Searcher::search(const uint8_t* pattern, size_t m, const uint8_t* text, size_t n, size_t k,
                                    bool all_minima) {
  return search_with_pam(pattern, m, text, n, k, all_minima, nullptr, 0);
}

// As convert_v1 (one pattern, one text) + the C ABI's record layout, in one pass.
void Searcher::search_flat(const uint8_t* pattern, size_t m, const DeviceText& text, size_t k, bool all_minima,
                           const uint8_t* pam, size_t pam_len, FlatMatches& out) {
  raw_only_ = true;
  try {
    search_with_pam(pattern, m, text, k, all_minima, pam, pam_len);
  } catch (...) {
    raw_only_ = false;
    throw;
  }
  raw_only_ = false;
  const MatchSet& ms = ms_;
  const size_t nq = rc_ ? 2 : 1;
  const uint64_t n = text.n;
  const size_t cnt = ms.m.size();
  out.m.resize(cnt);
  size_t total = 0;
  for (size_t i = 0; i < cnt; i++) {
    if (ms.m[i].failed & 1u) throw std::runtime_error("Trace failed (text contains bytes outside the profile's alphabet?)");
    total += without_trace_ ? 0 : ms.m[i].nops;
  }
  out.ops.resize(total);
  char* ops = total ? &out.ops[0] : nullptr;
  size_t off = 0;
  for (size_t i = 0; i < cnt; i++) {
    const GpuMatch& g = ms.m[i];
    sassy_gpu_Match& o = out.m[i];
    memset(&o, 0, sizeof o);
    const size_t q = g.qs % nq;
    o.cost = g.cost;
    o.pattern_start = without_trace_ ? ~0ull : (uint64_t)((g.failed >> 8) & 0xFFFu);
    o.pattern_end = m - (uint64_t)(g.failed >> 20);
    if (q == 0) {
      o.strand = (uint8_t)kFwd;
      o.text_start = g.text_start;
      o.text_end = g.text_end;
    } else {
      o.strand = (uint8_t)kRc;
      o.text_start = n - g.text_end;
      o.text_end = without_trace_ ? ~0ull : n - g.text_start;
    }
    o.ops_off = off;
    if (!without_trace_) {
      o.ops_len = g.nops;
      const uint32_t* w = &ms.ops[i * ms.ops_words];
      for (uint32_t a = 0; a < g.nops; a++) ops[off + a] = kOpChars[(w[a >> 4] >> ((a & 15) * 2)) & 3u];
      off += g.nops;
    }
  }
}

void Searcher::search_flat(const uint8_t* pattern, size_t m, const uint8_t* text, size_t n, size_t k, bool all_minima,
                           const uint8_t* pam, size_t pam_len, FlatMatches& out) {
  if (m == 0) throw std::invalid_argument("empty pattern");
  validate_pattern(pattern, m);
  DeviceText* t = engine_->stage_text(text, n);
  search_flat(pattern, m, *t, k, all_minima, pam, pam_len, out);
}

std::vector<Match> Searcher::search_with_pam(const uint8_t* pattern, size_t m, const uint8_t* text, size_t n,
                                             size_t k, bool all_minima, const uint8_t* pam, size_t pam_len) {
  if (m == 0) throw std::invalid_argument("empty pattern");
  validate_pattern(pattern, m);
  DeviceText* t = engine_->stage_text(text, n);
  return search_with_pam(pattern, m, *t, k, all_minima, pam, pam_len);
}

// Equal-length patterns x texts with v1 semantics per (pattern, text) pair.  Short texts go
// through the one-thread-per-pair kernel in a single launch; a long text is staged and
// scanned by the row-tiled kernels with all patterns as queries.
std::vector<Match> Searcher::search_group(const uint8_t* const* patterns, size_t n_pat, size_t m,
                                          const uint8_t* const* texts, const uint64_t* text_lens, size_t n_texts,
                                          size_t k) {
  if (m == 0) throw std::invalid_argument("empty pattern");
  if (rc_ && profile_ == kAscii)
    throw std::invalid_argument("reverse complement is not implemented for the Ascii profile");
  std::vector<uint8_t> comp(rc_ ? n_pat * m : 0);
  std::vector<Query> qs;
  for (size_t p = 0; p < n_pat; p++) {
    validate_pattern(patterns[p], m);
    qs.push_back(Query{patterns[p], false});
  }
  if (rc_)
    for (size_t p = 0; p < n_pat; p++) {
      for (size_t i = 0; i < m; i++) comp[p * m + i] = complement_byte(profile_, patterns[p][i]);
      qs.push_back(Query{&comp[p * m], true});
    }
  const int kk = (int)std::min<size_t>(k, 1u << 20);
  const SearchOpts o = v1_opts(false);
  std::vector<Match> out;
  constexpr uint64_t kShortText = 1ull << 17;  // longer texts are worth a row-tiled scan of their own
  std::vector<const uint8_t*> sp;
  std::vector<uint64_t> sl;
  std::vector<size_t> sidx;
  for (size_t t = 0; t < n_texts; t++) {
    if (text_lens[t] <= kShortText) {
      sp.push_back(texts[t]), sl.push_back(text_lens[t]), sidx.push_back(t);
      continue;
    }
    DeviceText* dt = engine_->stage_text(texts[t], text_lens[t]);
    engine_->search(*dt, qs, (int)m, kk, o, ms_);
    const uint64_t n = text_lens[t];
    std::vector<Match> part = convert_v1(ms_, n_pat, m, [n](size_t) { return n; });
    for (auto& mm : part) mm.text_idx = t;
    out.insert(out.end(), std::make_move_iterator(part.begin()), std::make_move_iterator(part.end()));
  }
  // pairs per launch are bounded by the 24-bit slot field of the candidate keys
  const size_t max_texts = std::max<size_t>(1, ((1u << 23) - 1) / qs.size());
  for (size_t first = 0; first < sp.size(); first += max_texts) {
    const size_t cnt = std::min(max_texts, sp.size() - first);
    engine_->search_texts(sp.data() + first, sl.data() + first, cnt, qs, (int)m, kk, o, ms_);
    const double t0 = HostTimers::on() ? HostTimers::now_us() : 0.0;
    std::vector<Match> part = convert_v1(ms_, n_pat, m, [&](size_t ti) { return sl[first + ti]; });
    for (auto& mm : part) mm.text_idx = sidx[first + mm.text_idx];
    out.insert(out.end(), std::make_move_iterator(part.begin()), std::make_move_iterator(part.end()));
    if (HostTimers::on()) HostTimers::add(HostTimers::kTextsConvert, HostTimers::now_us() - t0);
  }
  return out;
}

static void sort_single_mode(std::vector<Match>& v) {
  // SearchMode::Single order (src/search.rs:541-553,1519-1549): pattern-major, then text, forward
  // matches before reverse-complement ones; positions keep their scan order (stable)
  std::stable_sort(v.begin(), v.end(), [](const Match& a, const Match& b) {
    if (a.pattern_idx != b.pattern_idx) return a.pattern_idx < b.pattern_idx;
    if (a.text_idx != b.text_idx) return a.text_idx < b.text_idx;
    return a.strand < b.strand;
  });
}

std::vector<Match> Searcher::search_patterns(const uint8_t* const* patterns, size_t n_patterns, size_t m,
                                             const uint8_t* text, size_t n, size_t k) {
  if (n_patterns == 0) return {};
  const uint64_t len = n;
  std::vector<Match> out = search_group(patterns, n_patterns, m, &text, &len, 1, k);
  sort_single_mode(out);
  return out;
}

std::vector<Match> Searcher::search_texts(const uint8_t* pattern, size_t m, const uint8_t* const* texts,
                                          const uint64_t* text_lens, size_t n_texts, size_t k) {
  if (n_texts == 0) return {};
  std::vector<Match> out = search_group(&pattern, 1, m, texts, text_lens, n_texts, k);
  sort_single_mode(out);
  return out;
}

std::vector<Match> Searcher::search_many(const uint8_t* const* patterns, const uint64_t* pattern_lens,
                                         size_t n_patterns, const uint8_t* const* texts, const uint64_t* text_lens,
                                         size_t n_texts, size_t k) {
  std::vector<Match> out;
  if (n_patterns == 0 || n_texts == 0) return out;
  // group the patterns by length: one launch set per group
  std::vector<size_t> order(n_patterns);
  for (size_t i = 0; i < n_patterns; i++) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return pattern_lens[a] < pattern_lens[b]; });
  for (size_t g0 = 0; g0 < n_patterns;) {
    size_t g1 = g0;
    while (g1 < n_patterns && pattern_lens[order[g1]] == pattern_lens[order[g0]]) g1++;
    std::vector<const uint8_t*> grp;
    for (size_t i = g0; i < g1; i++) grp.push_back(patterns[order[i]]);
    std::vector<Match> part = search_group(grp.data(), grp.size(), pattern_lens[order[g0]], texts, text_lens, n_texts, k);
    for (auto& mm : part) mm.pattern_idx = order[g0 + mm.pattern_idx];
    out.insert(out.end(), std::make_move_iterator(part.begin()), std::make_move_iterator(part.end()));
    g0 = g1;
  }
  sort_single_mode(out);
  return out;
}

// Reference: TQueries::new (src/pattern_tiling/tqueries.rs:53-134): equal
// lengths, reverse complements appended as queries n..2n when rc (:75-80, it
// always uses the IUPAC complement table, tqueries.rs:2).
EncodedPatterns Searcher::encode_patterns(const uint8_t* const* patterns, size_t n_patterns, size_t m) const {
  if (m == 0) throw std::invalid_argument("empty pattern");
  if (m > 32 * kMaxWords) throw std::invalid_argument("pattern longer than 4096 characters");
  EncodedPatterns e;
  e.n_patterns = n_patterns;
  e.m = (int)m;
  e.rc = rc_;
  e.bytes.resize(e.n_queries() * m);
  for (size_t q = 0; q < n_patterns; q++) {
    validate_pattern(patterns[q], m);
    memcpy(&e.bytes[q * m], patterns[q], m);
    if (rc_) {
      uint8_t* dst = &e.bytes[(n_patterns + q) * m];
      for (size_t i = 0; i < m; i++) dst[i] = complement_byte(kIupac, patterns[q][m - 1 - i]);
    }
  }
  return e;
}

// Reference: search_with_options -> search_ranges + trace_batch_ranges
// (src/pattern_tiling/general.rs:369-404, trace.rs:262-452): every query is
// searched forward; query index >= n_patterns means strand Rc with
// pattern_idx = idx % n_patterns (trace.rs:444-449); coordinates and CIGAR are
// in the direction of the searched query.
std::vector<Match> Searcher::search_encoded(const EncodedPatterns& enc, const DeviceText& text, size_t k,
                                            bool all_minima) {
  search_encoded_raw(enc, text, k, all_minima);
  return convert_v2(ms_, enc.n_patterns, enc.m);
}

void Searcher::search_encoded_raw(const EncodedPatterns& enc, const uint8_t* text, size_t n, size_t k,
                                  bool all_minima) {
  DeviceText* t = engine_->stage_text(text, n);
  search_encoded_raw(enc, *t, k, all_minima);
}

void Searcher::search_encoded_raw(const EncodedPatterns& enc, const DeviceText& text, size_t k, bool all_minima) {
  std::vector<Query> qs(enc.n_queries());
  for (size_t q = 0; q < qs.size(); q++) qs[q] = Query{&enc.bytes[q * enc.m], false};
  const int kk = (int)std::min<size_t>(k, 1u << 20);
  // the v2 engine knows all_minima, max_n_frac (traced filter, general.rs:399-402) and the overhang
  // (PatterntilingSearcher::new(alpha)); without_trace / only_best_match do not reach it
  // (src/search.rs:415-433).  With overhang the reference pins v2 to the forward v1 search of every
  // query (fuzz_against_sassy_batch, src/pattern_tiling/search.rs:690-848,886-896): same options.
  SearchOpts o;
  o.all_minima = all_minima;
  o.max_n_frac = max_n_frac_;
  o.alpha = alpha_;
  o.max_overhang = max_overhang_;
  o.n_endpoint = alpha_ >= 0.f;
  engine_->search(text, qs, enc.m, kk, o, ms_);
}

std::vector<Match> Searcher::convert_v2(const MatchSet& ms, size_t n_patterns, int m) const {
  std::vector<Match> out(ms.m.size());
  for (size_t i = 0; i < ms.m.size(); i++) {
    const GpuMatch& g = ms.m[i];
    if (g.failed & 1u) throw std::runtime_error("Trace failed (text contains bytes outside the profile's alphabet?)");
    Match& mm = out[i];
    mm.pattern_idx = g.qs % n_patterns;
    mm.strand = g.qs >= n_patterns ? kRc : kFwd;
    mm.text_start = g.text_start;
    mm.text_end = g.text_end;
    mm.pattern_start = (uint64_t)((g.failed >> 8) & 0xFFFu);  // overhang (scan_core.cuh pack_overhang), else 0
    mm.pattern_end = (uint64_t)m - (uint64_t)(g.failed >> 20);
    mm.cost = g.cost;
    unpack_ops(ms, i, mm.ops);
  }
  return out;
}

// Copies rank r's records of the last exchange into a MatchSet.
static MatchSet slot_set(const PeerGather& pg, int r) {
  const PeerGather::Slot sl = pg.slot(r);
  MatchSet ms;
  ms.ops_words = sl.ops_words;
  ms.m.assign(sl.records, sl.records + sl.count);
  ms.ops.assign(sl.ops, sl.ops + sl.count * sl.ops_words);
  return ms;
}

std::vector<Match> Searcher::search_gathered(PeerGather& pg, const uint8_t* pattern, size_t m,
                                             const DeviceText& text, size_t k, bool all_minima, bool* complete) {
  engine_->set_gather(&pg, 1);
  std::vector<Match> local;
  try {
    local = search_with_pam(pattern, m, text, k, all_minima, nullptr, 0);
  } catch (...) {
    engine_->set_gather(nullptr, 0);
    throw;
  }
  engine_->set_gather(nullptr, 0);
  if (pg.pipelined()) {
    // the host mirror holds the records of the PREVIOUS search (same pattern length expected)
    if (!engine_->gather_ok())
      throw std::runtime_error("pipelined gather: this rank's result did not fit the exchange; use the lock-step mode");
    return collected_v1(pg, m, complete);
  }
  *complete = engine_->gather_ok();
  if (!*complete) return local;
  return collected_v1(pg, m, complete);
}

// All ranks' records of the collected step as v1 matches (text_idx = source rank).  Pipelined mode:
// *complete = 2 while the pipeline is being primed (nothing collected yet).
std::vector<Match> Searcher::collected_v1(PeerGather& pg, size_t m, bool* complete, int* state) {
  std::vector<Match> out;
  if (state) *state = 1;
  if (pg.pipelined()) {
    if (!pg.has_result()) {
      *complete = true;
      if (state) *state = 2;
      return out;
    }
    if (!pg.ok())
      throw std::runtime_error("pipelined gather: some rank's result did not fit the exchange; use the lock-step mode");
  }
  *complete = true;
  for (int r = 0; r < pg.world(); r++) {
    const uint64_t n = pg.slot(r).text_n;
    std::vector<Match> part = convert_v1(slot_set(pg, r), 1, m, [n](size_t) { return n; });
    for (auto& mm : part) mm.text_idx = (uint64_t)r;
    out.insert(out.end(), std::make_move_iterator(part.begin()), std::make_move_iterator(part.end()));
  }
  return out;
}

void Searcher::flush_sharded(PeerGather& pg, size_t m, bool all_minima, const SlabInfo* slabs, size_t n_slabs,
                             uint64_t n_global, int* state, FlatMatches& merged) {
  engine_->flush_gather(pg);
  *state = 1;
  if (!pg.pipelined() || !pg.has_result()) {
    *state = 2;
    return;
  }
  if (!pg.ok())
    throw std::runtime_error("pipelined gather: some rank's result did not fit the exchange; use the lock-step mode");
  merge_collected(pg, m, all_minima, slabs, n_slabs, n_global, merged);
}

std::vector<Match> Searcher::flush_gathered(PeerGather& pg, size_t m, int* state) {
  bool complete = false;
  engine_->flush_gather(pg);
  return collected_v1(pg, m, &complete, state);
}

std::vector<Match> Searcher::search_sharded_gathered(PeerGather& pg, const uint8_t* pattern, size_t m,
                                                     const DeviceText& window, size_t k, bool all_minima,
                                                     const SlabInfo* slabs, size_t n_slabs, uint64_t n_global,
                                                     bool* complete, FlatMatches& merged) {
  // search_all of the window; in lock-step mode the records of THIS search are in the host mirror
  // afterwards, in pipelined mode those of the previous one
  engine_->set_gather(&pg, 1);
  raw_only_ = true;  // no Match objects for this rank's own records: the merge reads the raw ones
  try {
    search_with_pam(pattern, m, window, k, /*all_minima=*/true, nullptr, 0);
  } catch (...) {
    raw_only_ = false;
    engine_->set_gather(nullptr, 0);
    throw;
  }
  raw_only_ = false;
  engine_->set_gather(nullptr, 0);
  if (pg.pipelined()) {
    if (!engine_->gather_ok())
      throw std::runtime_error("pipelined gather: this rank's result did not fit the exchange; use the lock-step mode");
    *complete = true;
    if (!pg.has_result()) return {};
    if (!pg.ok())
      throw std::runtime_error("pipelined gather: some rank's result did not fit the exchange; use the lock-step mode");
  } else {
    *complete = engine_->gather_ok();
    if (!*complete) {  // the caller's own collective: this rank's search_all matches, window coordinates
      const uint64_t wn = window.n;
      return convert_v1(ms_, 1, m, [wn](size_t) { return wn; });
    }
  }
  const double t0 = HostTimers::on() ? HostTimers::now_us() : 0.0;
  merge_collected(pg, m, all_minima, slabs, n_slabs, n_global, merged);
  if (HostTimers::on()) HostTimers::add(HostTimers::kMerge, HostTimers::now_us() - t0);
  return {};
}

// Ownership filter + local-minima rule on the RAW records of the collected step (scan-direction
// coordinates, 32 bytes each); only the selected records are turned into Matches (CIGAR strings).
// Every rank merges all ranks' search_all records every step: this is the host cost of a sharded
// search, so it avoids per-record allocations.
void Searcher::merge_collected(PeerGather& pg, size_t m, bool all_minima, const SlabInfo* slabs, size_t n_slabs,
                               uint64_t n_global, FlatMatches& merged) {
  struct Item {
    uint64_t key;
    uint32_t cost, rank, idx;
  };
  std::vector<Item> items;
  const size_t nq = rc_ ? 2 : 1;
  const int nr = std::min<int>(pg.world(), (int)n_slabs);
  // Every rank's records are sorted by (strand, scan-direction end) and the slabs own disjoint,
  // increasing ranges: forward records in rank order, reverse-complement records (whose scan runs
  // right to left) in reverse rank order, are sorted as they come -- the sort below only runs if a
  // caller's slabs break that order.
  bool sorted = true;
  for (size_t strand = 0; strand < nq; strand++) {
    for (int rr = 0; rr < nr; rr++) {
      const int r = strand == 0 ? rr : nr - 1 - rr;
      const PeerGather::Slot sl = pg.slot(r);
      const SlabInfo& sb_ = slabs[r];
      const uint64_t wlen = sl.text_n;
      for (unsigned long long i = 0; i < sl.count; i++) {
        const GpuMatch& g = sl.records[i];
        if (g.qs % nq != strand) continue;
        uint64_t pos;
        bool own;
        if (strand == 0) {  // forward: end position in the global text
          pos = sb_.window_off + g.text_end;
          own = (pos > sb_.own_lo && pos <= sb_.own_hi) || (pos == 0 && sb_.own_lo == 0);
        } else {  // reversed window: scan-direction end e' -> forward start of the match
          const uint64_t start = sb_.window_off + (wlen - g.text_end);
          pos = n_global - start;
          own = (start >= sb_.own_lo && start < sb_.own_hi) || (start == n_global && sb_.own_hi == n_global);
        }
        if (!own) continue;
        const uint64_t key = cand_key((uint32_t)strand, pos);
        if (!items.empty() && items.back().key > key) sorted = false;
        items.push_back(Item{key, (uint32_t)g.cost, (uint32_t)r, (uint32_t)i});
      }
    }
  }
  if (!sorted) std::stable_sort(items.begin(), items.end(), [](const Item& a, const Item& b) { return a.key < b.key; });
  std::vector<uint64_t> keys(items.size());
  std::vector<uint32_t> cost(items.size());
  for (size_t i = 0; i < items.size(); i++) keys[i] = items[i].key, cost[i] = items[i].cost;
  // the kept records straight into flat C records (as Searcher::convert_v1 + to_result would)
  merged.m.clear();
  merged.ops.clear();
  for (size_t i = 0; i < items.size(); i++) {
    if (!select_candidate(keys.data(), cost.data(), i, items.size(), all_minima)) continue;
    const int r = (int)items[i].rank;
    const PeerGather::Slot sl = pg.slot(r);
    const GpuMatch& g = sl.records[items[i].idx];
    if (g.failed & 1u) throw std::runtime_error("Trace failed (text contains bytes outside the profile's alphabet?)");
    const SlabInfo& sb_ = slabs[r];
    const uint64_t wn = sl.text_n;
    sassy_gpu_Match o;
    memset(&o, 0, sizeof o);
    o.pattern_idx = 0;
    o.text_idx = 0;
    o.cost = g.cost;
    o.pattern_start = without_trace_ ? ~0ull : (uint64_t)((g.failed >> 8) & 0xFFFu);
    o.pattern_end = m - (uint64_t)(g.failed >> 20);
    if (g.qs % nq == 0) {
      o.strand = (uint8_t)kFwd;
      o.text_start = sb_.window_off + g.text_start;
      o.text_end = sb_.window_off + g.text_end;
    } else {
      o.strand = (uint8_t)kRc;
      o.text_start = sb_.window_off + (wn - g.text_end);
      o.text_end = without_trace_ ? ~0ull : sb_.window_off + (wn - g.text_start);
    }
    o.ops_off = merged.ops.size();
    if (!without_trace_) {
      const uint32_t* w = sl.ops + (size_t)items[i].idx * sl.ops_words;
      o.ops_len = g.nops;
      const size_t off = merged.ops.size();
      merged.ops.resize(off + g.nops);
      for (uint32_t a = 0; a < g.nops; a++) merged.ops[off + a] = kOpChars[(w[a >> 4] >> ((a & 15) * 2)) & 3u];
    }
    merged.m.push_back(o);
  }
}

std::vector<Match> Searcher::merge_gathered(std::vector<Match>& all, bool all_minima, const SlabInfo* slabs,
                                            size_t n_slabs, uint64_t n_global) {
  const std::vector<size_t> keep = merge_slab_matches(all, slabs, n_slabs, n_global, all_minima);
  std::vector<Match> out;
  out.reserve(keep.size());
  for (size_t i : keep) out.push_back(std::move(all[i]));
  return out;
}

std::vector<Match> Searcher::search_encoded_gathered(PeerGather& pg, const EncodedPatterns& enc,
                                                     const DeviceText& text, size_t k, bool all_minima,
                                                     bool* complete) {
  engine_->set_gather(&pg, enc.n_patterns);
  std::vector<Match> local;
  try {
    local = search_encoded(enc, text, k, all_minima);
  } catch (...) {
    engine_->set_gather(nullptr, 0);
    throw;
  }
  engine_->set_gather(nullptr, 0);
  *complete = engine_->gather_ok();
  if (!*complete) return local;
  std::vector<Match> out;
  for (int r = 0; r < pg.world(); r++) {
    std::vector<Match> part = convert_v2(slot_set(pg, r), (size_t)pg.slot(r).user, enc.m);
    for (auto& mm : part) mm.text_idx = (uint64_t)r;
    out.insert(out.end(), std::make_move_iterator(part.begin()), std::make_move_iterator(part.end()));
  }
  return out;
}

std::vector<Match> Searcher::search_encoded(const EncodedPatterns& enc, const uint8_t* text, size_t n, size_t k,
                                            bool all_minima) {
  DeviceText* t = engine_->stage_text(text, n);
  return search_encoded(enc, *t, k, all_minima);
}

}  // namespace sb

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------

struct sassy_SearcherType {
  sb::Searcher s;
  sassy_SearcherType(const std::string& a, bool rc, float alpha, int dev) : s(a, rc, alpha, dev) {}
};
struct sassy_gpu_Text {
  sb::DeviceText* t;
};
struct sassy_gpu_Patterns {
  sb::EncodedPatterns e;
};
struct sassy_gpu_Result {
  std::vector<sassy_gpu_Match> m;
  std::string ops;
};
struct sassy_gpu_Gather {
  sb::PeerGather g;
  sassy_gpu_Gather(int dev, int world, int rank, size_t cap, size_t ops_words) : g(dev, world, rank, cap, ops_words) {}
};

namespace {

thread_local std::string g_last_error;

[[noreturn]] void die(const char* what) {
  // The reference panics across extern "C" (src/c.rs:57,68,74,98), which aborts the process.
  fprintf(stderr, "sassy_b200: %s\n", what);
  abort();
}

int default_device() {
  const char* d = getenv("SASSY_B200_DEVICE");
  return d ? atoi(d) : 0;
}

sassy_gpu_Result* to_result(const std::vector<sb::Match>& v) {
  sassy_gpu_Result* r = new sassy_gpu_Result;
  r->m.resize(v.size());
  size_t total = 0;
  for (auto& x : v) total += x.ops.size();
  r->ops.reserve(total);
  for (size_t i = 0; i < v.size(); i++) {
    sassy_gpu_Match& o = r->m[i];
    memset(&o, 0, sizeof o);
    o.pattern_idx = v[i].pattern_idx;
    o.text_idx = v[i].text_idx;
    o.text_start = v[i].text_start;
    o.text_end = v[i].text_end;
    o.pattern_start = v[i].pattern_start;
    o.pattern_end = v[i].pattern_end;
    o.cost = v[i].cost;
    o.strand = (uint8_t)v[i].strand;
    o.ops_off = r->ops.size();
    o.ops_len = (uint32_t)v[i].ops.size();
    r->ops += v[i].ops;
  }
  return r;
}

sassy_gpu_Result* flat_result(sb::Searcher::FlatMatches& f) {
  sassy_gpu_Result* r = new sassy_gpu_Result;
  r->m.swap(f.m);
  r->ops.swap(f.ops);
  return r;
}

// v2 records -> flat C result in one pass (no per-match heap allocation): same mapping as
// Searcher::convert_v2.
sassy_gpu_Result* to_result_v2(const sb::MatchSet& ms, size_t n_patterns, int m) {
  sassy_gpu_Result* r = new sassy_gpu_Result;
  const size_t n = ms.m.size();
  r->m.resize(n);
  size_t total = 0;
  for (size_t i = 0; i < n; i++) {
    if (ms.m[i].failed & 1u) {
      delete r;
      throw std::runtime_error("Trace failed (text contains bytes outside the profile's alphabet?)");
    }
    total += ms.m[i].nops;
  }
  r->ops.resize(total);
  char* ops = total ? &r->ops[0] : nullptr;
  size_t off = 0;
  for (size_t i = 0; i < n; i++) {
    const sb::GpuMatch& g = ms.m[i];
    sassy_gpu_Match& o = r->m[i];
    memset(&o, 0, sizeof o);
    o.pattern_idx = g.qs % n_patterns;
    o.strand = g.qs >= n_patterns ? 1 : 0;
    o.text_start = g.text_start;
    o.text_end = g.text_end;
    o.pattern_start = (uint64_t)((g.failed >> 8) & 0xFFFu);
    o.pattern_end = (uint64_t)m - (uint64_t)(g.failed >> 20);
    o.cost = g.cost;
    o.ops_off = off;
    o.ops_len = g.nops;
    const uint32_t* w = &ms.ops[i * ms.ops_words];
    for (uint32_t a = 0; a < g.nops; a++) ops[off + a] = "=XID"[(w[a >> 4] >> ((a & 15) * 2)) & 3u];
    off += g.nops;
  }
  return r;
}

template <class F>
auto guarded(F&& f) -> decltype(f()) {
  try {
    g_last_error.clear();
    return f();
  } catch (const std::exception& e) {
    g_last_error = e.what();
  } catch (...) {
    g_last_error = "unknown error";
  }
  return decltype(f())();
}

}  // namespace

extern "C" {

sassy_SearcherType* sassy_searcher(const char* alphabet, bool rc, float alpha) {
  if (!alphabet) die("Alphabet pointer must not be null");
  try {
    return new sassy_SearcherType(alphabet, rc, alpha, default_device());
  } catch (const std::exception& e) {
    die(e.what());
  }
}

void sassy_searcher_free(sassy_SearcherType* ptr) {
  if (!ptr) die("Pointer to SearcherType must not be null");
  delete ptr;
}

uintptr_t search(sassy_SearcherType* searcher, const uint8_t* pattern, uintptr_t pattern_len, const uint8_t* text,
                 uintptr_t text_len, uintptr_t k, sassy_Match** out_matches) {
  if (!searcher || !pattern || !text || !out_matches) die("Pointers in search() must not be null");
  std::vector<sb::Match> v;
  try {
    g_last_error.clear();
    v = searcher->s.search(pattern, pattern_len, text, text_len, k, /*all_minima=*/false);
  } catch (const sb::CapacityError& e) {
    // Not one of the reference's panics: a limit of this implementation (pattern longer than
    // 1024 characters, text of 2^40 bytes or more).  The process is not aborted: no matches are
    // returned, the message goes to stderr and to sassy_gpu_last_error() (include/sassy.h).
    g_last_error = e.what();
    fprintf(stderr, "sassy_b200: search() not run: %s\n", e.what());
    v.clear();
  } catch (const std::exception& e) {
    die(e.what());
  }
  // len == capacity; non-null even for zero matches (reference src/c.rs:112-117,127).
  sassy_Match* arr = static_cast<sassy_Match*>(malloc(std::max<size_t>(v.size(), 1) * sizeof(sassy_Match)));
  if (!arr) die("out of memory");
  for (size_t i = 0; i < v.size(); i++) {
    memset(&arr[i], 0, sizeof(sassy_Match));
    arr[i].text_start = v[i].text_start;
    arr[i].text_end = v[i].text_end;
    arr[i].pattern_start = v[i].pattern_start;
    arr[i].pattern_end = v[i].pattern_end;
    arr[i].cost = v[i].cost;
    arr[i].strand = (uint8_t)v[i].strand;
  }
  *out_matches = arr;
  return v.size();
}

void sassy_matches_free(sassy_Match* ptr, uintptr_t len) {
  (void)len;
  if (!ptr) die("Pointer to matches must not be null");
  free(ptr);
}

// ---- extensions (include/sassy_gpu.h) --------------------------------------

int sassy_gpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

const char* sassy_gpu_last_error(void) { return g_last_error.c_str(); }

int sassy_gpu_device_info(int device, char* name, size_t name_cap, int* sm_count, int* sm_clock_mhz,
                          size_t* total_mem, int* cc_major, int* cc_minor) {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    cudaGetLastError();
    g_last_error = "no such CUDA device";
    return 1;
  }
  if (name && name_cap) {
    strncpy(name, prop.name, name_cap - 1);
    name[name_cap - 1] = 0;
  }
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (sm_clock_mhz) *sm_clock_mhz = khz / 1000;
  if (total_mem) *total_mem = prop.totalGlobalMem;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return 0;
}

sassy_SearcherType* sassy_gpu_searcher(const char* alphabet, bool rc, float alpha, int device) {
  return guarded([&]() -> sassy_SearcherType* {
    if (!alphabet) throw std::invalid_argument("Alphabet pointer must not be null");
    return new sassy_SearcherType(alphabet, rc, alpha, device);
  });
}

int sassy_gpu_set_variant(sassy_SearcherType* searcher, int variant) {
  if (!searcher || (variant != sb::kVariantTma && variant != sb::kVariantLdg)) return 1;
  searcher->s.engine().set_variant(variant);
  return 0;
}

int sassy_gpu_set_filter(sassy_SearcherType* searcher, int mode) {
  if (!searcher || mode < 0 || mode > 2) return 1;
  searcher->s.engine().set_filter_mode(mode);
  return 0;
}

int sassy_gpu_set_transport(sassy_SearcherType* searcher, int mode) {
  if (!searcher || mode < 0 || mode > 1) return 1;
  searcher->s.engine().set_transport(mode);
  return 0;
}

int sassy_gpu_stats(const sassy_SearcherType* searcher, sassy_gpu_Stats* out) {
  if (!searcher || !out) return 1;
  const sb::SearchStats& st = const_cast<sassy_SearcherType*>(searcher)->s.engine().stats();
  memset(out, 0, sizeof *out);
  out->scan_ms = st.scan_ms;
  out->total_ms = st.total_ms;
  out->scan_launches = st.scan_launches;
  out->aux_launches = st.aux_launches;
  out->candidates = st.candidates;
  out->matches = st.matches;
  out->row_bytes = st.ltot;
  out->rows = st.rows;
  out->words = st.words;
  out->blocks_per_sm = st.blocks_per_sm;
  out->retries = st.retries;
  out->filter_words = st.filter_words;
  out->filter_ms = st.filter_ms;
  out->verify_ms = st.verify_ms;
  out->hits = st.hits;
  out->filter_len = st.filter_len;
  out->filter_fallback = st.filter_fallback;
  out->transfer_ms = st.transfer_ms;
  out->transfer_packed = st.transfer_packed;
  out->transfer_bytes = st.transfer_bytes;
  out->filter_kind = st.filter_kind;
  out->swar_lanes = st.swar_lanes;
  out->confirmed = st.confirmed;
  out->dense_tiles = st.dense_tiles;
  out->reserved3 = 0;
  return 0;
}

void* sassy_gpu_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
    g_last_error = "cudaHostAlloc failed";
    return nullptr;
  }
  return p;
}

void sassy_gpu_host_free(void* ptr) {
  if (ptr) cudaFreeHost(ptr);
}

sassy_gpu_Text* sassy_gpu_text_upload(sassy_SearcherType* searcher, const uint8_t* text, size_t text_len) {
  return guarded([&]() -> sassy_gpu_Text* {
    if (!searcher || (!text && text_len)) throw std::invalid_argument("null pointer");
    return new sassy_gpu_Text{searcher->s.engine().upload_text(text, text_len)};
  });
}

sassy_gpu_Text* sassy_gpu_text_from_device(sassy_SearcherType* searcher, const void* device_ptr, size_t text_len) {
  return guarded([&]() -> sassy_gpu_Text* {
    if (!searcher || (!device_ptr && text_len)) throw std::invalid_argument("null pointer");
    return new sassy_gpu_Text{searcher->s.engine().adopt_device_text(device_ptr, text_len)};
  });
}

size_t sassy_gpu_text_len(const sassy_gpu_Text* text) { return text ? (size_t)text->t->n : 0; }

void sassy_gpu_text_free(sassy_SearcherType* searcher, sassy_gpu_Text* text) {
  if (!searcher || !text) return;
  searcher->s.engine().free_text(text->t);
  delete text;
}

sassy_gpu_Result* sassy_gpu_search(sassy_SearcherType* searcher, const uint8_t* pattern, size_t pattern_len,
                                   const uint8_t* text, size_t text_len, size_t k, int all) {
  return guarded([&]() -> sassy_gpu_Result* {
    if (!searcher || !pattern || (!text && text_len)) throw std::invalid_argument("null pointer");
    sb::Searcher::FlatMatches f;
    searcher->s.search_flat(pattern, pattern_len, text, text_len, k, all != 0, nullptr, 0, f);
    return flat_result(f);
  });
}

sassy_gpu_Result* sassy_gpu_search_text(sassy_SearcherType* searcher, const uint8_t* pattern, size_t pattern_len,
                                        const sassy_gpu_Text* text, size_t k, int all) {
  return guarded([&]() -> sassy_gpu_Result* {
    if (!searcher || !pattern || !text) throw std::invalid_argument("null pointer");
    sb::Searcher::FlatMatches f;
    searcher->s.search_flat(pattern, pattern_len, *text->t, k, all != 0, nullptr, 0, f);
    return flat_result(f);
  });
}

int sassy_gpu_set_trace(sassy_SearcherType* searcher, int trace) {
  if (!searcher) return 1;
  searcher->s.set_trace(trace != 0);
  return 0;
}

int sassy_gpu_set_only_best_match(sassy_SearcherType* searcher, int on) {
  if (!searcher) return 1;
  searcher->s.set_only_best_match(on != 0);
  return 0;
}

int sassy_gpu_set_max_overhang(sassy_SearcherType* searcher, int max_overhang) {
  if (!searcher) return 1;
  searcher->s.set_max_overhang(max_overhang);
  return 0;
}

int sassy_gpu_set_max_n_frac(sassy_SearcherType* searcher, float max_n_frac) {
  if (!searcher || !(max_n_frac >= 0.f)) return 1;
  searcher->s.set_max_n_frac(max_n_frac);
  return 0;
}

sassy_gpu_Result* sassy_gpu_search_pam(sassy_SearcherType* searcher, const uint8_t* pattern, size_t pattern_len,
                                       const uint8_t* text, size_t text_len, size_t k, int all, const uint8_t* pam,
                                       size_t pam_len) {
  return guarded([&]() -> sassy_gpu_Result* {
    if (!searcher || !pattern || (!text && text_len) || (!pam && pam_len)) throw std::invalid_argument("null pointer");
    sb::Searcher::FlatMatches f;
    searcher->s.search_flat(pattern, pattern_len, text, text_len, k, all != 0, pam, pam_len, f);
    return flat_result(f);
  });
}

sassy_gpu_Result* sassy_gpu_search_pam_text(sassy_SearcherType* searcher, const uint8_t* pattern, size_t pattern_len,
                                            const sassy_gpu_Text* text, size_t k, int all, const uint8_t* pam,
                                            size_t pam_len) {
  return guarded([&]() -> sassy_gpu_Result* {
    if (!searcher || !pattern || !text || (!pam && pam_len)) throw std::invalid_argument("null pointer");
    sb::Searcher::FlatMatches f;
    searcher->s.search_flat(pattern, pattern_len, *text->t, k, all != 0, pam, pam_len, f);
    return flat_result(f);
  });
}

sassy_gpu_Result* sassy_gpu_search_patterns(sassy_SearcherType* searcher, const uint8_t* const* patterns,
                                            size_t n_patterns, size_t pattern_len, const uint8_t* text,
                                            size_t text_len, size_t k) {
  return guarded([&]() -> sassy_gpu_Result* {
    if (!searcher || (!patterns && n_patterns) || (!text && text_len)) throw std::invalid_argument("null pointer");
    return to_result(searcher->s.search_patterns(patterns, n_patterns, pattern_len, text, text_len, k));
  });
}

sassy_gpu_Result* sassy_gpu_search_texts(sassy_SearcherType* searcher, const uint8_t* pattern, size_t pattern_len,
                                         const uint8_t* const* texts, const size_t* text_lens, size_t n_texts,
                                         size_t k) {
  return guarded([&]() -> sassy_gpu_Result* {
    if (!searcher || !pattern || ((!texts || !text_lens) && n_texts)) throw std::invalid_argument("null pointer");
    return to_result(searcher->s.search_texts(pattern, pattern_len, texts,
                                              reinterpret_cast<const uint64_t*>(text_lens), n_texts, k));
  });
}

sassy_gpu_Result* sassy_gpu_search_many(sassy_SearcherType* searcher, const uint8_t* const* patterns,
                                        const size_t* pattern_lens, size_t n_patterns, const uint8_t* const* texts,
                                        const size_t* text_lens, size_t n_texts, size_t k) {
  return guarded([&]() -> sassy_gpu_Result* {
    if (!searcher || ((!patterns || !pattern_lens) && n_patterns) || ((!texts || !text_lens) && n_texts))
      throw std::invalid_argument("null pointer");
    return to_result(searcher->s.search_many(patterns, reinterpret_cast<const uint64_t*>(pattern_lens), n_patterns,
                                             texts, reinterpret_cast<const uint64_t*>(text_lens), n_texts, k));
  });
}

sassy_gpu_Patterns* sassy_gpu_encode_patterns(sassy_SearcherType* searcher, const uint8_t* patterns,
                                              size_t n_patterns, size_t pattern_len) {
  return guarded([&]() -> sassy_gpu_Patterns* {
    if (!searcher || (!patterns && n_patterns)) throw std::invalid_argument("null pointer");
    std::vector<const uint8_t*> ptrs(n_patterns);
    for (size_t i = 0; i < n_patterns; i++) ptrs[i] = patterns + i * pattern_len;
    return new sassy_gpu_Patterns{searcher->s.encode_patterns(ptrs.data(), n_patterns, pattern_len)};
  });
}

void sassy_gpu_patterns_free(sassy_gpu_Patterns* patterns) { delete patterns; }

sassy_gpu_Result* sassy_gpu_search_encoded(sassy_SearcherType* searcher, const sassy_gpu_Patterns* patterns,
                                           const sassy_gpu_Text* text, size_t k, int all) {
  return guarded([&]() -> sassy_gpu_Result* {
    if (!searcher || !patterns || !text) throw std::invalid_argument("null pointer");
    searcher->s.search_encoded_raw(patterns->e, *text->t, k, all != 0);
    return to_result_v2(searcher->s.last_set(), patterns->e.n_patterns, patterns->e.m);
  });
}

sassy_gpu_Result* sassy_gpu_search_encoded_host(sassy_SearcherType* searcher, const sassy_gpu_Patterns* patterns,
                                                const uint8_t* text, size_t text_len, size_t k, int all) {
  return guarded([&]() -> sassy_gpu_Result* {
    if (!searcher || !patterns || (!text && text_len)) throw std::invalid_argument("null pointer");
    searcher->s.search_encoded_raw(patterns->e, text, text_len, k, all != 0);
    return to_result_v2(searcher->s.last_set(), patterns->e.n_patterns, patterns->e.m);
  });
}

sassy_gpu_Gather* sassy_gpu_gather_create(sassy_SearcherType* searcher, int world, int rank, size_t cap_records,
                                          size_t max_ops) {
  return guarded([&]() -> sassy_gpu_Gather* {
    if (!searcher) throw std::invalid_argument("null pointer");
    if (cap_records < (size_t)sb::kSmallCandidates) cap_records = sb::kSmallCandidates;
    return new sassy_gpu_Gather(searcher->s.engine().device(), world, rank, cap_records, (max_ops + 15) / 16);
  });
}

int sassy_gpu_gather_handle(sassy_gpu_Gather* gather, uint8_t* out64) {
  return guarded([&]() -> int {
           if (!gather || !out64) throw std::invalid_argument("null pointer");
           gather->g.export_handle(out64);
           return 1;
         }) == 1 ? 0 : 1;
}

int sassy_gpu_gather_connect(sassy_gpu_Gather* gather, const uint8_t* handles) {
  return guarded([&]() -> int {
           if (!gather || !handles) throw std::invalid_argument("null pointer");
           gather->g.connect(handles);
           return 1;
         }) == 1 ? 0 : 1;
}

void sassy_gpu_gather_free(sassy_gpu_Gather* gather) { delete gather; }

sassy_gpu_Result* sassy_gpu_search_text_gathered(sassy_SearcherType* searcher, sassy_gpu_Gather* gather,
                                                 const uint8_t* pattern, size_t pattern_len,
                                                 const sassy_gpu_Text* text, size_t k, int all, int* complete) {
  return guarded([&]() -> sassy_gpu_Result* {
    if (!searcher || !gather || !pattern || !text || !complete) throw std::invalid_argument("null pointer");
    bool ok = false;
    auto v = searcher->s.search_gathered(gather->g, pattern, pattern_len, *text->t, k, all != 0, &ok);
    *complete = ok ? 1 : 0;
    return to_result(v);
  });
}

sassy_gpu_Result* sassy_gpu_search_encoded_gathered(sassy_SearcherType* searcher, sassy_gpu_Gather* gather,
                                                    const sassy_gpu_Patterns* patterns, const sassy_gpu_Text* text,
                                                    size_t k, int all, int* complete) {
  return guarded([&]() -> sassy_gpu_Result* {
    if (!searcher || !gather || !patterns || !text || !complete) throw std::invalid_argument("null pointer");
    bool ok = false;
    auto v = searcher->s.search_encoded_gathered(gather->g, patterns->e, *text->t, k, all != 0, &ok);
    *complete = ok ? 1 : 0;
    return to_result(v);
  });
}

sassy_gpu_Result* sassy_gpu_search_text_sharded(sassy_SearcherType* searcher, sassy_gpu_Gather* gather,
                                                const uint8_t* pattern, size_t pattern_len,
                                                const sassy_gpu_Text* window, size_t k, int all,
                                                const sassy_gpu_Slab* slabs, size_t n_slabs, uint64_t n_global,
                                                int* complete) {
  return guarded([&]() -> sassy_gpu_Result* {
    if (!searcher || !gather || !pattern || !window || !slabs || !complete) throw std::invalid_argument("null pointer");
    std::vector<sb::SlabInfo> info(n_slabs);
    for (size_t i = 0; i < n_slabs; i++) info[i] = sb::SlabInfo{slabs[i].window_off, slabs[i].own_lo, slabs[i].own_hi};
    bool ok = false;
    const double t_call = sb::HostTimers::on() ? sb::HostTimers::now_us() : 0.0;
    struct CallTimer {
      double t0;
      ~CallTimer() {
        if (sb::HostTimers::on()) sb::HostTimers::add(sb::HostTimers::kSearchCall, sb::HostTimers::now_us() - t0);
      }
    } call_timer{t_call};
    sb::Searcher::FlatMatches merged;
    auto v = searcher->s.search_sharded_gathered(gather->g, pattern, pattern_len, *window->t, k, all != 0, info.data(),
                                                 n_slabs, n_global, &ok, merged);
    *complete = ok ? 1 : 0;
    if (gather->g.pipelined() && !gather->g.has_result()) *complete = 2;  // pipeline priming: no result yet
    if (!ok) return to_result(v);  // this rank's unmerged matches for the caller's own collective
    sassy_gpu_Result* r = new sassy_gpu_Result;
    r->m.swap(merged.m);
    r->ops.swap(merged.ops);
    return r;
  });
}

int sassy_gpu_gather_set_pipelined(sassy_gpu_Gather* gather, int on) {
  if (!gather) return 1;
  gather->g.set_pipelined(on != 0);
  return 0;
}

int sassy_gpu_gather_has_result(const sassy_gpu_Gather* gather) { return gather && gather->g.has_result() ? 1 : 0; }

sassy_gpu_Result* sassy_gpu_text_sharded_flush(sassy_SearcherType* searcher, sassy_gpu_Gather* gather,
                                               size_t pattern_len, int all, const sassy_gpu_Slab* slabs,
                                               size_t n_slabs, uint64_t n_global, int* state) {
  return guarded([&]() -> sassy_gpu_Result* {
    if (!searcher || !gather || !slabs || !state) throw std::invalid_argument("null pointer");
    std::vector<sb::SlabInfo> info(n_slabs);
    for (size_t i = 0; i < n_slabs; i++) info[i] = sb::SlabInfo{slabs[i].window_off, slabs[i].own_lo, slabs[i].own_hi};
    sb::Searcher::FlatMatches merged;
    searcher->s.flush_sharded(gather->g, pattern_len, all != 0, info.data(), n_slabs, n_global, state, merged);
    sassy_gpu_Result* r = new sassy_gpu_Result;
    r->m.swap(merged.m);
    r->ops.swap(merged.ops);
    return r;
  });
}

sassy_gpu_Result* sassy_gpu_merge_slabs(const sassy_gpu_Match* records, size_t n_records, const char* ops,
                                        const sassy_gpu_Slab* slabs, size_t n_slabs, uint64_t n_global, int all) {
  return guarded([&]() -> sassy_gpu_Result* {
    if ((n_records && !records) || !slabs) throw std::invalid_argument("null pointer");
    std::vector<sassy_gpu_Match> recs(records, records + n_records);
    std::vector<sb::SlabInfo> info(n_slabs);
    for (size_t i = 0; i < n_slabs; i++) info[i] = sb::SlabInfo{slabs[i].window_off, slabs[i].own_lo, slabs[i].own_hi};
    const std::vector<size_t> keep = sb::merge_slab_matches(recs, info.data(), n_slabs, n_global, all != 0);
    sassy_gpu_Result* r = new sassy_gpu_Result;
    r->m.reserve(keep.size());
    for (size_t i : keep) {
      sassy_gpu_Match o = recs[i];
      const uint64_t off = o.ops_off;
      o.ops_off = r->ops.size();
      if (o.ops_len && ops) r->ops.append(ops + off, o.ops_len);
      r->m.push_back(o);
    }
    return r;
  });
}

size_t sassy_gpu_result_len(const sassy_gpu_Result* result) { return result ? result->m.size() : 0; }
const sassy_gpu_Match* sassy_gpu_result_matches(const sassy_gpu_Result* result) {
  return result ? result->m.data() : nullptr;
}
const char* sassy_gpu_result_ops(const sassy_gpu_Result* result) { return result ? result->ops.data() : nullptr; }

size_t sassy_gpu_cigar(const sassy_gpu_Result* result, size_t i, char* buf, size_t cap) {
  if (!result || i >= result->m.size()) return 0;
  sb::Match tmp;
  tmp.ops = result->ops.substr(result->m[i].ops_off, result->m[i].ops_len);
  const std::string c = tmp.cigar();
  if (buf && cap) {
    const size_t nn = std::min(c.size(), cap - 1);
    memcpy(buf, c.data(), nn);
    buf[nn] = 0;
  }
  return c.size();
}

void sassy_gpu_result_free(sassy_gpu_Result* result) { delete result; }

}  // extern "C"
