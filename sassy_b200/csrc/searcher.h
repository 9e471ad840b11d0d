// C++ mirror of the reference's public search surface for this path
// (reference src/search.rs: Searcher<P> :227-256, new/new_fwd/new_rc :364-371,
//  search :510, search_all :685, encode_patterns :404, search_encoded_patterns
//  :415, search_all_encoded_patterns :426; Match :35-62; Strand :116-119).
// The reference's generic parameter P becomes a runtime profile id; everything
// below L4 of the reference runs in CUDA (engine.h).
#pragma once
#include <stdint.h>

#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "engine.h"
#include "shard_merge.h"

namespace sb {

enum Strand : uint8_t { kFwd = 0, kRc = 1 };

struct Match {
  uint64_t pattern_idx = 0;
  uint64_t text_idx = 0;
  uint64_t text_start = 0;
  uint64_t text_end = 0;
  uint64_t pattern_start = 0;
  uint64_t pattern_end = 0;
  int32_t cost = 0;
  Strand strand = kFwd;
  std::string ops;  // one char per op ('=', 'X', 'I', 'D') in pattern direction

  std::string cigar() const;  // run-length "<cnt><op>" (pa_types::Cigar::to_string)
};

struct InvalidPattern : std::invalid_argument {
  using std::invalid_argument::invalid_argument;
};

// Reference: EncodedPatterns (src/pattern_tiling/general.rs:133-150).
struct EncodedPatterns {
  size_t n_patterns = 0;  // originals
  int m = 0;
  bool rc = false;
  std::vector<uint8_t> bytes;  // n_queries * m: originals, then reverse complements if rc
  size_t n_queries() const { return rc ? 2 * n_patterns : n_patterns; }
};

class Searcher {
 public:
  // alphabet: "ascii" | "dna" | "iupac" (case-insensitive), reference src/c.rs:62-67.
  Searcher(const std::string& alphabet, bool rc, float alpha, int device);

  int profile() const { return profile_; }
  bool rc() const { return rc_; }
  Engine& engine() { return *engine_; }

  // Searcher::search / search_all on a host text (copied to HBM for the call).
  std::vector<Match> search(const uint8_t* pattern, size_t m, const uint8_t* text, size_t n, size_t k,
                            bool all_minima);
  // Same on a text already resident in HBM.
  std::vector<Match> search(const uint8_t* pattern, size_t m, const DeviceText& text, size_t k, bool all_minima);

  EncodedPatterns encode_patterns(const uint8_t* const* patterns, size_t n_patterns, size_t m) const;
  std::vector<Match> search_encoded(const EncodedPatterns& enc, const DeviceText& text, size_t k, bool all_minima);
  std::vector<Match> search_encoded(const EncodedPatterns& enc, const uint8_t* text, size_t n, size_t k,
                                    bool all_minima);
  // Same search, result left in device-record form (last_set()): batch searches return matches by
  // the million, and the C ABI turns the records into its flat result without a Match per record.
  void search_encoded_raw(const EncodedPatterns& enc, const DeviceText& text, size_t k, bool all_minima);
  void search_encoded_raw(const EncodedPatterns& enc, const uint8_t* text, size_t n, size_t k, bool all_minima);
  const MatchSet& last_set() const { return ms_; }

  // ---- Searcher options (reference src/search.rs:441-483) ------------------------------------
  void set_trace(bool trace) { without_trace_ = !trace; }          // with_trace / without_trace / set_trace
  void set_only_best_match(bool on) { only_best_ = on; }          // only_best_match
  void set_max_n_frac(float f) { max_n_frac_ = (f == 1.0f) ? -1.f : f; }  // set_max_n_frac; 1.0 disables
  void set_max_overhang(int max_overhang) { max_overhang_ = max_overhang < 0 ? -1 : max_overhang; }  // with_max_overhang
  bool without_trace() const { return without_trace_; }

  // search_with_fn (src/search.rs:767-784) with the end filter the reference ships
  // (bin/crispr.rs:198-205): keep an end position only if the pam_len text characters before
  // it match `pam` exactly (profile is_match; complemented PAM on the reverse strand).
  std::vector<Match> search_with_pam(const uint8_t* pattern, size_t m, const DeviceText& text, size_t k,
                                     bool all_minima, const uint8_t* pam, size_t pam_len);
  std::vector<Match> search_with_pam(const uint8_t* pattern, size_t m, const uint8_t* text, size_t n, size_t k,
                                     bool all_minima, const uint8_t* pam, size_t pam_len);
  // The same searches with the result written as flat C records (one pass over the device records,
  // no per-match allocation: result sets of 10^5..10^6 matches are bound by this conversion).
  struct FlatMatches {
    std::vector<sassy_gpu_Match> m;
    std::string ops;
  };
  void search_flat(const uint8_t* pattern, size_t m, const DeviceText& text, size_t k, bool all_minima,
                   const uint8_t* pam, size_t pam_len, FlatMatches& out);
  void search_flat(const uint8_t* pattern, size_t m, const uint8_t* text, size_t n, size_t k, bool all_minima,
                   const uint8_t* pam, size_t pam_len, FlatMatches& out);

  // search_patterns (src/search.rs:648-683): equal-length patterns against one text (v1
  // semantics per pattern, pattern_idx set).
  std::vector<Match> search_patterns(const uint8_t* const* patterns, size_t n_patterns, size_t m,
                                     const uint8_t* text, size_t n, size_t k);
  // search_texts (src/search.rs:615-640): one pattern against many texts (text_idx set).
  std::vector<Match> search_texts(const uint8_t* pattern, size_t m, const uint8_t* const* texts,
                                  const uint64_t* text_lens, size_t n_texts, size_t k);
  // search_many (src/search.rs:531-603): every pattern against every text; the three modes of
  // the reference return the same set, ordered here like SearchMode::Single (pattern-major).
  std::vector<Match> search_many(const uint8_t* const* patterns, const uint64_t* pattern_lens, size_t n_patterns,
                                 const uint8_t* const* texts, const uint64_t* text_lens, size_t n_texts, size_t k);

  // ---- multi-GPU: search + gather of all ranks' matches over peer memory --------------------
  // Every rank of `pg` must make the same sequence of gathered calls.  *complete = true: the
  // returned matches are those of ALL ranks in rank order, text_idx = source rank (pattern_idx
  // local to the source rank's pattern set).  *complete = false (some rank's result did not fit
  // the fused exchange): only this rank's matches are returned and the caller runs its own
  // collective (sassy_b200/dist.py: NCCL all-gather).
  std::vector<Match> search_gathered(PeerGather& pg, const uint8_t* pattern, size_t m, const DeviceText& text,
                                     size_t k, bool all_minima, bool* complete);
  std::vector<Match> search_encoded_gathered(PeerGather& pg, const EncodedPatterns& enc, const DeviceText& text,
                                             size_t k, bool all_minima, bool* complete);
  // ONE text cut into world slabs (csrc/shard_merge.h): `window` = this rank's slab plus (m + k)
  // halos; every rank searches its window with search_all, the records travel through the fused
  // gather and the ownership filter + local-minima rule run on the merged list.  *complete = false:
  // this rank's unmerged search_all matches (window coordinates) are returned for the caller's own
  // collective + sassy_gpu_merge_slabs.
  // The merged result is written as flat C records (`merged`; no per-match allocation: every rank
  // merges all ranks' records every step); the returned vector only holds the fallback's matches.
  std::vector<Match> search_sharded_gathered(PeerGather& pg, const uint8_t* pattern, size_t m,
                                             const DeviceText& window, size_t k, bool all_minima,
                                             const SlabInfo* slabs, size_t n_slabs, uint64_t n_global,
                                             bool* complete, FlatMatches& merged);

  // Pipelined PeerGather: the records of the collected (previous) step / of the last pushed step.
  std::vector<Match> collected_v1(PeerGather& pg, size_t m, bool* complete, int* state = nullptr);
  std::vector<Match> flush_gathered(PeerGather& pg, size_t m, int* state);
  std::vector<Match> merge_gathered(std::vector<Match>& all, bool all_minima, const SlabInfo* slabs, size_t n_slabs,
                                    uint64_t n_global);
  void merge_collected(PeerGather& pg, size_t m, bool all_minima, const SlabInfo* slabs, size_t n_slabs,
                       uint64_t n_global, FlatMatches& merged);
  void flush_sharded(PeerGather& pg, size_t m, bool all_minima, const SlabInfo* slabs, size_t n_slabs,
                     uint64_t n_global, int* state, FlatMatches& merged);

  void validate_pattern(const uint8_t* p, size_t m) const;

 private:
  std::vector<Match> convert_v2(const MatchSet& ms, size_t n_patterns, int m) const;
  SearchOpts v1_opts(bool all_minima) const;
  // Converts slot-indexed device records of a v1 search over `n_pat` patterns (queries =
  // patterns, then their complements when rc) to Matches; text_len(text_idx) gives the text length.
  template <class LenFn>
  std::vector<Match> convert_v1(const MatchSet& ms, size_t n_pat, size_t m, LenFn text_len) const;
  std::vector<Match> search_group(const uint8_t* const* patterns, size_t n_pat, size_t m,
                                  const uint8_t* const* texts, const uint64_t* text_lens, size_t n_texts, size_t k);
  int profile_;
  bool rc_;
  bool without_trace_ = false;
  bool only_best_ = false;
  float max_n_frac_ = -1.f;
  float alpha_ = -1.f;      // overhang cost (Iupac only), < 0 = off
  int max_overhang_ = -1;
  std::unique_ptr<Engine> engine_;
  MatchSet ms_;
  bool raw_only_ = false;  // search_with_pam leaves the records in ms_ without building Matches
};

int parse_alphabet(const std::string& alphabet);  // -1 if unknown

}  // namespace sb
