// Launch interface between the host engine (engine.cu) and the CUDA kernels
// (scan_kernels.cu, post_kernels.cu).  sm_100a only.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "host_logic.h"

namespace sb {

#ifndef SB_STAGES
#define SB_STAGES 2
#endif
constexpr int kScanStages = SB_STAGES;  // measured: 2 x 64 B beats deeper rings (profiles/r01_pipeline_variants.md)  // per-warp shared-memory ring depth (64 B per thread per stage)

enum ScanVariant : int { kVariantTma = 0, kVariantLdg = 1 };


size_t scan_smem_bytes(int W, int variant, uint32_t nrows);
// Max resident blocks per SM for the given instantiation (for grid sizing).
int scan_blocks_per_sm(int W, bool rev, int variant, uint32_t nrows);
// Launches one scan over `a.nq` queries x ceil(rows / kScanThreads) row tiles.
cudaError_t launch_scan(int W, bool rev, int variant, const CUtensorMap* tmap, const ScanArgs& a,
                        cudaStream_t stream);

// Batches of one-word patterns (m <= 32): two patterns per thread share the text bytes, the row
// extraction and one 8-byte table load per character (scan2_kernel).  a.nq = number of patterns.
cudaError_t launch_scan2(bool rev, int variant, const CUtensorMap* tmap, const ScanArgs& a, cudaStream_t stream);

// Exact piece prefilter (scan_core.cuh): hits -> a.hit_keys; then one thread per hit re-scans
// the hit's neighbourhood with the full recurrences -> a.cand_*.
int filter_blocks_per_sm(int WF, int variant, bool pair);
cudaError_t launch_filter(int WF, bool rev, int variant, bool pair, const CUtensorMap* tmap, const ScanArgs& a,
                          cudaStream_t stream);
// q-gram bitmap prefilter (Dna, one pattern, both strands in one pass): a.feq = the 4^Q-bit table,
// a.fused = 1 reports every hit for the reversed partner slot too.  (Q, S) as planned by plan_qgram.
int qgram_blocks_per_sm(int Q, int S, int variant);
cudaError_t launch_qgram(int Q, int S, int variant, const CUtensorMap* tmap, const ScanArgs& a, cudaStream_t stream);
// The same filter over contiguous 2 KB tiles (tmap = the text as [ceil(n / 64)][64] bytes, box
// [32][64]); TMA data path only.
cudaError_t launch_qgram_seq(int Q, int S, const CUtensorMap* tmap, const ScanArgs& a, int sm_count,
                             cudaStream_t stream);
// Exact refinement of prefilter hits (refine_hit, Dna): every (share, alignment) that matches behind a
// hit becomes one entry (query slot, nominal end position) in `out` (capacity a.hit_cap, *out_count
// counts them); launch_verify then runs with ScanArgs::hit_exact = 1 on that list.
cudaError_t launch_refine(const ScanArgs& a, const uint8_t* rev_flags, uint64_t* out, uint32_t* out_span,
                          unsigned long long* out_count, cudaStream_t stream);
// Regional fallback: counts the prefilter hits per tile (a.tile_bytes) and marks / lists the tiles with
// more than max_hits_per_tile of them (nothing is marked when there are fewer than min_hits hits in
// all).  counts [ntiles] and *list_count must be zero on entry.
cudaError_t launch_tile_marks(const ScanArgs& a, uint32_t ntiles, uint32_t* counts, unsigned long long min_hits,
                              uint32_t max_hits_per_tile, uint8_t* dense, uint32_t* list, uint32_t* list_count,
                              cudaStream_t stream);
// Entries (query slot, nominal end, span) that cover a whole text in stretches of `stride` end
// positions: the full scan of patterns with more than 32 words runs as a re-scan of these.
cudaError_t launch_cover(uint64_t* keys, uint32_t* spans, unsigned long long* count, uint32_t nq, uint64_t n,
                         uint64_t stride, uint32_t k, cudaStream_t stream);
// nhits is read from a.hit_count on the device (clipped to a.hit_cap): no host round trip.
cudaError_t launch_verify(int W, const ScanArgs& a, const uint8_t* rev_flags, cudaStream_t stream);

// Many texts, one thread per (text, query) pair (search_texts / search_many).
struct TextsArgs {
  const uint8_t* base;       // texts back to back, every start 16-byte aligned, zero padded
  const uint64_t* offs;      // [ntexts]
  const uint64_t* lens;      // [ntexts]
  const uint8_t* rev_flags;  // [nq]
  uint32_t ntexts, nq;
  uint32_t include_pos0;
  uint32_t overhang;  // 1: positions <= min(n, m+k) of every text are left to the edge kernel
  uint32_t prefix;    // > 0: only the first `prefix` end positions (scan direction) of every text
};
cudaError_t launch_texts(int W, const ScanArgs& a, const TextsArgs& t, cudaStream_t stream);
// Many texts scanned as ONE concatenated text by the row-tiled kernels (Engine::search_texts):
// candidates (query, end position in the concatenation) become (text * nq + query, end position
// in the text); those in the padding between texts and those within the first `skip` end
// positions of their text -- where the state carried over from the previous text can lower a
// cost; texts_kernel with `prefix` owns them -- are dropped.
cudaError_t launch_concat_remap(const uint64_t* raw_keys, const uint32_t* raw_cost, const unsigned long long* raw_count,
                                uint64_t raw_cap, const ScanArgs& out, const TextsArgs& t, uint64_t total,
                                uint64_t skip, cudaStream_t stream);

// Overhang: the end positions that the overhang changes -- 0..min(n, m+k) (cheap left column)
// and n+1..n+steps (wildcard columns beyond the text, + floor(alpha * overshoot)) -- are
// computed per (slot, edge) by one thread each with the exact recurrences; the scan kernels
// leave positions <= ScanArgs::emit_min to it.
struct OverhangArgs {
  TextRef text;
  const uint8_t* rev_flags;  // [nq]
  uint32_t nslots;
  uint32_t init_pv[kMaxScanWords];  // vertical deltas of the left column L(j)
  int32_t left_total;           // L(m): cost of end position 0
  uint32_t steps;               // wildcard columns beyond the text
  float alpha;
};
cudaError_t launch_overhang_edges(int W, const ScanArgs& a, const OverhangArgs& o, cudaStream_t stream);

// flags[i] = 1 iff sorted candidate i is kept by the local-minima rule and, when `filter` is
// given, by the end-position predicates (end_filter_pass, scan_core.cuh).
cudaError_t launch_minima(const uint64_t* keys, const uint32_t* cost, uint64_t n, uint8_t* flags, bool all_minima,
                          const EndFilter* filter, cudaStream_t stream);
// only_best_match (reference src/search.rs:1392-1413): among the flagged candidates of every
// query slot keep the one with minimal cost, rightmost on ties.  best[nslots] is scratch.
cudaError_t launch_best(const uint64_t* keys, const uint32_t* cost, uint64_t n, uint8_t* flags,
                        unsigned long long* best, uint32_t nslots, cudaStream_t stream);

// One-block sort + selection for candidate lists of at most kSmallCandidates entries, driven by
// the device-side candidate count (*big = 1 and *nsel = 0 if the list is longer).
constexpr int kSmallCandidates = 2048;
cudaError_t launch_post_small(const uint64_t* keys, const uint32_t* cost, const unsigned long long* cand_count,
                              uint64_t cand_cap, uint64_t* sel_keys, unsigned long long* nsel,
                              unsigned long long* big, bool all_minima, int end_bit, cudaStream_t stream);

struct TraceArgs {
  TextRef text;
  int profile;
  const uint8_t* patterns;  // [nq_total][m] raw query bytes (already complemented for rc slots)
  const uint8_t* rev_flags;  // [nq_total] 1 = query scans the reversed text
  const uint32_t* eq;        // [nq_total][nrows][W]
  uint32_t nrows;
  uint32_t sh0, msk0;
  int32_t m, k;
  int32_t W;
  const uint64_t* keys;  // selected candidates (sorted)
  const uint32_t* costs; // without_trace: cost of every selected candidate (no traceback is run)
  float max_n_frac;      // >= 0: flag traced matches whose text slice holds too many N (src/n_filter.rs:59-61)
  float alpha;           // >= 0: overhang traceback (trace_one_ov)
  int32_t max_overhang;  // < 0: unlimited
  uint64_t first;        // slice [first, first+count) handled by this launch
  uint64_t count;
  const unsigned long long* count_dev;  // optional device-side total that clips the slice
  uint32_t smem_cols;      // set by launch_trace: the column store fits the block's shared memory
  uint32_t* scratch;       // trace_threads(count) * (m+k+1) * W * 2 words, interleaved by thread
  uint32_t* ops;           // [total][ops_words]
  uint32_t ops_words;
  GpuMatch* out;           // [total]
};
cudaError_t launch_trace(const TraceArgs& t, cudaStream_t stream);
uint64_t trace_threads(uint64_t count);  // threads launch_trace uses for a slice of `count` matches
// column stores launch_trace needs for a slice of `count` matches (one per thread, or per warp on the wide path)
uint64_t trace_slots(uint64_t count, int W, bool cost_only, bool overhang);

}  // namespace sb
