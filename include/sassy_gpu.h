/*
 * sassy_gpu.h -- entry points of libsassy_b200.so beyond the reference's C ABI
 * (include/sassy.h).  The reference's C ABI (src/c.rs) only exposes
 * Searcher::search without CIGARs; these functions expose the rest of the Rust
 * surface of the hot path so that a Rust/Python/C host can bind all of it:
 *
 *   reference Rust item (file:line)                      this header
 *   ---------------------------------------------------  ---------------------------------
 *   Searcher::search            src/search.rs:510        sassy_gpu_search(all=0)
 *   Searcher::search_all        src/search.rs:685        sassy_gpu_search(all=1)
 *   CachedRev (text reused)     src/search.rs:144-166    sassy_gpu_text_upload + *_text calls
 *   Searcher::encode_patterns   src/search.rs:404        sassy_gpu_encode_patterns
 *   search_encoded_patterns     src/search.rs:415        sassy_gpu_search_encoded(all=0)
 *   search_all_encoded_patterns src/search.rs:426        sassy_gpu_search_encoded(all=1)
 *   Match{.., cigar}            src/search.rs:35-62      sassy_gpu_Match + sassy_gpu_result_ops
 *   Cigar::to_string (pa-types) src/lib.rs:83,107        sassy_gpu_cigar
 *   with_trace / without_trace  src/search.rs:446-449    sassy_gpu_set_trace
 *   only_best_match             src/search.rs:441-444    sassy_gpu_set_only_best_match
 *   set_max_n_frac              src/search.rs:452-458    sassy_gpu_set_max_n_frac
 *   new_*_with_overhang(alpha)  src/search.rs:385-402    alpha of sassy_searcher / sassy_gpu_searcher
 *   with_max_overhang           src/search.rs:436-439    sassy_gpu_set_max_overhang
 *   search_with_fn + PAM filter src/search.rs:767-784,   sassy_gpu_search_pam(_text)
 *                               bin/crispr.rs:198-221
 *   search_patterns             src/search.rs:648-683    sassy_gpu_search_patterns
 *   search_texts                src/search.rs:615-640    sassy_gpu_search_texts
 *   search_many                 src/search.rs:531-603    sassy_gpu_search_many
 *
 * Error handling: functions returning a pointer return NULL on error and
 * functions returning int return non-zero; sassy_gpu_last_error() then holds a
 * message (thread-local).  Nothing here falls back to the CPU.
 */
#ifndef SASSY_GPU_H
#define SASSY_GPU_H

#include <stddef.h>
#include <stdint.h>

#include "sassy.h"

typedef struct sassy_gpu_Text sassy_gpu_Text;         /* a text resident in HBM */
typedef struct sassy_gpu_Patterns sassy_gpu_Patterns; /* EncodedPatterns */
typedef struct sassy_gpu_Result sassy_gpu_Result;     /* Vec<Match> incl. CIGAR ops */
typedef struct sassy_gpu_Gather sassy_gpu_Gather;     /* multi-GPU match gather over peer memory */

/* One match; same fields as the reference's Match (src/search.rs:35-62).
 * The CIGAR ops of match i are the ops_len bytes at sassy_gpu_result_ops() + ops_off,
 * one char per op out of "=XID", in pattern direction. 72 bytes. */
typedef struct sassy_gpu_Match {
  uint64_t pattern_idx;
  uint64_t text_idx;
  uint64_t text_start;
  uint64_t text_end;
  uint64_t pattern_start;
  uint64_t pattern_end;
  int32_t cost;
  uint8_t strand; /* 0 = Fwd, 1 = Rc */
  uint8_t reserved[3];
  uint32_t ops_len;
  uint32_t reserved2;
  uint64_t ops_off;
} sassy_gpu_Match;

/* Timing/shape of the last search of a searcher (CUDA events on its stream). */
typedef struct sassy_gpu_Stats {
  float scan_ms;          /* scan kernel(s) only */
  float total_ms;         /* tables + scan + sort + minima + traceback + result copy */
  uint32_t scan_launches; /* our scan kernel launches */
  uint32_t aux_launches;  /* our minima + traceback launches */
  uint64_t candidates;    /* end positions with cost <= k seen by the scan */
  uint64_t matches;
  uint32_t row_bytes;     /* bytes of text per thread */
  uint32_t rows;
  uint32_t words;         /* 32-bit words per pattern bit-vector */
  uint32_t blocks_per_sm;
  uint32_t retries;       /* re-scans after a candidate-buffer overflow */
  uint32_t filter_words;  /* 32-bit words of the exact piece prefilter automaton; 0 = full scan */
  float filter_ms;        /* prefilter kernel(s) */
  float verify_ms;        /* re-scan of the prefilter's hit neighbourhoods */
  uint64_t hits;          /* text words in which a pattern piece occurs exactly */
  uint32_t filter_len;    /* piece length */
  uint32_t filter_fallback; /* 1: prefilter produced too many hits, the full scan was used */
  float transfer_ms;        /* host->device transfer of the text (host-text entry points) */
  uint32_t transfer_packed; /* 1: (part of) the text crossed PCIe at 2 bits per character (Dna) */
  uint64_t transfer_bytes;  /* bytes that crossed PCIe for the text */
  uint32_t filter_kind;     /* 0 none, 1 piece automaton (Shift-And), 2 q-gram bitmap, 3 SWAR suffix scan */
  uint32_t swar_lanes;      /* patterns per 32-bit word in the scan that produced the candidates (0/1 = one) */
  uint64_t confirmed;       /* prefilter hits that were re-scanned (q-gram: after the exact confirmation) */
  uint32_t dense_tiles;     /* regional fallback: tiles with too many hits, scanned whole with the exact recurrences */
  uint32_t reserved3;
} sassy_gpu_Stats;

#ifdef __cplusplus
extern "C" {
#endif

int sassy_gpu_device_count(void);
const char *sassy_gpu_last_error(void);
/* Device facts for the self-check (the reference's `sassy test` prints CPU features, src/lib.rs:187-281). */
int sassy_gpu_device_info(int device, char *name, size_t name_cap, int *sm_count, int *sm_clock_mhz,
                          size_t *total_mem, int *cc_major, int *cc_minor);

/* Like sassy_searcher (src/c.rs:51-70) with an explicit CUDA device; NULL on error. */
sassy_SearcherType *sassy_gpu_searcher(const char *alphabet, bool rc, float alpha, int device);

/* Scan kernel data path: 0 = TMA-staged (default), 1 = per-thread global loads (A/B baseline). */
int sassy_gpu_set_variant(sassy_SearcherType *searcher, int variant);
/* Exact piece prefilter (result-neutral, like the reference's suffix prefilter,
 * src/pattern_tiling/general.rs:294-313): 0 = off, 1 = when profitable (default), 2 = whenever possible. */
int sassy_gpu_set_filter(sassy_SearcherType *searcher, int mode);
/* Host->device transport of large Dna texts: 1 = packed to 2 bits per character by host threads
 * (default; result-neutral, texts with bytes outside ACGTacgt are sent as bytes), 0 = bytes. */
int sassy_gpu_set_transport(sassy_SearcherType *searcher, int mode);
int sassy_gpu_stats(const sassy_SearcherType *searcher, sassy_gpu_Stats *out);

/* Pinned host memory for texts that are searched through the host-pointer entry points. */
void *sassy_gpu_host_alloc(size_t bytes);
void sassy_gpu_host_free(void *ptr);

sassy_gpu_Text *sassy_gpu_text_upload(sassy_SearcherType *searcher, const uint8_t *text, size_t text_len);
sassy_gpu_Text *sassy_gpu_text_from_device(sassy_SearcherType *searcher, const void *device_ptr, size_t text_len);
size_t sassy_gpu_text_len(const sassy_gpu_Text *text);
void sassy_gpu_text_free(sassy_SearcherType *searcher, sassy_gpu_Text *text);

/* Searcher::search (all = 0) / search_all (all = 1) with CIGARs. */
sassy_gpu_Result *sassy_gpu_search(sassy_SearcherType *searcher, const uint8_t *pattern, size_t pattern_len,
                                   const uint8_t *text, size_t text_len, size_t k, int all);
sassy_gpu_Result *sassy_gpu_search_text(sassy_SearcherType *searcher, const uint8_t *pattern, size_t pattern_len,
                                        const sassy_gpu_Text *text, size_t k, int all);

/* Searcher options; they apply to every later search of this searcher like the reference's
 * builder methods.  trace = 0: matches carry the end position and the cost only
 * (text_start = pattern_start = UINT64_MAX; on the reverse strand text_end = UINT64_MAX and
 * text_start holds the known coordinate, src/search.rs:866-872), no CIGAR.
 * max_n_frac = 1.0 disables the N filter (src/search.rs:452-458). */
int sassy_gpu_set_trace(sassy_SearcherType *searcher, int trace);
int sassy_gpu_set_only_best_match(sassy_SearcherType *searcher, int on);
int sassy_gpu_set_max_n_frac(sassy_SearcherType *searcher, float max_n_frac);
/* Searcher::with_max_overhang (src/search.rs:436-439); < 0 = unlimited.  Overhang itself is the
 * `alpha` of the constructor (Iupac only, 0 <= alpha <= 1, src/search.rs:373-400): pattern
 * characters hanging over a text end cost alpha each; matches then report pattern_start /
 * pattern_end inside the pattern.  Cannot be combined with a PAM filter. */
int sassy_gpu_set_max_overhang(sassy_SearcherType *searcher, int max_overhang);

/* search_with_fn with the end filter of the reference's CRISPR mode: an end position is kept
 * only if the pam_len (<= 16) text characters before it match `pam` exactly under the
 * profile's is_match (on the reverse strand: the complemented PAM on the reversed text,
 * bin/crispr.rs:198-205).  all = 1 reports every passing end position (crispr.rs:218). */
sassy_gpu_Result *sassy_gpu_search_pam(sassy_SearcherType *searcher, const uint8_t *pattern, size_t pattern_len,
                                       const uint8_t *text, size_t text_len, size_t k, int all,
                                       const uint8_t *pam, size_t pam_len);
sassy_gpu_Result *sassy_gpu_search_pam_text(sassy_SearcherType *searcher, const uint8_t *pattern,
                                            size_t pattern_len, const sassy_gpu_Text *text, size_t k, int all,
                                            const uint8_t *pam, size_t pam_len);

/* Searcher::search_patterns: n_patterns patterns of equal length against one text; v1 semantics
 * per pattern, sassy_gpu_Match::pattern_idx set. */
sassy_gpu_Result *sassy_gpu_search_patterns(sassy_SearcherType *searcher, const uint8_t *const *patterns,
                                            size_t n_patterns, size_t pattern_len, const uint8_t *text,
                                            size_t text_len, size_t k);
/* Searcher::search_texts: one pattern against n_texts texts (sassy_gpu_Match::text_idx set);
 * short texts are searched by one kernel launch, one thread per (text, strand). */
sassy_gpu_Result *sassy_gpu_search_texts(sassy_SearcherType *searcher, const uint8_t *pattern, size_t pattern_len,
                                         const uint8_t *const *texts, const size_t *text_lens, size_t n_texts,
                                         size_t k);
/* Searcher::search_many: every pattern against every text; result ordered like the reference's
 * SearchMode::Single (pattern-major, then text, forward before reverse-complement matches). */
sassy_gpu_Result *sassy_gpu_search_many(sassy_SearcherType *searcher, const uint8_t *const *patterns,
                                        const size_t *pattern_lens, size_t n_patterns,
                                        const uint8_t *const *texts, const size_t *text_lens, size_t n_texts,
                                        size_t k);

/* Searcher::encode_patterns: n_patterns patterns of equal length pattern_len, back to back. */
sassy_gpu_Patterns *sassy_gpu_encode_patterns(sassy_SearcherType *searcher, const uint8_t *patterns,
                                              size_t n_patterns, size_t pattern_len);
void sassy_gpu_patterns_free(sassy_gpu_Patterns *patterns);
/* search_encoded_patterns (all = 0) / search_all_encoded_patterns (all = 1). */
sassy_gpu_Result *sassy_gpu_search_encoded(sassy_SearcherType *searcher, const sassy_gpu_Patterns *patterns,
                                           const sassy_gpu_Text *text, size_t k, int all);
sassy_gpu_Result *sassy_gpu_search_encoded_host(sassy_SearcherType *searcher, const sassy_gpu_Patterns *patterns,
                                                const uint8_t *text, size_t text_len, size_t k, int all);

/* ---- multi-GPU (one process per GPU of one box) -------------------------------------------
 * The reference fans (pattern, text) tasks out over threads and concatenates the match lists
 * (src/search.rs:531-603,1519-1549).  Here every rank searches its shard and the records of all
 * ranks reach every rank through ONE fused exchange: the traceback leaves the records in the
 * rank's slot of a receive buffer, a kernel stores them into the same slot of every peer's
 * buffer over NVLink (CUDA IPC mapped peer memory) and releases a step flag, a second kernel
 * acquires all flags and moves the records to pinned host memory.  No collective library call.
 *   1. sassy_gpu_gather_create on every rank (cap_records per rank and search, max_ops = m + k + 1
 *      of the longest search), 2. exchange the 64-byte handles (any transport), 3. connect,
 *   4. sassy_gpu_search_*_gathered in lock step on all ranks.
 * *complete = 1: the result holds the matches of all ranks in rank order with text_idx = source
 * rank.  *complete = 0: some rank's result did not fit the exchange; the result holds this rank's
 * matches only and the caller gathers them itself (sassy_b200/dist.py uses an NCCL all-gather). */
sassy_gpu_Gather *sassy_gpu_gather_create(sassy_SearcherType *searcher, int world, int rank, size_t cap_records,
                                          size_t max_ops);
int sassy_gpu_gather_handle(sassy_gpu_Gather *gather, uint8_t *out64);
int sassy_gpu_gather_connect(sassy_gpu_Gather *gather, const uint8_t *handles /* world x 64 bytes */);
void sassy_gpu_gather_free(sassy_gpu_Gather *gather);
sassy_gpu_Result *sassy_gpu_search_text_gathered(sassy_SearcherType *searcher, sassy_gpu_Gather *gather,
                                                 const uint8_t *pattern, size_t pattern_len,
                                                 const sassy_gpu_Text *text, size_t k, int all, int *complete);
sassy_gpu_Result *sassy_gpu_search_encoded_gathered(sassy_SearcherType *searcher, sassy_gpu_Gather *gather,
                                                    const sassy_gpu_Patterns *patterns, const sassy_gpu_Text *text,
                                                    size_t k, int all, int *complete);

/* One text cut into slabs over several GPUs: every rank searches its slab plus an
 * (m + k) halo on both sides with all = 1, the records of all ranks are gathered (text_idx =
 * source rank, coordinates relative to that rank's window), and this function drops the records a
 * slab does not own, makes the coordinates global and applies the local-minima rule
 * (src/search.rs:1344-1368) to the merged list (all = 0) -- a run of minima that crosses a slab
 * border is selected exactly as by an unsharded search.  Host code only; `ops` = the op bytes the
 * records' ops_off refer to.  See sassy_b200/dist.py: search_text_sharded. */
typedef struct sassy_gpu_Slab {
  uint64_t window_off; /* global position of the first character of the rank's window */
  uint64_t own_lo;     /* the rank owns the global text range [own_lo, own_hi) */
  uint64_t own_hi;
} sassy_gpu_Slab;
/* The whole sharded search in one call (all ranks in lock step): search_all of this rank's
 * window, fused gather, merge.  *complete = 0: some rank's records did not fit the exchange; the
 * result then holds this rank's unmerged search_all matches (window coordinates) and the caller
 * gathers them itself and calls sassy_gpu_merge_slabs. */
sassy_gpu_Result *sassy_gpu_search_text_sharded(sassy_SearcherType *searcher, sassy_gpu_Gather *gather,
                                                const uint8_t *pattern, size_t pattern_len,
                                                const sassy_gpu_Text *window, size_t k, int all,
                                                const sassy_gpu_Slab *slabs, size_t n_slabs, uint64_t n_global,
                                                int *complete);
/* Pipelined exchange (set on every rank before the first gathered search): a gathered search of
 * step s pushes its records and collects those of step s - 1, which the peers pushed a whole search
 * earlier -- the wait for the slowest rank leaves the step, the ranks may drift one step apart.  The
 * *_gathered / text_sharded calls then return the result of the PREVIOUS call (*complete = 2 and an
 * empty result on the first call), and sassy_gpu_text_sharded_flush returns the last one (*state = 1;
 * 2 = nothing pending).  A result that does not fit a rank's slot is an error in this mode (NULL):
 * use the lock-step mode for result sets beyond cap_records per rank. */
int sassy_gpu_gather_set_pipelined(sassy_gpu_Gather *gather, int on);
int sassy_gpu_gather_has_result(const sassy_gpu_Gather *gather);
sassy_gpu_Result *sassy_gpu_text_sharded_flush(sassy_SearcherType *searcher, sassy_gpu_Gather *gather,
                                               size_t pattern_len, int all, const sassy_gpu_Slab *slabs,
                                               size_t n_slabs, uint64_t n_global, int *state);
sassy_gpu_Result *sassy_gpu_merge_slabs(const sassy_gpu_Match *records, size_t n_records, const char *ops,
                                        const sassy_gpu_Slab *slabs, size_t n_slabs, uint64_t n_global, int all);

size_t sassy_gpu_result_len(const sassy_gpu_Result *result);
const sassy_gpu_Match *sassy_gpu_result_matches(const sassy_gpu_Result *result);
const char *sassy_gpu_result_ops(const sassy_gpu_Result *result);
/* Writes the run-length CIGAR of match i ("3=1X") NUL-terminated into buf; returns its length
 * (excluding the NUL) even if it did not fit. */
size_t sassy_gpu_cigar(const sassy_gpu_Result *result, size_t i, char *buf, size_t cap);
void sassy_gpu_result_free(sassy_gpu_Result *result);

#ifdef __cplusplus
}
#endif

#endif /* SASSY_GPU_H */
