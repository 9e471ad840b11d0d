/*
 * sassy.h -- the reference's C ABI for Searcher::search, re-declared for the
 * B200 implementation (libsassy_b200.so).  Symbol names, argument order and the
 * record layout are those of the reference's generated header c/sassy.h:9-63
 * (implementation src/c.rs:16-131) so that a C caller written against the
 * reference (e.g. c/example.c) links against this library unchanged.
 *
 * Behavioural contract (reference src/c.rs):
 *   - sassy_searcher: alphabet is "ascii" | "dna" | "iupac", case-insensitive
 *     (c.rs:62-67); alpha = NAN disables overhang (c.rs:60).  The reference
 *     panics (process abort across extern "C") on a null pointer or unknown
 *     alphabet; so does this library, and likewise on overhang with a profile
 *     other than iupac or alpha outside [0, 1] (src/search.rs:373-383).
 *   - search: == Searcher::<P>::search, i.e. local-minima mode with traceback
 *     (c.rs:88-122).  The returned array is owned by the caller and must be
 *     released with sassy_matches_free(ptr, len) with the same len; for zero
 *     matches a non-null pointer is still returned (c.rs:112-127).
 *   - one searcher per thread; not re-entrant (search takes &mut Searcher).
 *   - Capacity limits of this implementation (the reference has none): patterns
 *     of at most 4096 characters (1024 with overhang, or in a search over many
 *     texts), texts shorter than 2^40 bytes.  A search()
 *     beyond a limit does NOT abort: it returns 0 matches (with a valid, freeable
 *     *out_matches), prints the reason to stderr and leaves it in
 *     sassy_gpu_last_error() (include/sassy_gpu.h), which is empty after every
 *     successful search().  Abort is reserved for the reference's own panics.
 * The GPU is selected with the environment variable SASSY_B200_DEVICE
 * (default 0).  There is no CPU fallback: without a B200 the constructor aborts.
 */
#ifndef SASSY_H
#define SASSY_H

#include <stdbool.h>
#include <stdint.h>
#include <stdlib.h>

typedef struct sassy_SearcherType sassy_SearcherType;

/* reference c/sassy.h:11-21 / src/c.rs:16-26: 40 bytes, cost at 32, strand at 36 */
typedef struct sassy_Match {
  uintptr_t text_start;
  uintptr_t text_end;
  uintptr_t pattern_start;
  uintptr_t pattern_end;
  int32_t cost;
  uint8_t strand; /* 0 = Fwd, 1 = Rc */
} sassy_Match;

#ifdef __cplusplus
extern "C" {
#endif

/* reference c/sassy.h:38 / src/c.rs:51-70 */
struct sassy_SearcherType *sassy_searcher(const char *alphabet, bool rc, float alpha);

/* reference c/sassy.h:43 / src/c.rs:73-81 */
void sassy_searcher_free(struct sassy_SearcherType *ptr);

/* reference c/sassy.h:52-58 / src/c.rs:88-122 (note: un-prefixed symbol, as in the reference) */
uintptr_t search(struct sassy_SearcherType *searcher, const uint8_t *pattern, uintptr_t pattern_len,
                 const uint8_t *text, uintptr_t text_len, uintptr_t k, struct sassy_Match **out_matches);

/* reference c/sassy.h:63 / src/c.rs:125-131 */
void sassy_matches_free(struct sassy_Match *ptr, uintptr_t len);

#ifdef __cplusplus
}
#endif

#endif /* SASSY_H */
