/* A plain C caller of the drop-in boundary, written against include/sassy.h only (the four
 * symbols of the reference's c/sassy.h:38-63) plus two calls of include/sassy_gpu.h.  Compiled and
 * linked by tests/test_c_abi.py (no GPU needed for that); executed by tests/test_gpu_options.py.
 * Prints one line per match: text_start text_end pattern_start pattern_end cost strand [cigar]. */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "sassy_gpu.h"

int main(int argc, char **argv) {
  const char *alphabet = argc > 1 ? argv[1] : "dna";
  const char *pattern = argc > 2 ? argv[2] : "ATCG";
  const char *text = argc > 3 ? argv[3] : "CCCATCACCC";
  unsigned long k = argc > 4 ? strtoul(argv[4], NULL, 10) : 1;
  sassy_SearcherType *s = sassy_searcher(alphabet, true, NAN);
  sassy_Match *ms = NULL;
  uintptr_t n = search(s, (const uint8_t *)pattern, strlen(pattern), (const uint8_t *)text, strlen(text), k, &ms);
  for (uintptr_t i = 0; i < n; i++)
    printf("%zu %zu %zu %zu %d %d\n", (size_t)ms[i].text_start, (size_t)ms[i].text_end, (size_t)ms[i].pattern_start,
           (size_t)ms[i].pattern_end, ms[i].cost, (int)ms[i].strand);
  sassy_matches_free(ms, n);
  /* the extended entry point also returns the CIGARs */
  sassy_gpu_Result *r = sassy_gpu_search(s, (const uint8_t *)pattern, strlen(pattern), (const uint8_t *)text,
                                         strlen(text), k, 0);
  if (!r) {
    fprintf(stderr, "%s\n", sassy_gpu_last_error());
    return 1;
  }
  for (size_t i = 0; i < sassy_gpu_result_len(r); i++) {
    char buf[256];
    sassy_gpu_cigar(r, i, buf, sizeof buf);
    printf("cigar %s\n", buf);
  }
  sassy_gpu_result_free(r);
  sassy_searcher_free(s);
  return 0;
}
