/*
 * sassy_cpu_port.c -- BASELINE / TEST INFRASTRUCTURE ONLY (never linked into sassy_b200).
 *
 * A C restatement ("port") of the reference's v1 CPU search engine, used as the
 * cpu_baseline / `--impl reference` leg of bench.py because the reference itself is
 * Rust and cannot be built in this image (no cargo/rustc; parity of this port is pinned
 * by tests/test_cpu_port.py against oracle/sassy_oracle.c, which in turn is pinned by the
 * reference's known-answer vectors).
 *
 * What is restated (reference file:line):
 *   - text-direction bit-vectors, 64 text characters per u64, LANES u64 lanes per SIMD
 *     register, each lane scanning its own chunk of the text with ceil((m+k)/64) blocks of
 *     overlap                                              src/search.rs:1008-1070,1100-1107
 *   - per block: profile-encode 64 text bytes into one eq mask per alphabet symbol
 *                                         src/profiles/dna.rs:26-40, iupac.rs:68-128
 *   - per pattern row: Myers block step with horizontal carry in/out
 *                                                          src/bitpacking.rs:63-85
 *   - early termination of the row loop with prefix-min checks, reset of pruned rows
 *                                                          src/search.rs:1128-1163,1244-1250
 *     prefix_min via pext + 8-bit table                    src/minima.rs:62-92
 *   - bottom-row walk emitting end positions with cost <= k, lane-overlap pruning
 *                                                          src/search.rs:1286-1369,1202-1240
 *   - rc strand = complement(pattern) on a reversed copy of the text
 *                                                          src/search.rs:813-878
 * SIMD width follows the reference's compile-time switch (src/lib.rs:177-185): 8 lanes with
 * AVX-512, else 4 lanes, via GCC vector extensions (compiled with -march=native).
 * Threads: the text is cut into one piece per thread with (m+k) overlap -- the most
 * favourable use of all host cores for a single (pattern, text) pair; the reference itself
 * would need several records to use more than one thread (src/search.rs:1520-1550).
 *
 * Output: every end position with cost <= k ("search_all" candidates) per strand; the
 * local-minima rule and traceback are applied by the caller's choice through
 * oracle/sassy_oracle.c-equivalent routines below (run-based rule).
 */
#define _GNU_SOURCE
#include <immintrin.h>
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#if defined(__AVX512F__) && defined(__AVX512BW__)
#define LANES 8
#else
#define LANES 4
#endif

typedef uint64_t vec __attribute__((vector_size(8 * LANES)));
typedef int64_t svec __attribute__((vector_size(8 * LANES)));

enum { PROFILE_DNA = 0, PROFILE_IUPAC = 1 };

static uint8_t IUPAC_CODE[32];
static int8_t PM_MIN[256], PM_SUM[256]; /* prefix-min / sum of 8 steps, bit=1:+1, bit=0:-1 */
static int tables_ready = 0;

static void init_tables(void) {
  if (tables_ready) return;
  memset(IUPAC_CODE, 255, sizeof IUPAC_CODE);
  const char *letters = "ACTUGNRYSWKMBDHVX";
  const uint8_t codes[] = {1, 2, 4, 4, 8, 15, 9, 6, 10, 5, 12, 3, 14, 13, 7, 11, 0};
  for (int i = 0; letters[i]; i++) IUPAC_CODE[letters[i] & 31] = codes[i];
  for (int b = 0; b < 256; b++) {
    int s = 0, mn = 0;
    for (int i = 0; i < 8; i++) {
      s += (b >> i) & 1 ? 1 : -1;
      if (s < mn) mn = s;
    }
    PM_MIN[b] = (int8_t)mn;
    PM_SUM[b] = (int8_t)s;
  }
  tables_ready = 1;
}

/* min over the 65 prefix sums (incl. empty) of the +-1 word (p,m): zero steps are squeezed
 * out with pext, the rest goes through a 256-entry table (src/minima.rs:62-92). */
static inline int prefix_min(uint64_t p, uint64_t m) {
  const uint64_t mask = p | m;
  int cnt = __builtin_popcountll(mask);
  if (cnt == 0) return 0;
#if defined(__BMI2__)
  uint64_t seq = _pext_u64(p, mask);
#else
  uint64_t seq = 0;
  int o = 0;
  for (uint64_t mm = mask; mm; mm &= mm - 1, o++) seq |= (uint64_t)((p >> __builtin_ctzll(mm)) & 1) << o;
#endif
  if (cnt < 64) seq |= ~0ull << cnt; /* pad with +1 steps */
  int mn = 0, s = 0;
  for (int i = 0; i < cnt; i += 8) {
    const int b = (int)((seq >> i) & 0xFF);
    if (s + PM_MIN[b] < mn) mn = s + PM_MIN[b];
    s += PM_SUM[b];
  }
  return mn;
}

/* Profile encoding of one 64-byte block into eq masks, one per symbol class. */
static inline void encode_dna(const uint8_t *b, uint64_t out[4]) {
#if defined(__AVX2__)
  const __m256i c0 = _mm256_loadu_si256((const __m256i *)b), c1 = _mm256_loadu_si256((const __m256i *)(b + 32));
  const __m256i three = _mm256_set1_epi8(3);
  const __m256i k0 = _mm256_and_si256(_mm256_srli_epi16(c0, 1), three);
  const __m256i k1 = _mm256_and_si256(_mm256_srli_epi16(c1, 1), three);
  for (int x = 0; x < 4; x++) {
    const __m256i v = _mm256_set1_epi8((char)x);
    const uint32_t lo = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(k0, v));
    const uint32_t hi = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(k1, v));
    out[x] = ((uint64_t)hi << 32) | lo;
  }
#else
  out[0] = out[1] = out[2] = out[3] = 0;
  for (int i = 0; i < 64; i++) out[(b[i] >> 1) & 3] |= 1ull << i;
#endif
}

static inline void encode_iupac(const uint8_t *b, const uint8_t *bases, int nbases, uint64_t *out) {
#if defined(__AVX2__)
  uint8_t code[64] __attribute__((aligned(32)));
  for (int i = 0; i < 64; i++) code[i] = IUPAC_CODE[b[i] & 31] & 0x0F; /* non-letters act as N */
  const __m256i c0 = _mm256_load_si256((const __m256i *)code), c1 = _mm256_load_si256((const __m256i *)(code + 32));
  const __m256i zero = _mm256_setzero_si256();
  for (int x = 0; x < nbases; x++) {
    const __m256i v = _mm256_set1_epi8((char)bases[x]);
    const uint32_t lo = ~(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_and_si256(c0, v), zero));
    const uint32_t hi = ~(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_and_si256(c1, v), zero));
    out[x] = ((uint64_t)hi << 32) | lo;
  }
#else
  for (int x = 0; x < nbases; x++) {
    uint64_t mk = 0;
    for (int i = 0; i < 64; i++)
      if (IUPAC_CODE[b[i] & 31] & 0x0F & bases[x]) mk |= 1ull << i;
    out[x] = mk;
  }
#endif
}

typedef struct {
  uint64_t pos;
  int32_t cost;
} Cand;

typedef struct {
  Cand *c;
  size_t n, cap;
} CandList;

static void cand_push(CandList *l, uint64_t pos, int32_t cost) {
  if (l->n == l->cap) {
    l->cap = l->cap ? l->cap * 2 : 256;
    l->c = (Cand *)realloc(l->c, l->cap * sizeof(Cand));
  }
  l->c[l->n].pos = pos;
  l->c[l->n].cost = cost;
  l->n++;
}

#define CHECK_AT_LEAST_ROWS 8
#define MAX_SYM 16

/* One strand, one piece of text t[0..n): all end positions e in [own_from, n] with cost<=k
 * (end position 0 is the caller's business).  Positions are relative to t. */
static void search_piece(int profile, const uint8_t *pat, int m, const uint8_t *t, size_t n, int k,
                         uint64_t own_from, CandList *out) {
  if (n == 0) return;
  /* pattern symbols -> index into the per-block mask array */
  uint8_t bases[MAX_SYM];
  int nbases = 0;
  int *sym = (int *)malloc(sizeof(int) * (size_t)m);
  for (int j = 0; j < m; j++) {
    if (profile == PROFILE_DNA) {
      sym[j] = (pat[j] >> 1) & 3;
    } else {
      const uint8_t c = IUPAC_CODE[pat[j] & 31];
      int x = 0;
      while (x < nbases && bases[x] != c) x++;
      if (x == nbases) bases[nbases++] = c;
      sym[j] = x;
    }
  }
  const int nsym = profile == PROFILE_DNA ? 4 : nbases;

  const size_t nblocks = (n + 63) / 64;
  const size_t overlap = ((size_t)m + (size_t)k + 63) / 64;
  const size_t bpc = ((nblocks > overlap ? nblocks - overlap : 0) + LANES - 1) / LANES; /* blocks per chunk */
  size_t last_i = 0; /* last block index that was encoded (LaneState::lane_end, src/search.rs:192-194) */

  vec *hp = (vec *)aligned_alloc(64, sizeof(vec) * (size_t)m);
  vec *hm = (vec *)aligned_alloc(64, sizeof(vec) * (size_t)m);
  for (int j = 0; j < m; j++) {
    for (int l = 0; l < LANES; l++) hp[j][l] = 1, hm[j][l] = 0;
  }
  uint64_t masks[LANES][MAX_SYM];
  uint8_t blockbuf[64];
  size_t prev_max_j = 0, prev_end_last_below = 0;
  CandList lanes_out[LANES];
  memset(lanes_out, 0, sizeof lanes_out);
  const vec kk1 = (vec){0} + (uint64_t)(k + 1);

  for (size_t i = 0; i < bpc + overlap; i++) {
    vec vp = {0}, vm = {0};
    last_i = i;
    for (int l = 0; l < LANES; l++) {
      const size_t start = ((size_t)l * bpc + i) * 64;
      const uint8_t *src;
      if (start + 64 <= n) {
        src = t + start;
      } else { /* pad with a byte that matches nothing relevant ('X'), src/search.rs:192-215 */
        memset(blockbuf, 'X', 64);
        if (start < n) memcpy(blockbuf, t + start, n - start);
        src = blockbuf;
      }
      if (start < n + 64) __builtin_prefetch(t + start + 256);
      if (profile == PROFILE_DNA)
        encode_dna(src, masks[l]);
      else
        encode_iupac(src, bases, nbases, masks[l]);
      (void)nsym;
    }
    vec dist_to_start = {0}, dist_to_end = {0};
    size_t cur_end_last_below = 0;
    int completed = 1;
    for (int j = 0; j < m; j++) {
      dist_to_start += hp[j];
      dist_to_start -= hm[j];
      vec eq;
      for (int l = 0; l < LANES; l++) eq[l] = masks[l][sym[j]];
      /* compute_block_simd, src/bitpacking.rs:63-85 */
      {
        const vec hp0 = hp[j], hm0 = hm[j];
        const vec vx = eq | vm;
        const vec eq2 = eq | hm0;
        const vec hx = (((eq2 & vp) + vp) ^ vp) | eq2;
        vec hpw = vm | ~(hx | vp);
        vec hmw = vp & hx;
        hp[j] = hpw >> 63;
        hm[j] = hmw >> 63;
        hpw = (hpw << 1) | hp0;
        hmw = (hmw << 1) | hm0;
        vp = hmw | ~(vx | hpw);
        vm = hpw & vx;
      }
      dist_to_end += hp[j];
      dist_to_end -= hm[j];
      {
        const svec lt = (svec)(dist_to_end < kk1);
        int any = 0;
        for (int l = 0; l < LANES; l++) any |= lt[l] != 0;
        if (any) cur_end_last_below = (size_t)j;
      }
      if ((size_t)j > prev_end_last_below) {
        int promising = 0;
        for (int l = 0; l < LANES; l++) {
          const int mn = prefix_min(vp[l], vm[l]) + (int)dist_to_start[l];
          if (mn <= k) {
            const size_t need = (size_t)(k - mn);
            prev_end_last_below = (size_t)j + (need > CHECK_AT_LEAST_ROWS ? need : CHECK_AT_LEAST_ROWS);
            promising = 1;
            break;
          }
        }
        if (!promising) {
          for (size_t j2 = (size_t)j + 1; j2 <= prev_max_j && j2 < (size_t)m; j2++) {
            for (int l = 0; l < LANES; l++) hp[j2][l] = 1, hm[j2][l] = 0;
          }
          prev_end_last_below = cur_end_last_below > CHECK_AT_LEAST_ROWS ? cur_end_last_below : CHECK_AT_LEAST_ROWS;
          prev_max_j = (size_t)j;
          completed = 0;
          break;
        }
      }
    }
    if (!completed) {
      /* should_terminate_early, src/search.rs:1253-1271 */
      if (i >= bpc) {
        const size_t d = 64 * (i - bpc);
        const size_t dist = d > prev_max_j ? d - prev_max_j : 0;
        if (dist > (size_t)k) break;
      }
      continue;
    }
    for (int l = 0; l < LANES; l++) {
      const int base_cost = (int)dist_to_start[l];
      if (prefix_min(vp[l], vm[l]) + base_cost > k) continue;
      const size_t base_pos = ((size_t)l * bpc + i) * 64;
      if (base_pos >= n) continue;
      int cost = base_cost;
      const uint64_t p = vp[l], mm = vm[l];
      for (int b = 0; b < 64; b++) { /* find_minima (all-minima branch), src/search.rs:1323-1335 */
        cost += (int)((p >> b) & 1) - (int)((mm >> b) & 1);
        const size_t pos = base_pos + (size_t)b + 1;
        if (pos > n) break;
        if (cost <= k) cand_push(&lanes_out[l], pos, cost);
      }
    }
    prev_end_last_below = cur_end_last_below > CHECK_AT_LEAST_ROWS ? cur_end_last_below : CHECK_AT_LEAST_ROWS;
    prev_max_j = (size_t)m - 1;
  }
  /* prune_lane_overlaps, src/search.rs:1202-1240: lane l keeps [end of lane l-1, end of lane l),
   * where a lane's end is the end of the last block it encoded (incl. the overlap blocks). */
  for (int l = 0; l < LANES; l++) {
    const size_t lo = l == 0 ? 0 : ((size_t)(l - 1) * bpc + last_i + 1) * 64;
    const size_t hi = l == LANES - 1 ? (size_t)-1 : ((size_t)l * bpc + last_i + 1) * 64;
    for (size_t a = 0; a < lanes_out[l].n; a++) {
      const uint64_t pos = lanes_out[l].c[a].pos;
      if (pos >= lo && pos < hi && pos >= own_from) cand_push(out, pos, lanes_out[l].c[a].cost);
    }
    free(lanes_out[l].c);
  }
  free(hp);
  free(hm);
  free(sym);
}

typedef struct {
  int profile, m, k, rev;
  const uint8_t *pat;
  const uint8_t *text;
  size_t n;
  size_t from, to; /* this thread owns end positions in (from, to] of the scan-direction text */
  CandList out;
} Job;

static void *job_main(void *arg) {
  Job *jb = (Job *)arg;
  const size_t halo = (size_t)jb->m + (size_t)jb->k;
  const size_t s = jb->from > halo ? jb->from - halo : 0;
  const size_t len = jb->to - s;
  if (!jb->rev) {
    CandList tmp = {0};
    search_piece(jb->profile, jb->pat, jb->m, jb->text + s, len, jb->k, jb->from - s + 1, &tmp);
    for (size_t a = 0; a < tmp.n; a++) cand_push(&jb->out, tmp.c[a].pos + s, tmp.c[a].cost);
    free(tmp.c);
  } else {
    /* reversed copy of the piece (the reference reverses the whole text, src/search.rs:137-139) */
    uint8_t *r = (uint8_t *)malloc(len ? len : 1);
    for (size_t i = 0; i < len; i++) r[i] = jb->text[jb->n - 1 - (s + i)];
    CandList tmp = {0};
    search_piece(jb->profile, jb->pat, jb->m, r, len, jb->k, jb->from - s + 1, &tmp);
    for (size_t a = 0; a < tmp.n; a++) cand_push(&jb->out, tmp.c[a].pos + s, tmp.c[a].cost);
    free(tmp.c);
    free(r);
  }
  return NULL;
}

/* All end positions with cost <= k of one strand (scan-direction coordinates, ascending),
 * using `threads` threads.  Returns a malloc'd array; *n_out its length. */
static Cand *strand_candidates(int profile, const uint8_t *pat, int m, const uint8_t *text, size_t n, int k,
                               int rev, int threads, size_t *n_out) {
  if (threads < 1) threads = 1;
  const size_t min_piece = 1 << 16;
  if ((size_t)threads > n / min_piece + 1) threads = (int)(n / min_piece + 1);
  Job *jobs = (Job *)calloc((size_t)threads, sizeof(Job));
  pthread_t *th = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
  int cpus[CPU_SETSIZE], ncpu = 0;
  {
    cpu_set_t allowed;
    if (sched_getaffinity(0, sizeof allowed, &allowed) == 0)
      for (int c = 0; c < CPU_SETSIZE; c++)
        if (CPU_ISSET(c, &allowed)) cpus[ncpu++] = c;
  }
  for (int i = 0; i < threads; i++) {
    jobs[i].profile = profile, jobs[i].m = m, jobs[i].k = k, jobs[i].rev = rev;
    jobs[i].pat = pat, jobs[i].text = text, jobs[i].n = n;
    jobs[i].from = n * (size_t)i / (size_t)threads;
    jobs[i].to = n * (size_t)(i + 1) / (size_t)threads;
    if (threads > 1) {
      pthread_create(&th[i], NULL, job_main, &jobs[i]);
      /* spread the workers over the allowed CPUs right away (the VM's scheduler is slow to
       * migrate freshly created threads, which would serialise short runs) */
      if (ncpu > 0) {
        cpu_set_t one;
        CPU_ZERO(&one);
        CPU_SET(cpus[i % ncpu], &one);
        pthread_setaffinity_np(th[i], sizeof one, &one);
      }
    }
  }
  if (threads == 1) job_main(&jobs[0]);
  size_t total = 0;
  for (int i = 0; i < threads; i++) {
    if (threads > 1) pthread_join(th[i], NULL);
    total += jobs[i].out.n;
  }
  Cand *all = (Cand *)malloc((total ? total : 1) * sizeof(Cand));
  size_t o = 0;
  for (int i = 0; i < threads; i++) {
    memcpy(all + o, jobs[i].out.c, jobs[i].out.n * sizeof(Cand));
    o += jobs[i].out.n;
    free(jobs[i].out.c);
  }
  free(jobs);
  free(th);
  *n_out = total;
  return all;
}

/* Run-based local-minima rule on an ascending candidate list (== src/search.rs:1344-1368
 * restricted to runs, == src/pattern_tiling/minima.rs:9-52). Keeps in place; returns new n. */
static size_t select_minima(Cand *c, size_t n) {
  size_t o = 0;
  size_t i = 0;
  while (i < n) {
    size_t j = i; /* run [i, e) */
    size_t e = i + 1;
    while (e < n && c[e].pos == c[e - 1].pos + 1) e++;
    int decreasing = 1;
    for (j = i; j < e; j++) {
      const int last = j + 1 == e;
      if (decreasing && (last || c[j + 1].cost > c[j].cost)) c[o++] = c[j];
      if (!last) decreasing = (c[j + 1].cost < c[j].cost) || (decreasing && c[j + 1].cost == c[j].cost);
    }
    i = e;
  }
  return o;
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static const uint8_t *complement_of(int profile, const uint8_t *p, int m, uint8_t *buf) {
  static const char *from = "ACTGRYSWKMBDHVNX", *to = "TGACYRSWMKVHDBNX";
  for (int i = 0; i < m; i++) {
    uint8_t c = p[i];
    if (profile == PROFILE_DNA) {
      c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'T' ? 'A' : c == 'G' ? 'C' : c;
    } else {
      for (int x = 0; from[x]; x++)
        if ((c & ~0x20) == from[x]) {
          c = (uint8_t)(to[x] | (c & 0x20));
          break;
        }
    }
    buf[i] = c;
  }
  return buf;
}

/* Public: Searcher::search / search_all end positions for one pattern (both strands when rc).
 * Writes up to cap (pos, cost, strand) triples (scan-direction end positions, i.e. reversed-
 * text coordinates on strand 1) and returns the total number found.  *seconds = wall time of
 * the search proper (candidates + minima selection). */
size_t cpu_port_search(int profile, const uint8_t *pattern, int m, const uint8_t *text, size_t n, int k, int rc,
                       int all, int threads, uint64_t *out_pos, int32_t *out_cost, uint8_t *out_strand, size_t cap,
                       double *seconds) {
  init_tables();
  const double t0 = now_s();
  size_t total = 0;
  uint8_t *buf = (uint8_t *)malloc((size_t)m + 1);
  for (int strand = 0; strand < (rc ? 2 : 1); strand++) {
    const uint8_t *p = strand ? complement_of(profile, pattern, m, buf) : pattern;
    size_t nc = 0;
    Cand *c = strand_candidates(profile, p, m, text, n, k, strand, threads, &nc);
    if (m <= k && n > 0) { /* end position 0 has cost m (src/search.rs:1320-1322) */
      c = (Cand *)realloc(c, (nc + 1) * sizeof(Cand));
      memmove(c + 1, c, nc * sizeof(Cand));
      c[0].pos = 0, c[0].cost = m;
      nc++;
    }
    if (!all) nc = select_minima(c, nc);
    for (size_t a = 0; a < nc; a++) {
      if (total < cap) {
        out_pos[total] = c[a].pos;
        out_cost[total] = c[a].cost;
        out_strand[total] = (uint8_t)strand;
      }
      total++;
    }
    free(c);
  }
  free(buf);
  if (seconds) *seconds = now_s() - t0;
  return total;
}

int cpu_port_lanes(void) { return LANES; }

/* =====================================================================================
 * v2: the pattern-tiled engine for batches of equal-length patterns (BASELINE configs 3, 5).
 *
 * Restates (reference file:line):
 *   - TQueries: one equality vector per text symbol, patterns across SIMD lanes
 *                                              src/pattern_tiling/tqueries.rs:53-134
 *   - search_ranges / myers_step: per text character and lane block
 *       xh = (((eq & vp) + vp) ^ vp) | eq;  mh = vp & xh;  ph = vn | ~(xh | vp);  xv = eq | vn;
 *       vp' = (mh << 1) | ~(xv | (ph << 1));  vn' = (ph << 1) & xv;  cost += ph[m-1] - mh[m-1]
 *     lanes with cost <= k are hits                   src/pattern_tiling/search.rs:148-175,326-407
 *   - lane width by pattern length (u32 for m <= 32, u16 for m <= 16)
 *                                              src/pattern_tiling/backend.rs, general.rs:247-292
 *   - hierarchical suffix prefilter: for 1 <= k <= 3 and m > 16 the 16-character SUFFIX of every
 *     pattern is searched first in u16 lanes (twice the patterns per register); only where the
 *     suffix has cost <= k is the full pattern evaluated    general.rs:60-102,294-313
 *   - end positions per pattern -> run-based local-minima rule   src/pattern_tiling/minima.rs:9-52
 * The reference gathers maximal ranges of hit positions and re-runs a forward pass with history per
 * range for the traceback; here every hit position is recorded with its cost and the traceback
 * below (trace_start) is run per reported match.  Forward strand only (bench.py's default, as the
 * reference's evals); Iupac profile.
 * Work split: 32-pattern x text-piece tasks over the threads, pieces with (m + k) overlap -- the
 * shape of the reference's own benchmark driver (evals/src/benchsuite/bench.rs:244-299).
 */
#if defined(__AVX512F__) && defined(__AVX512BW__)
#define V2_BYTES 64
#else
#define V2_BYTES 32
#endif
typedef uint32_t v32 __attribute__((vector_size(V2_BYTES)));
typedef int32_t s32 __attribute__((vector_size(V2_BYTES)));
typedef uint16_t v16 __attribute__((vector_size(V2_BYTES)));
typedef int16_t s16 __attribute__((vector_size(V2_BYTES)));
#define L32 (V2_BYTES / 4)
#define L16 (V2_BYTES / 2)

/* lanes with a <= b as a bit mask (one movemask instead of a lane loop) */
static inline uint64_t le_mask32(s32 a, s32 b) {
#if V2_BYTES == 64
  return (uint64_t)_mm512_cmple_epi32_mask((__m512i)a, (__m512i)b);
#elif defined(__AVX2__)
  return (uint64_t)(uint32_t)_mm256_movemask_ps((__m256)~_mm256_cmpgt_epi32((__m256i)a, (__m256i)b)) & 0xFFu;
#else
  uint64_t r = 0;
  for (int l = 0; l < L32; l++) r |= (uint64_t)(a[l] <= b[l]) << l;
  return r;
#endif
}
static inline uint64_t le_mask16(s16 a, s16 b) {
#if V2_BYTES == 64
  return (uint64_t)_mm512_cmple_epi16_mask((__m512i)a, (__m512i)b);
#elif defined(__AVX2__)
  const uint32_t bytes = (uint32_t)_mm256_movemask_epi8(~_mm256_cmpgt_epi16((__m256i)a, (__m256i)b));
  return (uint64_t)_pext_u32(bytes, 0x55555555u);
#else
  uint64_t r = 0;
  for (int l = 0; l < L16; l++) r |= (uint64_t)(a[l] <= b[l]) << l;
  return r;
#endif
}

typedef struct {
  uint32_t pat;
  uint64_t pos;
  int32_t cost;
} Hit2;
typedef struct {
  Hit2 *h;
  size_t n, cap;
} HitList;
static void hit_push(HitList *l, uint32_t pat, uint64_t pos, int32_t cost) {
  if (l->n == l->cap) {
    l->cap = l->cap ? l->cap * 2 : 1024;
    l->h = (Hit2 *)realloc(l->h, l->cap * sizeof(Hit2));
  }
  l->h[l->n].pat = pat, l->h[l->n].pos = pos, l->h[l->n].cost = cost;
  l->n++;
}

static inline uint32_t iupac_eq_word(const uint8_t *p, int m, int sym /* text byte & 31 */) {
  uint32_t w = 0;
  const uint8_t tc = IUPAC_CODE[sym] == 255 ? 0x0F : (IUPAC_CODE[sym] & 0x0F); /* non-letters act as N */
  for (int j = 0; j < m; j++)
    if (IUPAC_CODE[p[j] & 31] & tc) w |= 1u << j;
  return w;
}

/* Scalar full-pattern pass over t[from, to): end positions e in (own_from, to] with cost <= k. */
static void v2_scalar_window(const uint32_t *peq /*[32]*/, int m, int k, const uint8_t *t, size_t from, size_t to,
                             size_t own_from, uint32_t pat, HitList *out) {
  uint32_t vp = ~0u, vn = 0;
  int cost = m;
  const uint32_t top = 1u << (m - 1);
  for (size_t i = from; i < to; i++) {
    const uint32_t eq = peq[t[i] & 31];
    const uint32_t xh = (((eq & vp) + vp) ^ vp) | eq;
    const uint32_t mh = vp & xh, ph = vn | ~(xh | vp), xv = eq | vn;
    cost += ((ph & top) != 0) - ((mh & top) != 0);
    vp = (mh << 1) | ~(xv | (ph << 1));
    vn = (ph << 1) & xv;
    if (cost <= k && i + 1 > own_from) hit_push(out, pat, i + 1, cost);
  }
}

typedef struct {
  const uint8_t *pats; /* P x m */
  int m, k;
  uint32_t p0, p1;     /* patterns [p0, p1) */
  const uint8_t *text;
  size_t n, from, to;  /* owns end positions in (from, to] */
  int prefilter;
  HitList out;
} Job2;

static void *job2_main(void *arg) {
  Job2 *jb = (Job2 *)arg;
  const int m = jb->m, k = jb->k;
  const size_t halo = (size_t)m + (size_t)k;
  const size_t s = jb->from > halo ? jb->from - halo : 0;
  const uint32_t np = jb->p1 - jb->p0;
  /* scalar eq words of the full patterns (verification / narrow path) */
  uint32_t *peq = (uint32_t *)malloc((size_t)np * 32 * sizeof(uint32_t));
  for (uint32_t q = 0; q < np; q++)
    for (int sym = 0; sym < 32; sym++) peq[(size_t)q * 32 + sym] = iupac_eq_word(jb->pats + (size_t)(jb->p0 + q) * m, m, sym);
  if (jb->prefilter && m > 16) {
    /* pass 1: 16-character suffixes in u16 lanes */
    const int ms = 16;
    const uint32_t nb = (np + L16 - 1) / L16;
    v16 *eqv = (v16 *)aligned_alloc(64, sizeof(v16) * 32 * nb);
    for (int sym = 0; sym < 32; sym++)
      for (uint32_t b = 0; b < nb; b++) {
        v16 e = {0};
        for (int l = 0; l < L16; l++) {
          const uint32_t q = b * L16 + (uint32_t)l;
          if (q < np) e[l] = (uint16_t)(peq[(size_t)q * 32 + sym] >> (m - ms));
        }
        eqv[(size_t)sym * nb + b] = e;
      }
    v16 *vp = (v16 *)aligned_alloc(64, sizeof(v16) * nb), *vn = (v16 *)aligned_alloc(64, sizeof(v16) * nb);
    s16 *cost = (s16 *)aligned_alloc(64, sizeof(s16) * nb);
    for (uint32_t b = 0; b < nb; b++) vp[b] = (v16){0} - 1, vn[b] = (v16){0}, cost[b] = (s16){0} + (int16_t)ms;
    const s16 kk = (s16){0} + (int16_t)k;
    /* per pattern: end of the last verified window, so that overlapping windows are scanned once */
    size_t *done = (size_t *)calloc(np, sizeof(size_t));
    for (size_t i = s; i < jb->to; i++) {
      const v16 *e = eqv + (size_t)(jb->text[i] & 31) * nb;
      for (uint32_t b = 0; b < nb; b++) {
        const v16 eq = e[b], p = vp[b], nn = vn[b];
        const v16 xh = (((eq & p) + p) ^ p) | eq;
        const v16 mh = p & xh, ph = nn | ~(xh | p), xv = eq | nn;
        cost[b] += (s16)(ph >> (ms - 1)) - (s16)(mh >> (ms - 1));
        vp[b] = (mh << 1) | ~(xv | (ph << 1));
        vn[b] = (ph << 1) & xv;
        uint64_t lem = le_mask16(cost[b], kk);
        if (!lem || i + 1 <= jb->from) continue;
        for (; lem; lem &= lem - 1) {
          const int l = __builtin_ctzll(lem);
          const uint32_t q = b * L16 + (uint32_t)l;
          if (q >= np) continue;
          /* the full pattern can only end here with cost <= k: evaluate it on this end position
           * (window of m + k characters; consecutive hit positions extend the same window) */
          const size_t e1 = i + 1;
          if (done[q] >= e1) continue;
          size_t wfrom = e1 > halo ? e1 - halo : 0;
          size_t own = e1 - 1;
          if (done[q] > own) own = done[q];
          /* extend over the run of positions the suffix filter will report next (cheap look-ahead is
           * not available: verify position by position, sharing the warm-up through `done`) */
          v2_scalar_window(peq + (size_t)q * 32, m, k, jb->text, wfrom, e1, own > jb->from ? own : jb->from,
                           jb->p0 + q, &jb->out);
          done[q] = e1;
        }
      }
    }
    free(done);
    free(eqv), free(vp), free(vn), free(cost);
  } else {
    const uint32_t nb = (np + L32 - 1) / L32;
    v32 *eqv = (v32 *)aligned_alloc(64, sizeof(v32) * 32 * nb);
    for (int sym = 0; sym < 32; sym++)
      for (uint32_t b = 0; b < nb; b++) {
        v32 e = {0};
        for (int l = 0; l < L32; l++) {
          const uint32_t q = b * L32 + (uint32_t)l;
          if (q < np) e[l] = peq[(size_t)q * 32 + sym];
        }
        eqv[(size_t)sym * nb + b] = e;
      }
    v32 *vp = (v32 *)aligned_alloc(64, sizeof(v32) * nb), *vn = (v32 *)aligned_alloc(64, sizeof(v32) * nb);
    s32 *cost = (s32 *)aligned_alloc(64, sizeof(s32) * nb);
    for (uint32_t b = 0; b < nb; b++) vp[b] = (v32){0} - 1, vn[b] = (v32){0}, cost[b] = (s32){0} + m;
    const s32 kk = (s32){0} + k;
    for (size_t i = s; i < jb->to; i++) {
      const v32 *e = eqv + (size_t)(jb->text[i] & 31) * nb;
      for (uint32_t b = 0; b < nb; b++) {
        const v32 eq = e[b], p = vp[b], nn = vn[b];
        const v32 xh = (((eq & p) + p) ^ p) | eq;
        const v32 mh = p & xh, ph = nn | ~(xh | p), xv = eq | nn;
        cost[b] += (s32)((ph >> (m - 1)) & 1) - (s32)((mh >> (m - 1)) & 1);
        vp[b] = (mh << 1) | ~(xv | (ph << 1));
        vn[b] = (ph << 1) & xv;
        uint64_t lem = le_mask32(cost[b], kk);
        if (!lem || i + 1 <= jb->from) continue;
        for (; lem; lem &= lem - 1) {
          const int l = __builtin_ctzll(lem);
          if (b * L32 + (uint32_t)l < np) hit_push(&jb->out, jb->p0 + b * L32 + (uint32_t)l, i + 1, cost[b][l]);
        }
      }
    }
    free(eqv), free(vp), free(vn), free(cost);
  }
  free(peq);
  return NULL;
}

/* Traceback start of the match ending at `end` (reference src/trace.rs:273-406 on the window of
 * m + k characters, src/search.rs:1477-1478): scalar DP + the greedy walk, Iupac matching.
 * Returns text_start. */
static size_t trace_start_iupac(const uint8_t *p, int m, int k, const uint8_t *t, size_t end) {
  const size_t fill = (size_t)m + (size_t)k;
  const size_t off = end > fill ? end - fill : 0;
  const int w = (int)(end - off);
  static __thread int *D = NULL;
  static __thread size_t Dcap = 0;
  const size_t need = (size_t)(m + 1) * (size_t)(w + 1);
  if (need > Dcap) {
    D = (int *)realloc(D, need * sizeof(int));
    Dcap = need;
  }
#define DD(j, i) D[(size_t)(j) * (size_t)(w + 1) + (size_t)(i)]
  for (int i = 0; i <= w; i++) DD(0, i) = 0;
  for (int j = 1; j <= m; j++) {
    DD(j, 0) = j;
    const uint8_t pc = IUPAC_CODE[p[j - 1] & 31];
    for (int i = 1; i <= w; i++) {
      const uint8_t c = IUPAC_CODE[t[off + (size_t)i - 1] & 31];
      const uint8_t tc = c == 255 ? 0x0F : (c & 0x0F);
      const int mt = (pc & tc) != 0;
      int v = DD(j - 1, i - 1) + !mt;
      if (DD(j - 1, i) + 1 < v) v = DD(j - 1, i) + 1;
      if (DD(j, i - 1) + 1 < v) v = DD(j, i - 1) + 1;
      DD(j, i) = v;
    }
  }
  int j = m, i = w, g = DD(m, w);
  while (j > 0) {
    const uint8_t pc = IUPAC_CODE[p[j - 1] & 31];
    uint8_t tc = 0;
    if (i > 0) {
      const uint8_t c = IUPAC_CODE[t[off + (size_t)i - 1] & 31];
      tc = c == 255 ? 0x0F : (c & 0x0F);
    }
    if (i > 0 && DD(j - 1, i - 1) == g && (pc & tc)) {
      j--, i--;
    } else {
      g -= 1;
      if (i > 0 && DD(j - 1, i - 1) == g) j--, i--;
      else if (i > 0 && DD(j, i - 1) == g) i--;
      else if (DD(j - 1, i) == g) j--;
      else break;
    }
  }
#undef DD
  return off + (size_t)i;
}

typedef struct {
  Job2 *jobs;
  size_t ntasks;
  size_t *next;
  pthread_mutex_t *mu;
} Pool2;

static void *pool2_worker(void *arg) {
  Pool2 *pl = (Pool2 *)arg;
  for (;;) {
    pthread_mutex_lock(pl->mu);
    const size_t t = (*pl->next)++;
    pthread_mutex_unlock(pl->mu);
    if (t >= pl->ntasks) return NULL;
    job2_main(&pl->jobs[t]);
  }
}

static int hit_cmp(const void *a, const void *b) {
  const Hit2 *x = (const Hit2 *)a, *y = (const Hit2 *)b;
  if (x->pat != y->pat) return x->pat < y->pat ? -1 : 1;
  if (x->pos != y->pos) return x->pos < y->pos ? -1 : 1;
  return 0;
}

/* Public: search_encoded_patterns (forward strand) of P equal-length Iupac patterns (m <= 32).
 * all = 0: local minima per pattern; trace = 1: also compute text_start of every reported match.
 * Writes up to cap (pattern, end, cost, start) records; returns the total.  prefilter: 1 = the
 * reference's rule (suffix filter for 1 <= k <= 3, m > 16), 0 = never. */
size_t cpu_port_search_batch(const uint8_t *patterns, uint32_t n_patterns, int m, const uint8_t *text, size_t n, int k,
                             int all, int trace, int prefilter, int threads, uint32_t *out_pat, uint64_t *out_pos,
                             int32_t *out_cost, uint64_t *out_start, size_t cap, double *seconds) {
  init_tables();
  if (m < 1 || m > 32 || n_patterns == 0) return 0;
  const double t0 = now_s();
  if (threads < 1) threads = 1;
  const int use_pf = prefilter && k >= 1 && k <= 3 && m > 16;
  /* tasks: 32-pattern chunks x text pieces (evals/src/benchsuite/bench.rs:244-299) */
  const uint32_t chunk = 32 > L16 ? 32 : L16;
  const uint32_t nchunks = (n_patterns + chunk - 1) / chunk;
  size_t pieces = ((size_t)threads * 4 + nchunks - 1) / nchunks;
  if (pieces < 1) pieces = 1;
  if (pieces > n / (1 << 20) + 1) pieces = n / (1 << 20) + 1;
  const size_t ntasks = (size_t)nchunks * pieces;
  Job2 *jobs = (Job2 *)calloc(ntasks, sizeof(Job2));
  for (uint32_t c = 0; c < nchunks; c++)
    for (size_t pc = 0; pc < pieces; pc++) {
      Job2 *jb = &jobs[(size_t)c * pieces + pc];
      jb->pats = patterns, jb->m = m, jb->k = k;
      jb->p0 = c * chunk, jb->p1 = (c + 1) * chunk < n_patterns ? (c + 1) * chunk : n_patterns;
      jb->text = text, jb->n = n;
      jb->from = n * pc / pieces, jb->to = n * (pc + 1) / pieces;
      jb->prefilter = use_pf;
    }
  /* a pool of `threads` workers pulling tasks */
  size_t next = 0;
  pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
  Pool2 pool = {jobs, ntasks, &next, &mu};
  pthread_t *th = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
  for (int i = 1; i < threads; i++) pthread_create(&th[i], NULL, pool2_worker, &pool);
  pool2_worker(&pool);
  for (int i = 1; i < threads; i++) pthread_join(th[i], NULL);
  free(th);
  size_t total_hits = 0;
  for (size_t t = 0; t < ntasks; t++) total_hits += jobs[t].out.n;
  Hit2 *hits = (Hit2 *)malloc((total_hits ? total_hits : 1) * sizeof(Hit2));
  size_t o = 0;
  for (size_t t = 0; t < ntasks; t++) {
    memcpy(hits + o, jobs[t].out.h, jobs[t].out.n * sizeof(Hit2));
    o += jobs[t].out.n;
    free(jobs[t].out.h);
  }
  free(jobs);
  qsort(hits, total_hits, sizeof(Hit2), hit_cmp);
  /* de-duplicate (overlapping verification windows), then the run-based rule per pattern */
  size_t nh = 0;
  for (size_t a = 0; a < total_hits; a++)
    if (nh == 0 || hits[nh - 1].pat != hits[a].pat || hits[nh - 1].pos != hits[a].pos) hits[nh++] = hits[a];
  size_t total = 0;
  for (size_t a = 0; a < nh;) {
    size_t b = a;
    while (b < nh && hits[b].pat == hits[a].pat) b++;
    size_t cnt = b - a;
    Cand *c = (Cand *)malloc(cnt * sizeof(Cand));
    for (size_t x = 0; x < cnt; x++) c[x].pos = hits[a + x].pos, c[x].cost = hits[a + x].cost;
    if (!all) cnt = select_minima(c, cnt);
    for (size_t x = 0; x < cnt; x++) {
      if (total < cap) {
        out_pat[total] = hits[a].pat;
        out_pos[total] = c[x].pos;
        out_cost[total] = c[x].cost;
        out_start[total] = trace ? trace_start_iupac(patterns + (size_t)hits[a].pat * m, m, k, text, c[x].pos) : 0;
      }
      total++;
    }
    free(c);
    a = b;
  }
  free(hits);
  if (seconds) *seconds = now_s() - t0;
  return total;
}

int cpu_port_v2_lanes(void) { return L32; }
