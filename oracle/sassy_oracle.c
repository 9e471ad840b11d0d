/*
 * sassy_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A scalar CPU restatement of the approximate-string-matching path of the
 * reference (RagnarGrootKoerkamp/sassy @ 9ee854e): Searcher::search /
 * search_all (v1) and search_encoded_patterns / search_all_encoded_patterns
 * (v2), including traceback to CIGAR.  It is the *checker* for the CUDA path:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load it.  The product (sassy_b200/) never links or calls it.
 *
 * Parity pinning: the reference is Rust and cannot be compiled in this image
 * (no cargo/rustc), so this restatement is pinned against the known-answer
 * vectors held by the reference's own tests/docs (tests/golden/kat.json, each
 * entry cites its source file:line), and for the Searcher options, the PAM end
 * filter, search_many and overhang against the reference's tests of those
 * (tests/test_oracle_options.py, tests/test_cli.py: src/n_filter.rs:66-106,
 * src/search.rs:2372-2607,2929-3058, bin/crispr.rs:264-362, bin/grep.rs:795-813).
 *
 * The DP is the plain O(m*n) column recurrence (no bit tricks), so that it is
 * independent of both the reference's and the CUDA path's bit-parallel code:
 *   D[0][i] = 0, D[j][0] = j,
 *   D[j][i] = min(D[j-1][i-1] + !match(p_j,t_i), D[j][i-1] + 1, D[j-1][i] + 1)
 * (boundary conditions: reference src/search.rs:1058-1061 (hp=1,hm=0 => left
 *  column +1 per row) and src/search.rs:1101 (vp=vm=0 => top row 0)).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { PROFILE_DNA = 0, PROFILE_IUPAC = 1, PROFILE_ASCII = 2 };
enum { MODE_LOCAL_MINIMA = 0, MODE_ALL = 1 };

/* One reported match.  ops_off/ops_len index the op-string buffer (one char
 * per op, '=', 'X', 'I', 'D', in pattern direction).                        */
typedef struct {
  uint64_t text_start;
  uint64_t text_end;
  uint32_t pattern_idx;
  uint32_t pattern_start;
  uint32_t pattern_end;
  int32_t cost;
  uint32_t strand; /* 0 = Fwd, 1 = Rc */
  uint32_t ops_len;
  uint64_t ops_off;
  uint64_t text_idx;
} OracleMatch;

/* Searcher options and the end filter (reference src/search.rs:227-256,441-483,767-784). */
typedef struct {
  int without_trace;   /* src/search.rs:446-449 */
  int only_best;       /* src/search.rs:441-444 */
  float max_n_frac;    /* < 0: off (src/search.rs:452-458) */
  const uint8_t *pam;  /* end filter of bin/crispr.rs:198-205; NULL: none */
  size_t pam_len;
  float alpha;          /* overhang cost per pattern character, src/search.rs:231-233; < 0: off */
  int64_t max_overhang; /* src/search.rs:238-239; < 0: unlimited */
} OracleOpts;

typedef struct {
  OracleMatch *m;
  size_t n, cap;
  char *ops;
  size_t ops_n, ops_cap;
} OracleOut;

/* ------------------------------------------------------------------------ */
/* Profiles                                                                  */

/* IUPAC_CODE: reference src/profiles/iupac.rs:281-317. Indexed by c & 0x1F. */
static uint8_t IUPAC_CODE[32];
static uint8_t RC_IUPAC[256]; /* src/profiles/iupac.rs:235-278 */
static uint8_t RC_DNA[256];   /* src/profiles/dna.rs:121-133   */
static int tables_ready = 0;

static void init_tables(void) {
  if (tables_ready) return;
  memset(IUPAC_CODE, 255, sizeof IUPAC_CODE);
  const uint8_t A = 1, C = 2, T = 4, G = 8;
  IUPAC_CODE['A' & 31] = A;
  IUPAC_CODE['C' & 31] = C;
  IUPAC_CODE['T' & 31] = T;
  IUPAC_CODE['U' & 31] = T;
  IUPAC_CODE['G' & 31] = G;
  IUPAC_CODE['N' & 31] = A | C | T | G;
  IUPAC_CODE['R' & 31] = A | G;
  IUPAC_CODE['Y' & 31] = C | T;
  IUPAC_CODE['S' & 31] = G | C;
  IUPAC_CODE['W' & 31] = A | T;
  IUPAC_CODE['K' & 31] = G | T;
  IUPAC_CODE['M' & 31] = A | C;
  IUPAC_CODE['B' & 31] = C | G | T;
  IUPAC_CODE['D' & 31] = A | G | T;
  IUPAC_CODE['H' & 31] = A | C | T;
  IUPAC_CODE['V' & 31] = A | C | G;
  IUPAC_CODE['X' & 31] = 0;
  for (int i = 0; i < 256; i++) RC_IUPAC[i] = RC_DNA[i] = (uint8_t)i;
  const char *from = "ACTGRYSWKMBDHVNX";
  const char *to = "TGACYRSWMKVHDBNX";
  for (int i = 0; from[i]; i++) {
    RC_IUPAC[(uint8_t)from[i]] = (uint8_t)to[i];
    RC_IUPAC[(uint8_t)(from[i] | 0x20)] = (uint8_t)(to[i] | 0x20);
  }
  RC_DNA['A'] = 'T';
  RC_DNA['C'] = 'G';
  RC_DNA['T'] = 'A';
  RC_DNA['G'] = 'C';
  tables_ready = 1;
}

/* Equality used by the *search* DP.
 * Dna: both sides reduced to (c>>1)&3 (src/profiles/dna.rs:19-23,26-45).
 * Ascii: byte equality (the default case-sensitive Ascii<true>).
 * Iupac: pattern code (validated, <=15) AND low nibble of the text code; text
 *        bytes outside the table act as N (src/profiles/iupac.rs:68-128,
 *        319-330).                                                          */
static inline int search_eq(int profile, uint8_t p, uint8_t t) {
  if (profile == PROFILE_DNA) return ((p >> 1) & 3) == ((t >> 1) & 3);
  if (profile == PROFILE_ASCII) return p == t; /* src/profiles/ascii.rs:30-41,76-91 (case-sensitive) */
  return (IUPAC_CODE[p & 31] & (IUPAC_CODE[t & 31] & 0x0F)) != 0;
}

/* Equality used by the *traceback* (Profile::is_match).
 * Dna:  (a|0x20)==(b|0x20)              src/profiles/dna.rs:48-50
 * Iupac: code(a) & code(b) != 0         src/profiles/iupac.rs:136-138      */
static inline int trace_eq(int profile, uint8_t p, uint8_t t) {
  if (profile == PROFILE_DNA) return (p | 0x20) == (t | 0x20);
  if (profile == PROFILE_ASCII) return p == t; /* src/profiles/ascii.rs:44-51 */
  return (IUPAC_CODE[p & 31] & IUPAC_CODE[t & 31]) != 0;
}

/* Iupac::valid_seq, scalar branch: src/profiles/iupac.rs:195-201 */
int oracle_iupac_valid(const uint8_t *s, size_t n) {
  init_tables();
  for (size_t i = 0; i < n; i++) {
    uint8_t c = s[i] & (uint8_t)~0x20;
    if (c <= '@' || c >= 'Z' || IUPAC_CODE[c & 31] == 255) return 0;
  }
  return 1;
}

void oracle_complement(int profile, const uint8_t *in, size_t n, uint8_t *out) {
  init_tables();
  const uint8_t *tab = profile == PROFILE_DNA ? RC_DNA : RC_IUPAC;
  for (size_t i = 0; i < n; i++) out[i] = tab[in[i]];
}

void oracle_reverse_complement(int profile, const uint8_t *in, size_t n, uint8_t *out) {
  init_tables();
  const uint8_t *tab = profile == PROFILE_DNA ? RC_DNA : RC_IUPAC;
  for (size_t i = 0; i < n; i++) out[i] = tab[in[n - 1 - i]];
}

/* ------------------------------------------------------------------------ */
/* Output helpers                                                            */

static void out_push(OracleOut *o, OracleMatch mm, const char *ops, size_t nops) {
  if (o->n == o->cap) {
    o->cap = o->cap ? o->cap * 2 : 64;
    o->m = (OracleMatch *)realloc(o->m, o->cap * sizeof(OracleMatch));
  }
  if (o->ops_n + nops > o->ops_cap) {
    o->ops_cap = (o->ops_cap + nops) * 2 + 256;
    o->ops = (char *)realloc(o->ops, o->ops_cap);
  }
  mm.ops_off = o->ops_n;
  mm.ops_len = (uint32_t)nops;
  memcpy(o->ops + o->ops_n, ops, nops);
  o->ops_n += nops;
  o->m[o->n++] = mm;
}

/* ------------------------------------------------------------------------ */
/* Bottom-row costs c_i = D[m][i], i in [0,n].                               */
/* `rev` != 0 scans the byte-reversed text (t'[i] = t[n-1-i]),               */
/* reference src/search.rs:137-139,813-836.                                  */

static inline uint8_t text_at(const uint8_t *t, size_t n, int rev, size_t i) {
  return rev ? t[n - 1 - i] : t[i];
}

/* Overhang (reference src/search.rs:347-356,1274-1282,1693-1710; src/trace.rs:36-53):
 * with alpha set, pattern characters hanging over either end of the text cost alpha each
 * instead of 1: the left column of D is floor(j * alpha) (up to max_overhang rows, then +1 per
 * row), the text is followed by `steps` wildcard characters and an end position o characters
 * beyond the text costs floor(alpha * o) on top of the DP value.  All in f32 like the reference. */
static int32_t left_cost(size_t j, float alpha, int64_t mo) {
  if (alpha < 0.f) return (int32_t)j;
  size_t jm = (mo >= 0 && (size_t)mo < j) ? (size_t)mo : j;
  return (int32_t)floorf((float)jm * alpha) + (int32_t)(j - jm);
}

static size_t overhang_steps(size_t m, size_t k, float alpha, int64_t mo) {
  if (alpha < 0.f) return 0;
  float r = ceilf(((float)k + alpha) / alpha); /* Rust `as usize`: NaN -> 0, +inf saturates */
  size_t s = isnan(r) ? 0 : (r >= 1e18f ? SIZE_MAX : (size_t)r);
  if (m < s) s = m;
  if (mo >= 0 && (size_t)mo < s) s = (size_t)mo;
  return s;
}

static int32_t overshoot_cost(float alpha, size_t o) {
  return (alpha < 0.f || o == 0) ? 0 : (int32_t)floorf(alpha * (float)o);
}

/* c_i for i in [0, n + steps]; beyond the text every character matches (the reference pads
 * with 'N', src/search.rs:203) and the overshoot cost is added (src/search.rs:1274-1282). */
static int32_t *bottom_row_ov(int profile, const uint8_t *p, size_t m, const uint8_t *t, size_t n,
                              int rev, float alpha, int64_t mo, size_t steps) {
  int32_t *c = (int32_t *)malloc((n + steps + 1) * sizeof(int32_t));
  int32_t *col = (int32_t *)malloc((m + 1) * sizeof(int32_t));
  for (size_t j = 0; j <= m; j++) col[j] = left_cost(j, alpha, mo);
  c[0] = col[m];
  for (size_t i = 1; i <= n + steps; i++) {
    int wild = i > n;
    uint8_t tc = wild ? 0 : text_at(t, n, rev, i - 1);
    int32_t diag = col[0]; /* D[j-1][i-1] */
    col[0] = 0;
    for (size_t j = 1; j <= m; j++) {
      int32_t up = col[j - 1];  /* D[j-1][i]   */
      int32_t left = col[j];    /* D[j][i-1]   */
      int32_t v = diag + ((wild || search_eq(profile, p[j - 1], tc)) ? 0 : 1);
      if (left + 1 < v) v = left + 1;
      if (up + 1 < v) v = up + 1;
      diag = left;
      col[j] = v;
    }
    c[i] = col[m] + overshoot_cost(alpha, wild ? i - n : 0);
  }
  free(col);
  return c;
}

static int32_t *bottom_row(int profile, const uint8_t *p, size_t m, const uint8_t *t, size_t n,
                           int rev) {
  return bottom_row_ov(profile, p, m, t, n, rev, -1.f, -1, 0);
}

/* ------------------------------------------------------------------------ */
/* Traceback on a window: reference src/trace.rs:57-104 (fill: left column   */
/* j, top row 0) and src/trace.rs:273-406 (get_trace: greedy preference      */
/* '=' then 'X' then 'D' (consumes text) then 'I' (consumes pattern)).       */
/* The window is t[off .. end) in scan direction.                            */
/* Returns 0 on success, -1 if the greedy walk finds no ancestor (the        */
/* reference panics with "Trace failed").                                    */

/* With alpha >= 0 (src/trace.rs:57-104,273-406): the matrix has `cols` = m + k columns, the
 * left column is the overhang cost, columns beyond the text slice are wildcards; an end position
 * beyond the slice first steps diagonally back into the text (right overshoot, pattern_end
 * shrinks), reaching column 0 with j rows left is a left overshoot (pattern_start = j).       */
static int trace_window_ov(int profile, const uint8_t *p, size_t m, const uint8_t *t, size_t n,
                           int rev, size_t off, size_t end, float alpha, int64_t mo, size_t cols,
                           OracleMatch *mm, char **ops_out, size_t *nops_out) {
  size_t slice = (end < n ? end : n) - off; /* t[off .. min(end, n)) */
  size_t w = end - off;
  if (alpha < 0.f) cols = w;
  if (cols < w) cols = w;
  size_t stride = cols + 1;
  int32_t *D = (int32_t *)malloc((m + 1) * stride * sizeof(int32_t));
  for (size_t i = 0; i <= cols; i++) D[i] = 0;
  for (size_t j = 1; j <= m; j++) {
    D[j * stride] = left_cost(j, alpha, mo);
    for (size_t i = 1; i <= cols; i++) {
      int wild = i > slice;
      uint8_t tc = wild ? 0 : text_at(t, n, rev, off + i - 1);
      int32_t v = D[(j - 1) * stride + i - 1] + ((wild || search_eq(profile, p[j - 1], tc)) ? 0 : 1);
      int32_t l = D[j * stride + i - 1] + 1;
      int32_t u = D[(j - 1) * stride + i] + 1;
      if (l < v) v = l;
      if (u < v) v = u;
      D[j * stride + i] = v;
    }
  }
  char *ops = (char *)malloc(m + cols + 1);
  size_t nops = 0;
  size_t j = m, i = w;
  int32_t g = D[j * stride + i];
  int32_t total = g;
  size_t pattern_start = 0, pattern_end = m;
  int rc = 0;
  if (i > slice) { /* src/trace.rs:298-309 */
    size_t o = i - slice;
    pattern_end -= o;
    total += overshoot_cost(alpha, o);
    i -= o;
    j -= o;
  }
  while (j > 0) {
    if (i == 0 && alpha >= 0.f) { /* src/trace.rs:320-334 */
      pattern_start = j;
      g -= left_cost(j, alpha, mo);
      break;
    }
    if (i > 0 && D[(j - 1) * stride + i - 1] == g &&
        trace_eq(profile, p[j - 1], text_at(t, n, rev, off + i - 1))) {
      ops[nops++] = '=';
      j--, i--;
      continue;
    }
    g -= 1;
    if (i > 0 && D[(j - 1) * stride + i - 1] == g) {
      ops[nops++] = 'X';
      j--, i--;
      continue;
    }
    if (i > 0 && D[j * stride + i - 1] == g) {
      ops[nops++] = 'D';
      i--;
      continue;
    }
    if (D[(j - 1) * stride + i] == g) {
      ops[nops++] = 'I';
      j--;
      continue;
    }
    rc = -1;
    break;
  }
  if (alpha >= 0.f && rc == 0 && g != 0) rc = -1; /* assert_eq!(g, 0), src/trace.rs:390 */
  /* cigar.reverse(): src/trace.rs:393 */
  for (size_t a = 0, b = nops; a + 1 < b; a++, b--) {
    char tmp = ops[a];
    ops[a] = ops[b - 1];
    ops[b - 1] = tmp;
  }
  mm->cost = total;
  mm->text_start = off + i;
  mm->text_end = off + slice;
  mm->pattern_start = (uint32_t)pattern_start;
  mm->pattern_end = (uint32_t)pattern_end;
  *ops_out = ops;
  *nops_out = nops;
  free(D);
  return rc;
}

static int trace_window(int profile, const uint8_t *p, size_t m, const uint8_t *t, size_t n,
                        int rev, size_t off, size_t end, OracleMatch *mm, char **ops_out,
                        size_t *nops_out) {
  return trace_window_ov(profile, p, m, t, n, rev, off, end, -1.f, -1, 0, mm, ops_out, nops_out);
}

/* ------------------------------------------------------------------------ */
/* End-position selection.                                                   */

/* v1 rule: reference src/search.rs:1286-1369 run over the whole text with a
 * single lane (LANES-independent semantics; the reference's lane chunking can
 * duplicate a plateau at a lane boundary, which is not part of the contract).
 * Returns the number of selected positions written to `sel`.                */
static size_t select_v1(const int32_t *c, size_t n, int32_t k, int all, uint64_t *sel) {
  size_t ns = 0;
  if (n == 0) return 0; /* base_pos >= max_pos: src/search.rs:1314-1316 */
  int32_t prev_cost = c[0];
  size_t prev_pos = 0;
  int decreasing = 1; /* src/search.rs:1053 */
  if (all && c[0] <= k) sel[ns++] = 0; /* src/search.rs:1320-1322 */
  for (size_t pos = 1; pos <= n; pos++) {
    int32_t cost = c[pos];
    if (all) {
      if (cost <= k) sel[ns++] = pos;
    } else {
      if (decreasing && cost > prev_cost && prev_cost <= k) sel[ns++] = prev_pos;
      decreasing = (cost < prev_cost) || (decreasing && cost == prev_cost);
    }
    prev_cost = cost;
    prev_pos = pos;
  }
  if (!all && decreasing && prev_cost <= k) sel[ns++] = prev_pos; /* :1365-1368 */
  return ns;
}

/* v2 rule: positions are 0-based text indices idx in [0,n) whose end is idx+1
 * (reference src/pattern_tiling/search.rs:363-407: cost after consuming
 * text[idx]); maximal ranges of idx with cost<=k; inside each range
 * local_minima_indices (src/pattern_tiling/minima.rs:9-52).  Here the list of
 * passing (pos,cost) pairs is global and "gap > 1" splits ranges, exactly as
 * the function does for a list that spans several ranges.                   */
static size_t select_v2(const int32_t *c, size_t n, int32_t k, int all, uint64_t *sel) {
  size_t ns = 0;
  /* collect passing */
  size_t np = 0;
  int64_t *pp = (int64_t *)malloc((n + 1) * sizeof(int64_t));
  for (size_t idx = 0; idx < n; idx++)
    if (c[idx + 1] <= k) pp[np++] = (int64_t)idx;
  if (all) {
    for (size_t a = 0; a < np; a++) sel[ns++] = (uint64_t)pp[a] + 1;
    free(pp);
    return ns;
  }
  if (np == 0) {
    free(pp);
    return 0;
  }
  int64_t prev_pos = pp[0];
  int32_t prev_cost = c[pp[0] + 1];
  size_t prev_idx = 0;
  int last_trend = 2;
  for (size_t a = 1; a < np; a++) {
    int64_t pos = pp[a];
    int32_t cost = c[pos + 1];
    if (pos - prev_pos > 1) {
      if (last_trend != 1) sel[ns++] = (uint64_t)pp[prev_idx] + 1;
      last_trend = 2;
      prev_cost = cost;
      prev_idx = a;
      prev_pos = pos;
      continue;
    }
    if (cost > prev_cost && last_trend != 1) {
      sel[ns++] = (uint64_t)pp[prev_idx] + 1;
      last_trend = 1;
    } else if (cost < prev_cost) {
      last_trend = -1;
    } else if (cost == prev_cost && last_trend == 2) {
      last_trend = 0;
    }
    prev_cost = cost;
    prev_idx = a;
    prev_pos = pos;
  }
  if (last_trend != 1) sel[ns++] = (uint64_t)pp[prev_idx] + 1;
  free(pp);
  return ns;
}

/* ------------------------------------------------------------------------ */
/* One strand of v1: src/search.rs:884-937 + process_matches :1372-1517.     */

/* check_n_fraction: src/n_filter.rs:8-34 (f32 arithmetic as there). */
static int n_fraction_ok(const uint8_t *t, size_t n, int rev, size_t start, size_t end,
                         float max_n_frac, size_t denom) {
  if (start >= n) return 1;
  if (end <= start) return 1;
  size_t cnt = 0;
  for (size_t i = start; i < end; i++) {
    uint8_t c = text_at(t, n, rev, i);
    cnt += (c == 'N' || c == 'n');
  }
  float frac = (float)cnt / (float)(denom ? denom : end - start);
  return frac <= max_n_frac;
}

static int v1_one_strand(int profile, const uint8_t *p, size_t m, const uint8_t *t, size_t n,
                         int32_t k, int all, int rev, uint32_t pattern_idx, uint64_t text_idx,
                         const OracleOpts *o, OracleOut *out) {
  const float alpha = o ? o->alpha : -1.f;
  const int64_t mo = o ? o->max_overhang : -1;
  const size_t steps = overhang_steps(m, (size_t)k, alpha, mo);
  int32_t *c = bottom_row_ov(profile, p, m, t, n, rev, alpha, mo, steps);
  uint64_t *sel = (uint64_t *)malloc((n + steps + 2) * sizeof(uint64_t));
  size_t ns = select_v1(c, n + steps, k, all, sel); /* max_pos = n + steps, src/search.rs:1298-1308 */
  int rc = 0;
  /* end filter (search_with_fn, src/search.rs:895-905) with the closure of bin/crispr.rs:
   * the pam_len characters before the end match the PAM (complemented on the rc strand).
   * An end position shorter than the PAM panics in the reference; it is rejected here. */
  if (o && o->pam && o->pam_len) {
    uint8_t pamc[64];
    const uint8_t *pam = o->pam;
    if (rev) {
      oracle_complement(profile, o->pam, o->pam_len, pamc);
      pam = pamc;
    }
    size_t w = 0;
    for (size_t a = 0; a < ns; a++) {
      size_t end = sel[a];
      int ok = end >= o->pam_len;
      for (size_t i = 0; ok && i < o->pam_len; i++)
        ok = trace_eq(profile, text_at(t, n, rev, end - o->pam_len + i), pam[i]);
      if (ok) sel[w++] = end;
    }
    ns = w;
  }
  /* N end-point filter: src/search.rs:907-919, src/n_filter.rs:41-53 */
  if (o && o->max_n_frac >= 0.f) {
    size_t w = 0;
    for (size_t a = 0; a < ns; a++) {
      size_t end = sel[a] < n ? sel[a] : n;
      size_t mand = m > (size_t)k ? m - (size_t)k : 0;
      size_t start = end > mand ? end - mand : 0;
      if (n_fraction_ok(t, n, rev, start, end, o->max_n_frac, m + (size_t)k)) sel[w++] = sel[a];
    }
    ns = w;
  }
  /* only_best_match: rightmost end with minimal cost, src/search.rs:1392-1413 */
  if (o && o->only_best && ns > 0) {
    size_t best = 0;
    for (size_t a = 1; a < ns; a++)
      if (c[sel[a]] < c[sel[best]] || (c[sel[a]] == c[sel[best]] && sel[a] > sel[best])) best = a;
    sel[0] = sel[best];
    ns = 1;
  }
  for (size_t a = 0; a < ns; a++) {
    size_t end = sel[a];
    size_t fill = m + (size_t)k;
    size_t off = end > fill ? end - fill : 0; /* saturating_sub: :1477 */
    OracleMatch mm;
    memset(&mm, 0, sizeof mm);
    char *ops = NULL;
    size_t nops = 0;
    if (o && o->without_trace) {
      /* src/search.rs:1464-1475 and the rc mapping :859-872 */
      mm.text_start = UINT64_MAX;
      mm.text_end = end < n ? end : n;
      mm.pattern_start = UINT32_MAX;
      mm.pattern_end = (uint32_t)(m - (end > n ? end - n : 0)); /* src/search.rs:1470 */
      mm.cost = c[end];
    } else {
      if (trace_window_ov(profile, p, m, t, n, rev, off, end, alpha, mo, fill, &mm, &ops, &nops) != 0) rc = -1;
      /* traced N filter: src/search.rs:924-934, src/n_filter.rs:59-61 */
      if (o && o->max_n_frac >= 0.f &&
          !n_fraction_ok(t, n, rev, mm.text_start, mm.text_end, o->max_n_frac, 0)) {
        free(ops);
        continue;
      }
    }
    mm.pattern_idx = pattern_idx;
    mm.text_idx = text_idx;
    mm.strand = rev ? 1 : 0;
    if (rev) {
      /* map to forward coordinates: src/search.rs:859-877 */
      uint64_t rs = mm.text_start, re = mm.text_end;
      mm.text_start = n - re;
      mm.text_end = (o && o->without_trace) ? UINT64_MAX : n - rs;
    }
    out_push(out, mm, ops ? ops : "", nops);
    free(ops);
  }
  free(sel);
  free(c);
  return rc;
}

/* ------------------------------------------------------------------------ */
/* Public entry points (C ABI, loaded through ctypes by tests/ and bench.py) */

OracleOut *oracle_out_new(void) { return (OracleOut *)calloc(1, sizeof(OracleOut)); }
void oracle_out_free(OracleOut *o) {
  if (!o) return;
  free(o->m);
  free(o->ops);
  free(o);
}
size_t oracle_out_len(const OracleOut *o) { return o->n; }
const OracleMatch *oracle_out_matches(const OracleOut *o) { return o->m; }
const char *oracle_out_ops(const OracleOut *o) { return o->ops; }

/* Searcher::<P>::search / search_all (src/search.rs:510-525,685-700,787-881).
 * Output order: forward matches by ascending end, then rc matches by
 * ascending end in the reversed text.  Returns 0, or -1 when a traceback
 * failed (the reference would panic), -2 on an invalid IUPAC pattern.       */
static int search_pair(int profile, const uint8_t *pattern, size_t m, const uint8_t *text, size_t n,
                       uint32_t k, int rc_strand, int all, uint32_t pattern_idx, uint64_t text_idx,
                       const OracleOpts *o, OracleOut *out) {
  int rc = v1_one_strand(profile, pattern, m, text, n, (int32_t)k, all, 0, pattern_idx, text_idx, o, out);
  if (rc_strand) {
    uint8_t *cp = (uint8_t *)malloc(m + 1);
    oracle_complement(profile, pattern, m, cp);
    if (v1_one_strand(profile, cp, m, text, n, (int32_t)k, all, 1, pattern_idx, text_idx, o, out) != 0) rc = -1;
    free(cp);
  }
  return rc;
}

int oracle_search(int profile, const uint8_t *pattern, size_t m, const uint8_t *text, size_t n,
                  uint32_t k, int rc_strand, int all, OracleOut *out) {
  init_tables();
  if (profile == PROFILE_IUPAC && !oracle_iupac_valid(pattern, m)) return -2;
  return search_pair(profile, pattern, m, text, n, k, rc_strand, all, 0, 0, NULL, out);
}

/* search / search_all / search_with_fn(PAM) under the Searcher options. */
int oracle_search_opts(int profile, const uint8_t *pattern, size_t m, const uint8_t *text, size_t n,
                       uint32_t k, int rc_strand, int all, int without_trace, int only_best,
                       float max_n_frac, const uint8_t *pam, size_t pam_len, float alpha,
                       int64_t max_overhang, OracleOut *out) {
  init_tables();
  if (profile == PROFILE_IUPAC && !oracle_iupac_valid(pattern, m)) return -2;
  if (pam_len > 64) return -3;
  OracleOpts o = {without_trace, only_best, max_n_frac, pam, pam_len, alpha, max_overhang};
  if (alpha >= 0.f && profile != PROFILE_IUPAC) return -4; /* src/search.rs:373-383 */
  if (alpha >= 0.f && pam_len) return -5; /* the PAM closure would index beyond the text */
  return search_pair(profile, pattern, m, text, n, k, rc_strand, all, 0, 0, &o, out);
}

/* Searcher::search_many in SearchMode::Single (src/search.rs:531-553,1519-1549): every pattern
 * against every text with `search`, pattern-major; the other modes must return the same set
 * (src/search.rs:3624-3730).  Patterns / texts are concatenated; *_lens give the lengths.   */
int oracle_search_many(int profile, const uint8_t *patterns, const uint64_t *pattern_lens,
                       size_t n_patterns, const uint8_t *texts, const uint64_t *text_lens,
                       size_t n_texts, uint32_t k, int rc_strand, int without_trace, int only_best,
                       float max_n_frac, float alpha, int64_t max_overhang, OracleOut *out) {
  init_tables();
  OracleOpts o = {without_trace, only_best, max_n_frac, NULL, 0, alpha, max_overhang};
  if (alpha >= 0.f && profile != PROFILE_IUPAC) return -4;
  int rc = 0;
  const uint8_t *p = patterns;
  for (size_t pi = 0; pi < n_patterns; pi++) {
    if (profile == PROFILE_IUPAC && !oracle_iupac_valid(p, pattern_lens[pi])) return -2;
    const uint8_t *t = texts;
    for (size_t ti = 0; ti < n_texts; ti++) {
      if (search_pair(profile, p, pattern_lens[pi], t, text_lens[ti], k, rc_strand, 0, (uint32_t)pi, ti,
                      &o, out) != 0)
        rc = -1;
      t += text_lens[ti];
    }
    p += pattern_lens[pi];
  }
  return rc;
}

/* Searcher::search_encoded_patterns / search_all_encoded_patterns
 * (src/search.rs:415-433 -> src/pattern_tiling/general.rs:335-404).
 * `patterns` holds n_patterns patterns of equal length m back to back.
 * With rc_strand the reverse complement of every pattern is searched in the
 * forward text as query n_patterns + idx (src/pattern_tiling/tqueries.rs:75-80)
 * and reported with pattern_idx = idx, strand = Rc, coordinates and CIGAR in
 * the direction of the searched (rc) pattern (src/pattern_tiling/trace.rs:444-449).
 * Output order here: by query index, then ascending end (the reference's
 * order depends on range-closing time and is not part of the contract; its
 * own fuzz test sorts before comparing, src/pattern_tiling/search.rs:748-796).
 * The traceback window starts at max(0, range.start - (m+k)) like the
 * reference (src/pattern_tiling/trace.rs:65-82), not at end-(m+k).           */
int oracle_search_encoded_nfrac(int profile, const uint8_t *patterns, size_t n_patterns, size_t m,
                                const uint8_t *text, size_t n, uint32_t k, int rc_strand, int all,
                                float max_n_frac, OracleOut *out);
int oracle_search_encoded_opts(int profile, const uint8_t *patterns, size_t n_patterns, size_t m,
                               const uint8_t *text, size_t n, uint32_t k, int rc_strand, int all,
                               float max_n_frac, float alpha, int64_t max_overhang, OracleOut *out);

int oracle_search_encoded(int profile, const uint8_t *patterns, size_t n_patterns, size_t m,
                          const uint8_t *text, size_t n, uint32_t k, int rc_strand, int all,
                          OracleOut *out) {
  return oracle_search_encoded_nfrac(profile, patterns, n_patterns, m, text, n, k, rc_strand, all, -1.f, out);
}

/* max_n_frac >= 0: traced N filter of the v2 engine (src/pattern_tiling/general.rs:399-402). */
int oracle_search_encoded_nfrac(int profile, const uint8_t *patterns, size_t n_patterns, size_t m,
                                const uint8_t *text, size_t n, uint32_t k, int rc_strand, int all,
                                float max_n_frac, OracleOut *out) {
  init_tables();
  if (m == 0 || m > 64) return -3; /* general.rs:285-291, tqueries.rs:60-66 */
  for (size_t q = 0; q < n_patterns; q++)
    if (profile == PROFILE_IUPAC && !oracle_iupac_valid(patterns + q * m, m)) return -2;
  int rc = 0;
  size_t nq = n_patterns * (rc_strand ? 2 : 1);
  uint8_t *buf = (uint8_t *)malloc(m + 1);
  uint64_t *sel = (uint64_t *)malloc((n + 2) * sizeof(uint64_t));
  for (size_t q = 0; q < nq; q++) {
    const uint8_t *p = patterns + (q % n_patterns) * m;
    if (q >= n_patterns) {
      oracle_reverse_complement(PROFILE_IUPAC, p, m, buf); /* tqueries.rs:2,78 uses iupac rc */
      p = buf;
    }
    int32_t *c = bottom_row(profile, p, m, text, n, 0);
    size_t ns = select_v2(c, n, (int32_t)k, all, sel);
    for (size_t a = 0; a < ns; a++) {
      size_t end = sel[a];
      /* range start: walk left while cost <= k (idx = end-1 is in the range) */
      size_t idx = end - 1;
      while (idx > 0 && c[idx] <= (int32_t)k) idx--; /* c[idx] is cost of idx-1 */
      size_t range_start = idx; /* first idx of the range */
      size_t fill = m + (size_t)k;
      size_t off = range_start > fill ? range_start - fill : 0;
      OracleMatch mm;
      memset(&mm, 0, sizeof mm);
      char *ops;
      size_t nops;
      if (trace_window(profile, p, m, text, n, 0, off, end, &mm, &ops, &nops) != 0) rc = -1;
      mm.pattern_idx = (uint32_t)(q % n_patterns);
      mm.strand = q >= n_patterns ? 1 : 0;
      if (max_n_frac < 0.f || n_fraction_ok(text, n, 0, mm.text_start, mm.text_end, max_n_frac, 0))
        out_push(out, mm, ops, nops);
      free(ops);
    }
    free(c);
  }
  free(sel);
  free(buf);
  return rc;
}

/* Encoded patterns with overhang (PatterntilingSearcher::new(alpha), src/pattern_tiling/search.rs:
 * 91-117,223-323).  The reference pins this engine to the v1 one: its fuzz_against_sassy_batch
 * (src/pattern_tiling/search.rs:690-848, run with alpha = 0.5 at :886-896) requires the v2 matches
 * to EQUAL -- all Match fields incl. pattern_start / pattern_end / CIGAR -- the forward v1 overhang
 * search of every pattern and, for the rc strand, of its reverse complement with the strand
 * relabelled.  That definition is restated here.                                                */
int oracle_search_encoded_opts(int profile, const uint8_t *patterns, size_t n_patterns, size_t m,
                               const uint8_t *text, size_t n, uint32_t k, int rc_strand, int all,
                               float max_n_frac, float alpha, int64_t max_overhang, OracleOut *out) {
  if (alpha < 0.f)
    return oracle_search_encoded_nfrac(profile, patterns, n_patterns, m, text, n, k, rc_strand, all, max_n_frac, out);
  init_tables();
  if (profile != PROFILE_IUPAC) return -4;
  if (m == 0 || m > 64) return -3;
  for (size_t q = 0; q < n_patterns; q++)
    if (!oracle_iupac_valid(patterns + q * m, m)) return -2;
  OracleOpts o = {0, 0, max_n_frac, NULL, 0, alpha, max_overhang};
  int rc = 0;
  size_t nq = n_patterns * (rc_strand ? 2 : 1);
  uint8_t *buf = (uint8_t *)malloc(m + 1);
  for (size_t q = 0; q < nq; q++) {
    const uint8_t *p = patterns + (q % n_patterns) * m;
    if (q >= n_patterns) {
      oracle_reverse_complement(PROFILE_IUPAC, p, m, buf);
      p = buf;
    }
    size_t first = out->n;
    if (v1_one_strand(profile, p, m, text, n, (int32_t)k, all, 0, (uint32_t)(q % n_patterns), 0, &o, out) != 0) rc = -1;
    if (q >= n_patterns)
      for (size_t i = first; i < out->n; i++) out->m[i].strand = 1;
  }
  free(buf);
  return rc;
}

/* Bottom row for inspection in tests: out[i] = D[m][i], i in [0,n]. */
void oracle_bottom_row(int profile, const uint8_t *pattern, size_t m, const uint8_t *text,
                       size_t n, int rev, int32_t *out) {
  init_tables();
  int32_t *c = bottom_row(profile, pattern, m, text, n, rev);
  memcpy(out, c, (n + 1) * sizeof(int32_t));
  free(c);
}
