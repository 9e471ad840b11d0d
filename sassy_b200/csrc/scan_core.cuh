// Per-thread logic of the search path, shared by the CUDA kernels and by the
// host emulator used in the CPU-only tests (tests/test_host_emulation.py).
//
// What it computes (reference semantics, restated):
//   * the bottom row c_i = D[m][i] of the semi-global edit-distance matrix
//     (top row 0, left column j): reference src/search.rs:1058-1061,1101 and
//     src/pattern_tiling/search.rs:91-117,148-175;
//   * every end position with c_i <= k is emitted as a candidate
//     (reference src/search.rs:1320-1335 "all minima" branch,
//      src/pattern_tiling/search.rs:389-407).
// How it computes it is NOT the reference's layout: the bit-vector runs along
// the pattern in W 32-bit words (Myers'99 / Hyyro'01 recurrences), the pattern
// sits in the TOP m bits of the 32*W-bit vector and the unused low rows are
// wildcards with vertical delta 0, so the score delta is always bit 31 of the
// top word and never needs a per-pattern shift.  The text is cut into rows of
// `ltot` bytes (one thread per row, cold start + (m+k)-byte warm-up from the
// neighbouring row), an idea shared with the reference's lane chunking
// (src/search.rs:1018-1049) but at 10^5-way width.
#pragma once
#include <stdint.h>

#include "profile.h"

#include <math.h>
#include <string.h>

namespace sb {

// Characters per score check.  The score can change by at most 1 per text
// character, so if the score after a group of kGroup characters is above
// k + kGroup - 1 no position inside the group can be <= k.
#ifndef SB_GROUP
#define SB_GROUP 8
#endif
constexpr int kGroup = SB_GROUP;  // 4, 8 or 16 (whole 32-bit text words)
// Bytes per thread per pipeline stage: 64 (half a line, SWIZZLE_64B) or 128 (SWIZZLE_128B).
#ifndef SB_STAGE_BYTES
#define SB_STAGE_BYTES 64
#endif
constexpr int kStageBytes = SB_STAGE_BYTES;

constexpr int kMaxFilterWords = 8;
// Delay-line bits behind every piece of the prefilter: together with the piece's last
// bit they keep an occurrence visible for 4 characters = one text word = one hit test.
constexpr int kFilterDelay = 3;

// Geometry of one scan: the text is viewed as `rows` rows of `ltot` bytes.
struct ScanGeom {
  uint32_t ltot;    // bytes per row, multiple of kStageBytes
  uint32_t rows;    // ceil(n / ltot)
  uint32_t nstage;  // ltot / kStageBytes
  uint32_t nwarm;   // warm-up stages taken from the neighbouring row: ceil((m+k)/kStageBytes)
};

// Everything that is uniform over one scan launch.
struct ScanArgs {
  const uint8_t* text;  // padded device text (used by the non-TMA variant and the emulator)
  uint64_t n;           // text length in bytes
  uint64_t reset_idx;   // forward index at which the scan state restarts (0 fwd, n-1 rev)
  ScanGeom g;
  uint32_t sh0;   // text byte -> equality row: row = (byte >> sh0) & msk0 (applied to 4 packed bytes)
  uint32_t msk0;
  uint32_t nrows;  // equality rows per query (4 Dna, 32 Iupac, 256 Ascii)
  uint32_t rowbytes;  // W * 4
  int32_t m;       // pattern length
  int32_t k;       // threshold
  uint32_t nq;       // queries in this launch (scan2_kernel: pattern PAIRS)
  uint32_t nq_odd;   // scan2_kernel: 1 = the last pair's second pattern is a padding copy
  uint32_t qs_base;  // query slot of the first query (strand * n_patterns + pattern)
  const uint32_t* eq;  // [nq][nrows][W] equality words
  uint64_t* cand_keys;
  uint32_t* cand_cost;
  unsigned long long* cand_count;
  uint64_t cand_cap;
  // exact piece prefilter (filter_kernel / verify_kernel)
  const uint32_t* feq;      // [nq][256][WF] filter automaton masks, indexed by the raw text byte
  uint32_t finit[kMaxFilterWords];   // first bit of every piece
  uint32_t fdelay[kMaxFilterWords];  // delay-line bits behind every piece
  uint32_t rev_lead;        // strand-fused prefilter: extra window length for hits of reversed queries
  uint32_t fused;           // 1: automaton words [WF/2, WF) belong to the reversed partner query (slot + nq)
  uint64_t* hit_keys;       // (query slot << 40) | forward index of the hit's 16-byte text chunk / 16
  unsigned long long* hit_count;
  uint64_t hit_cap;
  uint64_t emit_min;        // overhang: end positions <= emit_min come from the edge kernel instead
  // q-gram bitmap prefilter (qgram_kernel): exact confirmation of a hit before its re-scan
  // exact refinement of prefilter hits (refine_hit): Dna only
  const uint32_t* qconf;    // [query slot][qnp][kConfWords] piece records, see PieceConf
  uint32_t qnp;             // pieces per query (k + 1); 0 = hits are not refined
  uint32_t hit_exact;       // 1: hit keys hold a nominal END POSITION (scan direction) instead of a 16-byte chunk
  const uint32_t* hit_span; // hit_exact: per entry, how many positions beyond the nominal one the entry covers
  uint32_t max_span;        // upper bound of hit_span (sizes the window staging of the wide re-scan kernel)
  uint32_t wide_few;        // launch_verify: both re-scan kernels run, the device-side entry count picks one
  // Regional fallback of the prefilter routes.  A tile = the kScanThreads rows of one block of the
  // scan geometry.  Tiles with so many hits that re-scanning their neighbourhoods would cost more
  // than scanning the tile (repeats, low-complexity sequence) are marked dense: their hits are
  // dropped and the scan kernels run over exactly those tiles (tile_list) instead.
  // First pass of a prefilter route: refine / verify do nothing when the prefilter produced more than
  // guard_limit hits (the host then runs the regional pass); guard_limit = 0: no guard.
  const unsigned long long* guard_count;
  unsigned long long guard_limit;
  uint64_t tile_bytes;         // text bytes per tile (kScanThreads * ltot of the SCAN geometry); 0 = no tiles
  const uint8_t* dense;        // refine / verify: [tile] != 0 -> skip the hit
  const uint32_t* tile_list;   // scan kernels: block b scans tile tile_list[b / nq] if b / nq < *tile_count
  const uint32_t* tile_count;
};

// May the hit (16-byte chunk at forward index `fwd_pos`) be dropped because the scan of dense tiles
// covers everything it stands for?  A hit speaks for end positions up to 16 + m + k characters away
// (to the right for forward queries, to the left for reversed ones), so it is dropped only when the
// tiles at both ends of that reach are dense (tiles are much longer than the reach: at most two are
// involved).  Hits near the border between a dense and a sparse tile are verified as usual.
SB_HD bool hit_in_dense_tile(const ScanArgs& a, uint64_t fwd_pos) {
  if (!a.dense || !a.tile_bytes || a.n == 0) return false;
  const uint64_t reach = (uint64_t)a.m + (uint64_t)a.k + 32;
  const uint64_t lo = fwd_pos > reach ? fwd_pos - reach : 0;
  uint64_t hi = fwd_pos + reach;
  if (hi > a.n - 1) hi = a.n - 1;
  return a.dense[lo / a.tile_bytes] != 0 && a.dense[hi / a.tile_bytes] != 0;
}

template <int W>
struct Lane {
  uint32_t pv[W];  // vertical +1 deltas of the current column (rows = bits, wildcard rows are 0)
  uint32_t mv[W];  // vertical -1 deltas
};

SB_HD int popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}

template <int W>
SB_HD void lane_reset(Lane<W>& s, int m) {
  const int pad = 32 * W - m;  // wildcard rows below the pattern
#pragma unroll
  for (int w = 0; w < W; w++) {
    const int lo = pad - 32 * w;  // first pattern bit inside word w
    s.pv[w] = lo <= 0 ? 0xFFFFFFFFu : (lo >= 32 ? 0u : (0xFFFFFFFFu << lo));
    s.mv[w] = 0;
  }
}

// D[m][i] of the current column = sum of the vertical deltas (D[0][i] = 0).
// POPC runs on its own pipe, so checking the score once per kGroup characters
// costs the saturated ALU pipe one add and one compare.
template <int W>
SB_HD int lane_score(const Lane<W>& s) {
  int v = 0;
#pragma unroll
  for (int w = 0; w < W; w++) v += popc32(s.pv[w]) - popc32(s.mv[w]);
  return v;
}

// Candidate keys: (query slot << 40) | end position.  Sorting the 64-bit keys
// orders candidates by query slot, then by end position.
constexpr int kPosBits = 40;
SB_HD uint64_t cand_key(uint32_t qs, uint64_t pos) { return ((uint64_t)qs << kPosBits) | pos; }
SB_HD uint32_t key_qs(uint64_t key) { return (uint32_t)(key >> kPosBits); }
SB_HD uint64_t key_pos(uint64_t key) { return key & ((1ull << kPosBits) - 1); }

SB_HD void emit_candidate(const ScanArgs& a, uint32_t qs, uint64_t pos, int score) {
#if defined(__CUDA_ARCH__)
  const unsigned long long i = atomicAdd(a.cand_count, 1ull);
#else
  const unsigned long long i = (*a.cand_count)++;
#endif
  if (i < a.cand_cap) {
    a.cand_keys[i] = cand_key(qs, pos);
    a.cand_cost[i] = (uint32_t)score;
  }
}

// One text character.  eq[w]: bit set where the pattern row matches the
// character (all ones on the wildcard rows).  Hyyro's formulation of Myers'
// recurrences with zero horizontal input at row 0 (top row of D is 0).
// On sm_100a this compiles to 7 LOP3 (ALU pipe) + 3 IMAD (add and the two
// shifts, FMA pipe) per word.
template <int W>
SB_HD void myers_step(Lane<W>& s, const uint32_t* __restrict__ eq) {
  uint32_t carry = 0, phc = 0, mhc = 0;
#pragma unroll
  for (int w = 0; w < W; w++) {
    const uint32_t pv = s.pv[w], mv = s.mv[w];
    const uint32_t x = eq[w] | mv;
    const uint32_t t = x & pv;
    uint32_t u;
    if (W == 1) {
      u = t + pv;
    } else {
      const uint64_t sum = (uint64_t)t + pv + carry;
      u = (uint32_t)sum;
      carry = (uint32_t)(sum >> 32);
    }
    const uint32_t d0 = (u ^ pv) | x;
    const uint32_t ph = mv | ~(d0 | pv);
    const uint32_t mh = pv & d0;
    const uint32_t ph1 = (ph << 1) | phc;
    const uint32_t mh1 = (mh << 1) | mhc;
    if (W > 1) {
      phc = ph >> 31;
      mhc = mh >> 31;
    }
    s.pv[w] = mh1 | ~(d0 | ph1);
    s.mv[w] = ph1 & d0;
  }
}

// A query's equality table [nrows][W].  On the device it lives in shared
// memory and is addressed by its 32-bit shared-window address.
struct EqTab {
  const uint32_t* p;
  uint32_t saddr;     // device only: shared-window address of p
  uint32_t rowbytes;  // W * 4, kept in a register so the row offset is an IMAD, not a shift/LEA
};

// Equality words of the text byte in lane b (0..3) of `pre` (4 packed row
// indices).  Device: PRMT (ALU pipe) extracts the row, IMAD (FMA pipe) forms
// the address, LDS fetches; the ALU pipe - the bottleneck of this kernel -
// spends one instruction per character here.
template <int W>
SB_HD void load_eq(uint32_t (&eq)[W], const EqTab& t, uint32_t pre, int b) {
#if defined(__CUDA_ARCH__)
  const uint32_t row = __byte_perm(pre, 0u, 0x4440u + (uint32_t)b);
  uint32_t addr;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(addr) : "r"(row), "r"(t.rowbytes), "r"(t.saddr));
  if (W % 4 == 0) {
#pragma unroll
    for (int w = 0; w < W; w += 4)
      asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
          : "=r"(eq[w]), "=r"(eq[w + 1]), "=r"(eq[w + 2]), "=r"(eq[w + 3])
          : "r"(addr + 4 * w));
  } else if (W % 2 == 0) {
#pragma unroll
    for (int w = 0; w < W; w += 2)
      asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(eq[w]), "=r"(eq[w + 1]) : "r"(addr + 4 * w));
  } else {
#pragma unroll
    for (int w = 0; w < W; w++) asm("ld.shared.u32 %0, [%1];" : "=r"(eq[w]) : "r"(addr + 4 * w));
  }
#else
  const uint32_t row = (pre >> (8 * b)) & 0xFFu;
  const uint32_t* p = t.p + row * W;
  for (int w = 0; w < W; w++) eq[w] = p[w];
#endif
}

#if defined(__CUDACC__)
#define SB_SLOW __host__ __device__ __noinline__
#else
#define SB_SLOW
#endif

// Exact (per character) path: used for the rare groups that may contain a
// position with score <= k, and for the chunk containing the restart index.
// Deliberately not inlined and by-value, so that the hot path keeps the lane
// in registers and stays small in the instruction cache.
template <int W, bool REV>
SB_SLOW Lane<W> slow_word(Lane<W> s, uint32_t x, uint64_t base_idx, const ScanArgs& a,
                          EqTab eqs, uint32_t qs, bool own) {
  const uint32_t pre = (x >> a.sh0) & a.msk0;
  for (int bb = 0; bb < 4; bb++) {
    const int b = REV ? 3 - bb : bb;
    const uint64_t idx = base_idx + (uint64_t)b;
    if (idx == a.reset_idx) lane_reset<W>(s, a.m);
    uint32_t eq[W];
    load_eq<W>(eq, eqs, pre, b);
    myers_step<W>(s, eq);
    const int score = lane_score<W>(s);
    if (score <= a.k && own && idx < a.n) {
      const uint64_t pos = REV ? a.n - idx : idx + 1;
      if (pos > a.emit_min) emit_candidate(a, qs, pos, score);
    }
  }
  return s;
}

// prev_score = score at the end of the previous group (the thread keeps it in a
// register).  With s0 the score before a group of G characters and sG after it, a
// position i in the group can only have score <= k if s0 - i <= k and sG - (G - i) <= k,
// i.e. only if s0 + sG <= 2k + G.  Otherwise the group is skipped with one add and one
// compare; else it is replayed exactly.  xs = the kGroup/4 text words of the group in
// forward order; base_idx = forward index of xs[0] byte 0.
template <int W, bool REV>
SB_HD void fast_group(Lane<W>& s, int& prev_score, const uint32_t* xs, uint64_t base_idx, const ScanArgs& a,
                      const EqTab& eqs, uint32_t qs, bool own) {
  constexpr int NW = kGroup / 4;
  const Lane<W> saved = s;
#pragma unroll
  for (int ww = 0; ww < NW; ww++) {
    const int w = REV ? NW - 1 - ww : ww;
    const uint32_t pre = (xs[w] >> a.sh0) & a.msk0;
#pragma unroll
    for (int bb = 0; bb < 4; bb++) {
      const int b = REV ? 3 - bb : bb;
      uint32_t eq[W];
      load_eq<W>(eq, eqs, pre, b);
      myers_step<W>(s, eq);
    }
  }
  const int score = lane_score<W>(s);
  if (prev_score + score <= 2 * a.k + kGroup) {
    s = saved;
    for (int ww = 0; ww < NW; ww++) {
      const int w = REV ? NW - 1 - ww : ww;
      s = slow_word<W, REV>(s, xs[w], base_idx + 4 * w, a, eqs, qs, own);
    }
  }
  prev_score = score;
}

// 16 text bytes x[0..3] (little endian, x[0] byte 0 = forward index base_idx).
// EXACT = every word through the per-character path (stage holding the restart index).
template <int W, bool REV, bool EXACT>
SB_HD void process16(Lane<W>& s, int& prev_score, const uint32_t (&x)[4], uint64_t base_idx, const ScanArgs& a,
                     const EqTab& eqs, uint32_t qs, bool own) {
  if (EXACT) {
    for (int ww = 0; ww < 4; ww++) {
      const int w = REV ? 3 - ww : ww;
      s = slow_word<W, REV>(s, x[w], base_idx + 4 * w, a, eqs, qs, own);
    }
    prev_score = lane_score<W>(s);
    return;
  }
  constexpr int NW = kGroup / 4;
  constexpr int NG = 4 / NW;
#pragma unroll
  for (int gg = 0; gg < NG; gg++) {
    const int g = REV ? NG - 1 - gg : gg;
    fast_group<W, REV>(s, prev_score, &x[g * NW], base_idx + 4 * NW * g, a, eqs, qs, own);
  }
}

// ---------------------------------------------------------------------------
// Two one-word patterns per thread (batches of patterns of at most 32 characters): the text bytes
// are fetched, the row is extracted and the table address is formed ONCE for both patterns -- the
// pair table holds {eq of pattern A, eq of pattern B} per text byte, so one 8-byte shared-memory
// load feeds two recurrences -- and the two dependency chains interleave.  The pair table is
// indexed by the RAW text byte (256 rows, expanded from the per-class tables when the block
// starts), which also removes the per-word class extraction.
struct EqPair {
  uint32_t x, y;  // equality words of pattern A and pattern B for one text byte
};
struct Lane2 {
  uint32_t pv[2], mv[2];
};

SB_HD void lane2_reset(Lane2& s, int m) {
  const int pad = 32 - m;
  s.pv[0] = s.pv[1] = pad <= 0 ? 0xFFFFFFFFu : (0xFFFFFFFFu << pad);
  s.mv[0] = s.mv[1] = 0;
}

SB_HD void myers_step2(Lane2& s, uint32_t eqa, uint32_t eqb) {
#pragma unroll
  for (int q = 0; q < 2; q++) {
    const uint32_t eq = q ? eqb : eqa;
    const uint32_t pv = s.pv[q], mv = s.mv[q];
    const uint32_t x = eq | mv;
    const uint32_t u = (x & pv) + pv;
    const uint32_t d0 = (u ^ pv) | x;
    const uint32_t ph = mv | ~(d0 | pv);
    const uint32_t mh = pv & d0;
    const uint32_t ph1 = ph << 1, mh1 = mh << 1;
    s.pv[q] = mh1 | ~(d0 | ph1);
    s.mv[q] = ph1 & d0;
  }
}

// pair = the 256-entry table {eqA, eqB} indexed by the raw text byte (shared memory on the device)
SB_HD void load_eq2(uint32_t& eqa, uint32_t& eqb, const EqPair* pair, uint32_t saddr, uint32_t x, int b) {
#if defined(__CUDA_ARCH__)
  (void)pair;
  // The ALU pipe (LOP3 + PRMT) is the bottleneck of this kernel (ncu: 90 % busy).  Extracting the
  // text byte with integer multiplies on the FMA pipe instead (x * 2^(24-8b), then the high word of
  // that times 2^8) was measured 3.4 % SLOWER (1.97 vs 2.04 T lane-steps/s on c3: IMAD.HI is not a
  // full-rate instruction), so the PRMT stays; -DSB_EQ2_IMAD builds the other form.
#ifdef SB_EQ2_IMAD
  uint32_t top = x;
  if (b < 3) asm("mul.lo.u32 %0, %1, %2;" : "=r"(top) : "r"(x), "r"(1u << (8 * (3 - b))));
  uint32_t row;
  asm("mul.hi.u32 %0, %1, 256;" : "=r"(row) : "r"(top));
#else
  const uint32_t row = __byte_perm(x, 0u, 0x4440u + (uint32_t)b);
#endif
  uint32_t addr;
  asm("mad.lo.u32 %0, %1, 8, %2;" : "=r"(addr) : "r"(row), "r"(saddr));
  asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(eqa), "=r"(eqb) : "r"(addr));
#else
  (void)saddr;
  const EqPair e = pair[(x >> (8 * b)) & 0xFFu];
  eqa = e.x, eqb = e.y;
#endif
}

// Exact path of one text word for both patterns (rare: a group that may hold a position <= k, or the
// stage with the restart index).  qs = slot of pattern A, B = qs + 1 (B is a padding copy when
// `has_b` is false: nothing is reported for it).
template <bool REV>
SB_SLOW Lane2 slow_word2(Lane2 s, uint32_t x, uint64_t base_idx, const ScanArgs& a, const EqPair* pair, uint32_t saddr,
                         uint32_t qs, bool has_b, bool own) {
  for (int bb = 0; bb < 4; bb++) {
    const int b = REV ? 3 - bb : bb;
    const uint64_t idx = base_idx + (uint64_t)b;
    if (idx == a.reset_idx) lane2_reset(s, a.m);
    uint32_t eqa, eqb;
    load_eq2(eqa, eqb, pair, saddr, x, b);
    myers_step2(s, eqa, eqb);
    if (own && idx < a.n) {
      const uint64_t pos = REV ? a.n - idx : idx + 1;
      if (pos > a.emit_min) {
        const int sa = popc32(s.pv[0]) - popc32(s.mv[0]);
        if (sa <= a.k) emit_candidate(a, qs, pos, sa);
        const int sb = popc32(s.pv[1]) - popc32(s.mv[1]);
        if (has_b && sb <= a.k) emit_candidate(a, qs + 1, pos, sb);
      }
    }
  }
  return s;
}

// kGroup characters (as fast_group): both scores are checked once per group.
template <bool REV>
SB_HD void fast_group2(Lane2& s, int& prev_a, int& prev_b, const uint32_t* xs, uint64_t base_idx, const ScanArgs& a,
                       const EqPair* pair, uint32_t saddr, uint32_t qs, bool has_b, bool own) {
  constexpr int NW = kGroup / 4;
  const Lane2 saved = s;
#pragma unroll
  for (int ww = 0; ww < NW; ww++) {
    const int w = REV ? NW - 1 - ww : ww;
#pragma unroll
    for (int bb = 0; bb < 4; bb++) {
      const int b = REV ? 3 - bb : bb;
      uint32_t eqa, eqb;
      load_eq2(eqa, eqb, pair, saddr, xs[w], b);
      myers_step2(s, eqa, eqb);
    }
  }
  const int sa = popc32(s.pv[0]) - popc32(s.mv[0]);
  const int sb = popc32(s.pv[1]) - popc32(s.mv[1]);
  const int lim = 2 * a.k + kGroup;
  if (prev_a + sa <= lim || prev_b + sb <= lim) {
    s = saved;
    for (int ww = 0; ww < NW; ww++) {
      const int w = REV ? NW - 1 - ww : ww;
      s = slow_word2<REV>(s, xs[w], base_idx + 4 * w, a, pair, saddr, qs, has_b, own);
    }
  }
  prev_a = sa;
  prev_b = sb;
}

template <bool REV, bool EXACT>
SB_HD void process16_2(Lane2& s, int& prev_a, int& prev_b, const uint32_t (&x)[4], uint64_t base_idx,
                       const ScanArgs& a, const EqPair* pair, uint32_t saddr, uint32_t qs, bool has_b, bool own) {
  if (EXACT) {
    for (int ww = 0; ww < 4; ww++) {
      const int w = REV ? 3 - ww : ww;
      s = slow_word2<REV>(s, x[w], base_idx + 4 * w, a, pair, saddr, qs, has_b, own);
    }
    prev_a = popc32(s.pv[0]) - popc32(s.mv[0]);
    prev_b = popc32(s.pv[1]) - popc32(s.mv[1]);
    return;
  }
  constexpr int NW = kGroup / 4;
  constexpr int NG = 4 / NW;
#pragma unroll
  for (int gg = 0; gg < NG; gg++) {
    const int g = REV ? NG - 1 - gg : gg;
    fast_group2<REV>(s, prev_a, prev_b, &x[g * NW], base_idx + 4 * NW * g, a, pair, saddr, qs, has_b, own);
  }
}

// ---------------------------------------------------------------------------
// Exact piece prefilter (pigeonhole).  k+1 pairwise disjoint pieces of the
// pattern are searched EXACTLY with a Shift-And automaton; an alignment with at
// most k edits leaves at least one piece intact, so every end position with
// cost <= k lies within m+k characters behind an exact piece occurrence.  Only
// those neighbourhoods are then scanned with the Myers recurrences
// (verify_hit).  The result is identical to the full scan; the reference uses
// the same idea of an exact prefilter in its v2 engine (suffix prefilter,
// src/pattern_tiling/general.rs:60-102,294-313), with a different filter.
//
// Automaton word layout: pieces are packed from bit 0, each followed by
// kFilterDelay (3) delay bits whose mask accepts every character.  The hit mask
// holds every piece's last bit and its delay bits, so an occurrence ending at
// any of the 4 characters of a text word is visible when the word has been
// consumed: one test per text word.  Per character: PRMT + LOP3 on the ALU pipe,
// 2 IMAD on the FMA pipe, 1 LDS.
template <int WF>
struct FLane {
  uint32_t st[WF];
  uint32_t init[WF];   // copies of ScanArgs::finit / fdelay kept in registers
  uint32_t delay[WF];
};

template <int WF>
SB_HD void flane_reset(FLane<WF>& s, const ScanArgs& a) {
#pragma unroll
  for (int w = 0; w < WF; w++) s.st[w] = 0, s.init[w] = a.finit[w], s.delay[w] = a.fdelay[w];
}

template <int WF>
SB_HD void load_feq(uint32_t (&eq)[WF], const EqTab& t, uint32_t x, int b) {
#if defined(__CUDA_ARCH__)
  const uint32_t row = __byte_perm(x, 0u, 0x4440u + (uint32_t)b);
  uint32_t addr;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(addr) : "r"(row), "r"(t.rowbytes), "r"(t.saddr));
  if (WF % 4 == 0) {
#pragma unroll
    for (int w = 0; w + 3 < WF; w += 4)
      asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
          : "=r"(eq[w]), "=r"(eq[w + 1 < WF ? w + 1 : 0]), "=r"(eq[w + 2 < WF ? w + 2 : 0]), "=r"(eq[w + 3 < WF ? w + 3 : 0])
          : "r"(addr + 4 * w));
  } else if (WF == 2) {
    asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(eq[0]), "=r"(eq[WF > 1 ? 1 : 0]) : "r"(addr));
  } else {
#pragma unroll
    for (int w = 0; w < WF; w++) asm("ld.shared.u32 %0, [%1];" : "=r"(eq[w]) : "r"(addr + 4 * w));
  }
#else
  const uint32_t row = (x >> (8 * b)) & 0xFFu;
  const uint32_t* p = t.p + row * WF;
  for (int w = 0; w < WF; w++) eq[w] = p[w];
#endif
}

// Per-warp staging queue for hits (shared memory on the device; absent on the host).
// A single global counter cannot absorb millions of atomics per millisecond, so hits
// are batched: one global atomic per flush of up to kHitQueueCap hits.
constexpr uint32_t kHitQueueCap = 128;
struct HitQueue {
  uint64_t* q;   // [kHitQueueCap], or nullptr: write straight to the global list
  uint32_t* n;
};

// A hit = a 16-byte text chunk in which some piece occurrence ends.
constexpr int kHitChars = 16;

// Reports the hit chunks of one 64-byte stage (bit c of `mask` = chunk c, forward order).
// Deliberately not inlined: hits are rare, and the ownership / text-end tests must not be
// hoisted into the per-character fast path.
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#endif
static void emit_stage_hits(const ScanArgs& a, HitQueue hq, uint32_t qs, uint64_t stage_idx, uint32_t mask,
                            bool own) {
  if (!own) return;
  for (int c = 0; c < kStageBytes / kHitChars; c++) {
    if (!((mask >> c) & 1u)) continue;
    const uint64_t base_idx = stage_idx + (uint64_t)(kHitChars * c);
    if (base_idx >= a.n) continue;
    const uint64_t key = cand_key(qs, base_idx / kHitChars);
#if defined(__CUDA_ARCH__)
    if (hq.q) {
      const uint32_t pos = atomicAdd(hq.n, 1u);
      if (pos < kHitQueueCap) {
        hq.q[pos] = key;
        continue;
      }
    }
    const unsigned long long i = atomicAdd(a.hit_count, 1ull);
#else
    (void)hq;
    const unsigned long long i = (*a.hit_count)++;
#endif
    if (i < a.hit_cap) a.hit_keys[i] = key;
  }
}

// 16 text bytes through the automaton.  Returns non-zero iff a piece occurrence ended
// inside the chunk (the hit mask is sampled after every text word, see above).
// acc[0] collects the hit bits of automaton words [0, max(1, WF/2)), acc[1] those of the rest
// (the reversed partner query when the two strands share one pass).
template <int WF, bool REV>
SB_HD void filter16(FLane<WF>& s, const uint32_t (&x)[4], const EqTab& feq, uint32_t (&acc)[2]) {
  constexpr int kHalf = WF > 1 ? WF / 2 : 1;
  acc[0] = acc[1] = 0;
#pragma unroll
  for (int ww = 0; ww < 4; ww++) {
    const int w4 = REV ? 3 - ww : ww;
#pragma unroll
    for (int bb = 0; bb < 4; bb++) {
      const int b = REV ? 3 - bb : bb;
      uint32_t eq[WF];
      load_feq<WF>(eq, feq, x[w4], b);
#pragma unroll
      for (int w = 0; w < WF; w++) s.st[w] = ((s.st[w] << 1) | s.init[w]) & eq[w];
    }
#pragma unroll
    for (int w = 0; w < WF; w++) acc[w < kHalf ? 0 : 1] |= s.st[w] & s.delay[w];
  }
}

// Two-characters-per-step form of the same automaton for the Dna profile (4 character
// classes -> 16 class pairs).  With E0, E1 the masks of two consecutive characters and
// I the piece-start mask,
//   s2 = ((((s << 1) | I) & E0) << 1 | I) & E1 = ((s << 2) & A) | B,
//   A = (E0 << 1) & E1,   B = (((I & E0) << 1) | I) & E1,
// so a pair costs PRMT + LDS (A and B of the pair) + IMAD.SHL + LOP3.  The pair table
// [16][2*WF] is indexed by first | second << 2 in scan order; the pair offsets of a text
// word come from two shift/mask operations on the whole word.
// Pair table layout: tab[word][pair] = {A, B}: one 128-byte sub-table per automaton word, so
// that a warp access touches at most 16 distinct 8-byte entries in 16 distinct bank pairs (no
// bank conflicts, one shared-memory wavefront per half warp).  Interleaving the words of an
// entry (16- or 32-byte entries) was measured 2.5x slower: conflicts + bytes per access.
template <int WF>
SB_HD void load_pair(uint32_t (&a)[WF], uint32_t (&b)[WF], const EqTab& t, uint32_t off) {
  // plain indexing: on the device t.p points into shared memory and the compiler folds the
  // table base into the LDS address (register + uniform register + immediate), no add
  const uint8_t* base = reinterpret_cast<const uint8_t*>(t.p) + off;
#pragma unroll
  for (int w = 0; w < WF; w++) {
#if defined(__CUDA_ARCH__)
    const uint2 v = *reinterpret_cast<const uint2*>(base + 128 * w);
    a[w] = v.x, b[w] = v.y;
#else
    uint32_t v[2];
    memcpy(v, base + 128 * w, 8);
    a[w] = v[0], b[w] = v[1];
#endif
  }
}

// Bits of the pair offset: class c = (byte >> 1) & 3 is moved to bits 3-4 of its byte
// ((x << 2) & 0x18), then first | second << 2 is assembled by one more shift, so that a byte
// of the result IS the offset pair * 8 inside a word's sub-table.
template <int WF, bool REV>
SB_HD void filter16_pair(FLane<WF>& s, const uint32_t (&x)[4], const EqTab& ftab, uint32_t (&acc)[2]) {
  constexpr int kHalf = WF > 1 ? WF / 2 : 1;
  acc[0] = acc[1] = 0;
#pragma unroll
  for (int ww = 0; ww < 4; ww++) {
    const int w4 = REV ? 3 - ww : ww;
    const uint32_t pre = (x[w4] << 2) & 0x18181818u;
    const uint32_t t = REV ? (pre | (pre << 10)) : (pre | (pre >> 6));
#pragma unroll
    for (int pp = 0; pp < 2; pp++) {
      // forward: pairs live in bytes 0 and 2; reverse: bytes 3 and 1
      const int byte = REV ? 3 - 2 * pp : 2 * pp;
#if defined(__CUDA_ARCH__)
      const uint32_t off = __byte_perm(t, 0u, 0x4440u + (uint32_t)byte);
#else
      const uint32_t off = (t >> (8 * byte)) & 0xFFu;
#endif
      uint32_t a[WF], b[WF];
      load_pair<WF>(a, b, ftab, off);
#pragma unroll
      for (int w = 0; w < WF; w++) s.st[w] = ((s.st[w] << 2) & a[w]) | b[w];
    }
#pragma unroll
    for (int w = 0; w < WF; w++) acc[w < kHalf ? 0 : 1] |= s.st[w] & s.delay[w];
  }
}

// ---------------------------------------------------------------------------
// q-gram bitmap prefilter (Dna profile, one pattern, both strands in one pass).
//
// Same pigeonhole argument as the piece automaton above -- an alignment with at most k edits
// leaves one of the k+1 shares of the pattern intact -- but the shares are found through a
// direct-addressed bitmap instead of an automaton whose width grows with k: the thread keeps the
// last 16 text characters as 2-bit classes in one register (newest character in the top bits),
// and every S characters (S = 4, 8 or 16: once per 1, 2 or 4 text words) tests ONE bit of a
// 4^Q-bit table (Q <= 8: 8 KB of shared memory) for the Q-gram that ends there.  The table holds,
// for every share F (forward orientation; the reverse-complement strand contributes the shares of
// the reverse complement) the Q-grams F[o, o+Q) for o = 0..S-1; an occurrence of F covers S
// consecutive Q-gram ends, one of which lies on the sampling grid (|F| >= Q + S - 1).  The cost per
// text character does not depend on k or m: 3 instructions per text word to update the window
// (LOP3, IMAD.HI, SHF) + 5 per sample (SHF, LOP3, LDS, SHF, LOP3).
// Expected hits on uniform text: (k+1) * strands / 4^Q per character whatever S; a hit is
// confirmed exactly (qgram_confirm: the first min(|F|, 16) characters of the share are compared
// at the 16 possible alignments) before its neighbourhood is re-scanned.
struct QLane {
  uint32_t w;     // classes of the last 16 characters, character i-back at bits [30-2i, 32-2i)
  uint32_t prev;  // IMAD.HI result of the previous text word: its low byte = that word's 4 classes
};

SB_HD uint32_t umulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

// 4 text bytes -> their classes ((c >> 1) & 3) as one byte, first character in the low bits, in the
// LOW byte of the result (bits above it hold partial products: callers shift them out).
SB_HD uint32_t pack4_classes(uint32_t x) {
  // classes sit at bits 1-2, 9-10, 17-18, 25-26; the multiplier moves them to bits 32-39 of the
  // 64-bit product (shifts 31, 25, 19, 13), no two partial products overlap
  return umulhi32(x & 0x06060606u, 0x82082000u);
}

SB_HD int first_bit(uint32_t x) {  // index of the lowest set bit, x != 0
#if defined(__CUDA_ARCH__)
  return __ffs((int)x) - 1;
#else
  return __builtin_ctz(x);
#endif
}

SB_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) {  // low word of (hi:lo) >> sh, sh in [0, 32)
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(lo, hi, sh);
#else
  return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
#endif
}

// 16 text bytes.  Returns a word whose bit 0 is set iff a sampled Q-gram ending inside the chunk is
// in the table (other bits are garbage).  tab = the bitmap (shared memory on the device).
template <int Q, int S>
SB_HD uint32_t qgram16(QLane& s, const uint32_t (&x)[4], const uint32_t* __restrict__ tab) {
  uint32_t acc = 0;
#pragma unroll
  for (int ww = 0; ww < 4; ww++) {
    const uint32_t hi = pack4_classes(x[ww]);
    s.w = funnel_r(s.w, hi, 8);
    if ((4 * (ww + 1)) % S == 0) {
      if (Q == 8) {
        // index = top 16 bits of the window: word = index >> 5, bit = index & 31 = the low 5 bits of
        // the previous word's class byte (a shift uses the low 5 bits of its amount register)
        acc |= tab[s.w >> 21] >> (s.prev & 31u);
      } else {
        const uint32_t idx = s.w >> (32 - 2 * Q);
        acc |= tab[idx >> 5] >> (idx & 31u);
      }
    }
    s.prev = hi;
  }
  return acc;
}

// Window state of a thread that starts in the middle of the text: the 16 characters before its
// first character (hist = 16 text bytes, little endian words in text order).
SB_HD void qlane_init(QLane& s, const uint32_t (&hist)[4]) {
  uint32_t w = 0, hi = 0;
#pragma unroll
  for (int ww = 0; ww < 4; ww++) {
    hi = pack4_classes(hist[ww]);
    w = funnel_r(w, hi, 8);
  }
  s.w = w;
  s.prev = hi;
}

// Exact refinement of a prefilter hit (Dna).  A hit only says "some share of the pattern may occur
// around this 16-byte chunk"; here the first min(|F|, 16) characters of every share F (forward
// orientation, 2-bit classes) are compared at the 16 alignments the hit allows, and every
// (share, alignment) that matches is turned into the NOMINAL END POSITION of the pattern,
//   E0 = start of F - offset of F in the pattern + m        (scan direction of the query slot):
// in an alignment with at most k edits that leaves this share intact, the pattern ends within k
// characters of E0.  The re-scan then covers end positions E0 - k .. E0 + k only (plus m + k
// characters of warm-up) instead of the 16 + m + k positions behind an unrefined hit.
// Which alignments a hit allows (start of F relative to the chunk, rel0 .. rel0 + 15):
//   q-gram bitmap: the sampled Q-gram ends in the chunk and starts < S characters into F: rel0 = 1 - Q;
//   piece automaton, forward pass: the LAST character of F lies in the chunk: rel0 = 1 - |F|;
//   piece automaton, right-to-left pass (reversed queries): the FIRST character of F does: rel0 = 0.
// A share occurrence is never missed (the tests force every route); a false match costs a re-scan.
constexpr int kConfWords = 5;  // code, mask, rel0 (int32), offset in the pattern, length
struct PieceConf {
  uint32_t code, mask;
  int32_t rel0;
  uint32_t off, len;
};

// Returns false when no share matches behind the hit; else [lo, hi] = the smallest and the largest
// nominal end position over all matching (share, alignment) pairs of this chunk: ONE refined entry
// per hit, whose re-scan reports end positions lo - k .. hi + k.  (Almost always lo == hi; on
// repetitive text, where every alignment of every share matches, the entry degrades to about the
// window of an unrefined hit instead of multiplying.)
SB_HD bool refine_hit(const ScanArgs& a, uint32_t qs, bool rev, uint64_t base, int64_t& lo, int64_t& hi) {
  bool found = false;
  lo = hi = 0;
  // characters [base - 32, base + 32) as 2-bit classes, first character in the low bits
  uint32_t c[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    uint32_t code = 0;
    const int64_t at = (int64_t)base - 32 + 16 * j;
    if (at >= 0) {
      uint32_t x[4];
#if defined(__CUDA_ARCH__)
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(a.text + at));
      x[0] = v.x, x[1] = v.y, x[2] = v.z, x[3] = v.w;
#else
      memcpy(x, a.text + at, 16);
#endif
#pragma unroll
      for (int ww = 0; ww < 4; ww++) code |= (pack4_classes(x[ww]) & 0xFFu) << (8 * ww);
    }
    c[j] = code;
  }
  const uint32_t* conf = a.qconf + (size_t)qs * a.qnp * kConfWords;
  // vals[i] = the 16 characters of alignment i (the alignment starts at character 32 + rel0 + i of the
  // window: 4 .. 47).  All shares of the q-gram route have the same rel0, so the 16 funnel shifts
  // are done once per hit and a share costs 16 compares (statically indexed registers throughout).
  uint32_t vals[16];
  int32_t vals_rel0 = 1 << 30;
  for (uint32_t p = 0; p < a.qnp; p++) {
    const uint32_t code = conf[p * kConfWords], mask = conf[p * kConfWords + 1];
    const int32_t rel0 = (int32_t)conf[p * kConfWords + 2];
    const uint32_t off = conf[p * kConfWords + 3], len = conf[p * kConfWords + 4];
    if (rel0 != vals_rel0) {
      vals_rel0 = rel0;
      const uint32_t o0 = (uint32_t)(32 + rel0);
      const uint32_t wi = o0 >> 4, sh = o0 & 15u;
      const uint32_t w0 = wi == 0 ? c[0] : (wi == 1 ? c[1] : c[2]);
      const uint32_t w1 = wi == 0 ? c[1] : (wi == 1 ? c[2] : c[3]);
      const uint32_t w2 = wi == 0 ? c[2] : (wi == 1 ? c[3] : 0u);
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const uint32_t x = sh + (uint32_t)i;  // 0 .. 30
        vals[i] = x < 16 ? funnel_r(w0, w1, 2 * x) : funnel_r(w1, w2, 2 * (x - 16));
      }
    }
    uint32_t cand = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) cand |= (uint32_t)(((vals[i] ^ code) & mask) == 0) << i;
    while (cand) {
      const int i = first_bit(cand);
      cand &= cand - 1;
      const int64_t start = (int64_t)base + rel0 + i;  // forward index of the first character of F
      if (start < 0 || start + (int64_t)len > (int64_t)a.n) continue;
      // An intact copy of the pattern matches with ALL its shares, on one diagonal, and every share
      // would contribute an entry with the same nominal end.  Keep the entry of the first share
      // only: skip when an earlier share (one that fits the 16 compared characters entirely, so
      // that its own hit and refinement are certain) also matches on this diagonal.
      {
        const int64_t diag = start - (rev ? (int64_t)a.m - (int64_t)off - (int64_t)len : (int64_t)off);
        bool covered = false;
        for (uint32_t p2 = 0; p2 < p && !covered; p2++) {
          const uint32_t len2 = conf[p2 * kConfWords + 4];
          if (len2 > 16) continue;
          const uint32_t off2 = conf[p2 * kConfWords + 3];
          const int64_t s2 = diag + (rev ? (int64_t)a.m - (int64_t)off2 - (int64_t)len2 : (int64_t)off2);
          if (s2 < 0 || s2 + (int64_t)len2 > (int64_t)a.n) continue;
          if (hit_in_dense_tile(a, (uint64_t)s2)) continue;  // that share's own hit may have been dropped
          uint32_t v2 = 0;
          for (uint32_t j = 0; j < len2; j++) v2 |= (((uint32_t)a.text[s2 + j] >> 1) & 3u) << (2 * j);
          covered = v2 == conf[p2 * kConfWords];
        }
        if (covered) continue;
      }
      // nominal end position of the whole pattern in the slot's scan direction
      const int64_t e0 = rev ? (int64_t)a.n - start - (int64_t)len + ((int64_t)a.m - (int64_t)off)
                             : start - (int64_t)off + (int64_t)a.m;
      if (!found || e0 < lo) lo = e0;
      if (!found || e0 > hi) hi = e0;
      found = true;
    }
  }
  if (!found || hi + a.k < 1) return false;
  if (lo < 0) lo = 0;
  if (hi < lo) hi = lo;
  return true;
}

// Re-scan of the neighbourhood of one hit with the exact recurrences.  The hit
// is the text chunk at forward index 16*unit; in scan direction it starts at G0.
// A piece occurrence ending inside the chunk implies end positions in
// (G0, G0 + 16 + m + k]; starting m+k characters before G0 makes them exact.
// Window of a refined hit (nominal end position e0): scan [w0, end), report end positions > emit_from.
SB_HD void hit_window_exact(const ScanArgs& a, uint64_t e0, uint32_t span, int64_t& w0, int64_t& end,
                            int64_t& emit_from) {
  const int64_t n = (int64_t)a.n;
  emit_from = (int64_t)e0 - (int64_t)a.k - 1;
  if (emit_from < 0) emit_from = 0;
  w0 = emit_from - ((int64_t)a.m + (int64_t)a.k);
  if (w0 < 0) w0 = 0;
  end = (int64_t)e0 + (int64_t)span + (int64_t)a.k;
  if (end > n) end = n;
}

template <int W>
SB_HD void verify_hit(const ScanArgs& a, const uint32_t* eq /*[nrows][W] of this query*/, uint32_t qs, bool rev,
                      uint64_t unit, uint32_t span = 0) {
  const int64_t n = (int64_t)a.n;
  const int64_t reach = (int64_t)a.m + (int64_t)a.k;
  int64_t w0, end, emit_from;
  if (a.hit_exact) {  // refined hit: `unit` is the nominal end position E0; end positions E0 - k .. E0 + span + k
    hit_window_exact(a, unit, span, w0, end, emit_from);
  } else {
    const int64_t base = (int64_t)(unit * kHitChars);
    const int64_t g0 = rev ? n - kHitChars - base : base;  // may be negative for the chunk straddling the text end
    w0 = g0 - reach;
    if (w0 < 0) w0 = 0;
    // strand-fused prefilter: a reversed query's hit marks where its piece STARTS in scan direction
    end = g0 + kHitChars + reach + (rev ? (int64_t)a.rev_lead : 0);
    if (end > n) end = n;
    emit_from = g0 < 0 ? 0 : g0;
  }
  Lane<W> s;
  lane_reset<W>(s, a.m);
  for (int64_t idx = w0; idx < end; idx++) {
    const uint8_t c = a.text[rev ? n - 1 - idx : idx];
    const uint32_t row = ((uint32_t)c >> a.sh0) & (a.msk0 & 0xFFu);
    myers_step<W>(s, eq + row * W);
    if (idx >= emit_from) {
      const int score = lane_score<W>(s);
      if (score <= a.k) emit_candidate(a, qs, (uint64_t)idx + 1, score);
    }
  }
}

// Concatenated texts (offs ascending, every text followed by padding): the text that holds the
// end position `pos` (scan direction over `total` bytes) of a candidate, and the end position
// inside that text, again in scan direction.  False for positions in the padding.
SB_HD bool concat_locate(uint64_t pos, bool rev, uint64_t total, const uint64_t* offs, const uint64_t* lens,
                         uint32_t ntexts, uint32_t& ti, uint64_t& local) {
  if (pos == 0 || pos > total || ntexts == 0) return false;
  const uint64_t g = rev ? total - pos : pos - 1;  // forward index of the last character consumed
  uint32_t lo = 0, hi = ntexts;                    // last text with offs <= g
  while (hi - lo > 1) {
    const uint32_t mid = lo + (hi - lo) / 2;
    if (offs[mid] <= g) lo = mid; else hi = mid;
  }
  if (offs[lo] > g || g >= offs[lo] + lens[lo]) return false;
  ti = lo;
  local = rev ? offs[lo] + lens[lo] - g : g - offs[lo] + 1;
  return true;
}

// True when the stage starting at forward index stage_idx holds the restart index.
SB_HD bool stage_is_special(const ScanArgs& a, uint64_t stage_idx) {
  return (a.reset_idx - stage_idx) < (uint64_t)kStageBytes;  // unsigned wrap-around intended
}

// Which 128-byte piece a thread reads in pipeline iteration `it`
// (it < nwarm: warm-up from the neighbouring row; else own row).
template <bool REV>
SB_HD void stage_coord(const ScanGeom& g, uint32_t it, int64_t row, int64_t& r, uint32_t& col, bool& own) {
  own = it >= g.nwarm;
  if (!REV) {
    if (!own) {
      r = row - 1;
      col = g.ltot - (g.nwarm - it) * kStageBytes;
    } else {
      r = row;
      col = (it - g.nwarm) * kStageBytes;
    }
  } else {
    if (!own) {
      r = row + 1;
      col = (g.nwarm - 1 - it) * kStageBytes;
    } else {
      r = row;
      col = g.ltot - (it - g.nwarm + 1) * kStageBytes;
    }
  }
}

// ---------------------------------------------------------------------------
// ---------------------------------------------------------------------------
// Selection on the sorted candidate list (all positions with cost<=k; a position may be
// listed several times when re-scan windows of the prefilter overlap -- copies carry the
// same cost and only the first copy can be selected).
// Local-minima rule: a maximal run = same query slot, consecutive positions.  A candidate
// is kept iff the next step is up (or the run ends) and the last non-equal step before it
// inside the run was down (or the run starts).  This is the reference's streaming rule
// (src/search.rs:1344-1368) restricted to runs, identical to
// src/pattern_tiling/minima.rs:9-52.
SB_HD bool select_candidate(const uint64_t* keys, const uint32_t* cost, uint64_t i, uint64_t n, bool all_minima) {
  const uint64_t key = keys[i];
  if (i > 0 && keys[i - 1] == key) return false;  // a later copy
  if (all_minima) return true;
  const uint32_t c = cost[i];
  uint64_t j = i + 1;
  while (j < n && keys[j] == key) j++;  // next distinct candidate
  if (j < n && keys[j] == key + 1 && cost[j] <= c) return false;
  uint64_t p = i;  // first copy of the position we stand on
  uint64_t kk = key;
  while (p > 0 && keys[p - 1] == kk - 1) {
    p--;
    kk--;
    if (cost[p] != c) return cost[p] > c;
    while (p > 0 && keys[p - 1] == kk) p--;  // move to the first copy
  }
  return true;
}

// ---------------------------------------------------------------------------
// End-position predicates evaluated together with the selection, in the reference's order
// (src/search.rs:895-919): the caller's end filter, then the N end-point filter.  The
// reference takes an arbitrary closure (search_with_fn, :767-784); a device kernel cannot,
// so the one closure the reference itself ships is built in: "the last pam_len characters
// before the end position match the PAM exactly" (bin/crispr.rs:136-143,198-205), with the
// profile's is_match and, for reversed queries, the complemented PAM on the reversed text.
// Several texts (search_texts / search_many): query slot = text index * nq_per_text + query.
struct TextRef {
  const uint8_t* base;      // one text, or the concatenation of all texts
  uint64_t n;               // length of the single text
  const uint64_t* offs;     // nullptr for a single text, else start of every text in `base`
  const uint64_t* lens;
  uint32_t nq_per_text;
};

SB_HD void text_of_slot(const TextRef& t, uint32_t qs, const uint8_t*& text, uint64_t& n, uint32_t& q) {
  if (t.offs) {
    const uint32_t ti = qs / t.nq_per_text;
    q = qs - ti * t.nq_per_text;
    text = t.base + t.offs[ti];
    n = t.lens[ti];
  } else {
    q = qs;
    text = t.base;
    n = t.n;
  }
}

constexpr int kMaxPam = 16;
struct EndFilter {
  TextRef text;
  const uint8_t* rev_flags;  // per query (not per slot)
  int32_t profile;
  int32_t m, k;
  int32_t pam_len;           // 0: no end filter
  uint8_t pam[2][kMaxPam];   // [0] forward queries, [1] reversed queries (complemented, not reversed)
  int32_t n_endpoint;        // 1: N end-point filter (v1 only, src/n_filter.rs:41-53)
  float max_n_frac;
};

SB_HD bool profile_is_match(int profile, uint8_t a, uint8_t b) {
  switch (profile) {
    case kDna: return trace_match<kDna>(a, b);
    case kIupac: return trace_match<kIupac>(a, b);
    default: return trace_match<kAscii>(a, b);
  }
}

// count(N or n in td[start, end)) / denom <= max_n_frac, in f32 like src/n_filter.rs:8-34.
SB_HD bool n_fraction_ok(const uint8_t* text, uint64_t n, bool rev, uint64_t start, uint64_t end, float max_n_frac,
                         uint64_t denom /*0 = slice length*/) {
  if (start >= n) return true;
  if (end <= start) return true;
  uint64_t cnt = 0;
  for (uint64_t i = start; i < end; i++) {
    const uint8_t c = rev ? text[n - 1 - i] : text[i];
    cnt += (c | 0x20) == 'n';
  }
  const float frac = (float)cnt / (float)(denom ? denom : end - start);
  return frac <= max_n_frac;
}

SB_HD bool end_filter_pass(const EndFilter& f, uint64_t key) {
  const uint8_t* text;
  uint64_t n;
  uint32_t q;
  text_of_slot(f.text, key_qs(key), text, n, q);
  const bool rev = f.rev_flags[q] != 0;
  const uint64_t end = key_pos(key);
  if (f.pam_len > 0) {
    // the reference's closure indexes text[len - pam_len..] and panics when the prefix is
    // shorter than the PAM; such an end position cannot hold the PAM: rejected
    if (end < (uint64_t)f.pam_len) return false;
    const uint8_t* pam = f.pam[rev ? 1 : 0];
    for (int i = 0; i < f.pam_len; i++) {
      const uint64_t idx = end - (uint64_t)f.pam_len + (uint64_t)i;
      const uint8_t c = rev ? text[n - 1 - idx] : text[idx];
      if (!profile_is_match(f.profile, c, pam[i])) return false;
    }
  }
  if (f.n_endpoint) {
    const uint64_t e = end < n ? end : n;
    const uint64_t mand = f.m > f.k ? (uint64_t)(f.m - f.k) : 0;
    const uint64_t s = e > mand ? e - mand : 0;
    if (!n_fraction_ok(text, n, rev, s, e, f.max_n_frac, (uint64_t)(f.m + f.k))) return false;
  }
  return true;
}

// ---------------------------------------------------------------------------
// Traceback of one match.  Window = the (m+k) text characters before the end
// position, recomputed with the same recurrences while storing every column's
// vertical deltas (reference src/search.rs:1477-1478, src/trace.rs:57-104;
// column lookup as src/pattern_tiling/trace.rs:85-116), then the greedy walk
// of src/trace.rs:314-388 ('=' > 'X' > 'D' (text only) > 'I' (pattern only)).
struct TraceOut {
  uint64_t text_start;  // in scan direction (reversed-text coordinates for REV queries)
  uint64_t text_end;
  int32_t cost;
  uint32_t nops;
  uint32_t failed;
};

// Device-side match record (scan-direction coordinates).
struct GpuMatch {
  uint64_t text_start;  // scan-direction coordinates
  uint64_t text_end;
  uint32_t qs;
  int32_t cost;
  uint32_t nops;
  uint32_t failed;  // bit 0: traceback failed; bit 1: dropped by the traced N-fraction filter
};

enum : uint32_t { kOpEq = 0, kOpX = 1, kOpI = 2, kOpD = 3 };

// Column store accessor: word w of column i for this match.
struct ColStore {
  uint32_t* base;
  uint64_t stride;  // distance between consecutive (column, word) slots
  SB_HD uint32_t& at(uint32_t slot) const { return base[(uint64_t)slot * stride]; }
  // hint: the line holding `slot` will be read a few steps from now (device only)
  SB_HD void prefetch(uint32_t slot) const {
#if defined(__CUDA_ARCH__)
    // generic address: a column store in shared memory makes this a no-op
    asm volatile("prefetch.L1 [%0];" ::"l"(base + (uint64_t)slot * stride));
#else
    (void)slot;
#endif
  }
};

SB_HD uint8_t text_at_dir(const uint8_t* text, uint64_t n, bool rev, uint64_t i) {
  return rev ? text[n - 1 - i] : text[i];
}

// Column store of the traceback: per column i (0..m+k) and word w the vertical deltas (pv, mv);
// for patterns of more than 4 words also the horizontal deltas (ph, mh) of the step that produced
// the column, so that the greedy walk looks its three neighbours up with single-bit reads instead
// of prefix popcounts over up to 32 words (a 1000-character pattern walks 1000 steps).
// (from 3 words on: a 100-character pattern walks ~100 steps, each of which would otherwise read and
// popcount up to 3 x W x 2 words of the column store)
SB_HD int trace_fields(int W) { return W > 2 ? 4 : 2; }
SB_HD uint64_t trace_words_per_match(int m, int k, int W) {
  return (uint64_t)(m + k + 1) * (uint64_t)W * (uint64_t)trace_fields(W);
}

// +1 / -1 / 0 from a (plus, minus) pair of delta words at bit b
SB_HD int delta_at(uint32_t p, uint32_t mn, int bit) { return (int)((p >> bit) & 1u) - (int)((mn >> bit) & 1u); }

// The greedy walk of src/trace.rs:314-388 over a filled column store (columns 0..wlen of the
// window starting at `off`): '=' > 'X' > 'D' (text only) > 'I' (pattern only).
// `win`: the window's characters in scan order when the caller has them at hand (shared memory),
// else nullptr (they are read from the text).
template <int P>
SB_HD void trace_walk(const uint8_t* text, uint64_t n, bool rev, const uint8_t* pattern, int m, int W, uint64_t off,
                      uint32_t wlen, uint64_t end, const ColStore& cs, uint32_t* ops_out, uint32_t ops_words,
                      TraceOut& out, const uint8_t* win = nullptr) {
  const int pad = 32 * W - m;
  const int F = trace_fields(W);
  const bool wide = F == 4;
  const bool unit = cs.stride == 1;  // the warp-per-match kernels: the fields of a word are 16 adjacent bytes
  // The op codes are built in a local buffer and written out once: `ops_out` may be pinned host
  // memory (the single-synchronisation tail writes results over PCIe), where every
  // read-modify-write of a word would cost a round trip.  The walk produces the ops back to front
  // (src/trace.rs:393 reverses them): they are written downwards from the top of the buffer, 16
  // per word through a register, and moved to the front with one funnel shift per word at the end.
  constexpr uint32_t kLocalOpsWords = 80;  // 1280 ops: any m + k up to 1024 + 255
  uint32_t local_ops[kLocalOpsWords];
  uint32_t* const ops = ops_words <= kLocalOpsWords ? local_ops : ops_out;
  // D[j][i] for the narrow layout: column 0 is j, row 0 is 0, else the sum of the vertical deltas
  auto cost = [&](int j, uint32_t i) -> int {
    if (j == 0) return 0;
    if (i == 0) return j;
    int bits = pad + j, v = 0;
    for (int w = 0; w < W && bits > 0; w++) {
      const uint32_t msk = bits >= 32 ? 0xFFFFFFFFu : ((1u << bits) - 1u);
      v += popc32(cs.at((i * W + w) * F) & msk) - popc32(cs.at((i * W + w) * F + 1) & msk);
      bits -= 32;
    }
    return v;
  };
  int j = m;
  uint32_t i = wlen;
  int g = cost(j, i);
  out.cost = g;
  out.failed = 0;
  uint32_t nops = 0;
  const uint32_t max_ops = ops_words * 16;
  uint32_t cur = 0;  // the word that holds op number max_ops - 1 - nops
  while (j > 0) {
    int diag, left, up;
    if (wide) {
      // neighbours of (j, i) with value g: single-bit reads of word (pad + j - 1) / 32
      //   up    D[j-1][i]   = g - v(j, i)
      //   left  D[j][i-1]   = g - h(j, i)
      //   diag  D[j-1][i-1] = left - v(j, i-1)
      const int b = pad + j - 1;
      const int bit = b & 31;
      const uint32_t slot = (i * (uint32_t)W + (uint32_t)(b >> 5)) * 4u;
      uint32_t pv_i, mv_i, ph_i = 0, mh_i = 0, pv_p = 0, mv_p = 0;
#if defined(__CUDA_ARCH__)
      if (unit) {
        // the walk moves one column per step at most: fetch the words around the path 6 columns ahead
        if (i >= 6) cs.prefetch(slot - 6u * (uint32_t)W * 4u);
        const uint4 f = *reinterpret_cast<const uint4*>(cs.base + slot);
        pv_i = f.x, mv_i = f.y, ph_i = f.z, mh_i = f.w;
        if (i > 0) {
          const uint2 f2 = *reinterpret_cast<const uint2*>(cs.base + (slot - (uint32_t)W * 4u));
          pv_p = f2.x, mv_p = f2.y;
        }
      } else
#endif
      {
        (void)unit;
        pv_i = cs.at(slot), mv_i = cs.at(slot + 1);
        if (i > 0) {
          ph_i = cs.at(slot + 2), mh_i = cs.at(slot + 3);
          pv_p = cs.at(slot - (uint32_t)W * 4u), mv_p = cs.at(slot - (uint32_t)W * 4u + 1);
        }
      }
      up = g - delta_at(pv_i, mv_i, bit);
      if (i > 0) {
        left = g - delta_at(ph_i, mh_i, bit);
        diag = left - delta_at(pv_p, mv_p, bit);
      } else {
        left = diag = 0;
      }
    } else {
      up = cost(j - 1, i);
      left = i > 0 ? cost(j, i - 1) : 0;
      diag = i > 0 ? cost(j - 1, i - 1) : 0;
    }
    uint32_t op;
    const uint8_t tc = i > 0 ? (win ? win[i - 1] : text_at_dir(text, n, rev, off + i - 1)) : (uint8_t)0;
    if (i > 0 && diag == g && trace_match<P>(pattern[j - 1], tc)) {
      op = kOpEq;
      j--, i--;
    } else {
      g -= 1;
      if (i > 0 && diag == g) {
        op = kOpX;
        j--, i--;
      } else if (i > 0 && left == g) {
        op = kOpD;
        i--;
      } else if (up == g) {
        op = kOpI;
        j--;
      } else {
        out.failed = 1;  // the reference panics with "Trace failed" (src/trace.rs:367-387)
        break;
      }
    }
    if (nops >= max_ops) {  // more ops than the record holds: cannot happen for cost <= k
      out.failed = 1;
      break;
    }
    const uint32_t pos = max_ops - 1 - nops;
    cur |= op << ((pos & 15u) * 2);
    if ((pos & 15u) == 0) {
      ops[pos >> 4] = cur;
      cur = 0;
    }
    nops++;
  }
  // ops[first_pos ..] hold the ops in pattern direction: move them to position 0
  const uint32_t first_pos = max_ops - nops;
  if (first_pos & 15u) ops[first_pos >> 4] = cur;  // the partly filled word
  const uint32_t w0 = first_pos >> 4, sh = (first_pos & 15u) * 2;
  const uint32_t nwords = (nops + 15) / 16;
  for (uint32_t w = 0; w < ops_words; w++) {
    uint32_t v = 0;
    if (w < nwords) {
      const uint32_t lo = ops[w0 + w];
      const uint32_t hi = (w0 + w + 1 < ops_words) ? ops[w0 + w + 1] : 0u;
      v = sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
      const uint32_t valid = nops - w * 16;  // ops in this word
      if (valid < 16) v &= (1u << (2 * valid)) - 1u;
    }
    ops_out[w] = v;
  }
  out.text_start = off + i;
  out.text_end = end;
  out.nops = nops;
}

// `ops` receives 2-bit op codes, 16 per word, in pattern direction.
template <int P>
SB_HD void trace_one(const uint8_t* text, uint64_t n, bool rev, const uint8_t* pattern, int m, int k,
                     const uint32_t* eq /*[rows][W]*/, int W, uint32_t sh0, uint32_t msk0,
                     uint64_t end, const ColStore& cs, uint32_t* ops, uint32_t ops_words, TraceOut& out) {
  const int pad = 32 * W - m;
  const int F = trace_fields(W);
  const bool wide = F == 4;
  const uint64_t fill = (uint64_t)m + (uint64_t)k;
  const uint64_t off = end > fill ? end - fill : 0;
  const uint32_t wlen = (uint32_t)(end - off);
  // previous column: registers for the first 4 words, a local array beyond
  // (the device runs this one-thread fill up to 32 words only; the emulation library up to 128)
#if defined(__CUDACC__)
  constexpr int kRegW = 4, kMaxW = 32;
#else
  constexpr int kRegW = 4, kMaxW = 128;
#endif
  uint32_t ppv[kRegW], pmv[kRegW];
  uint32_t lpv[kMaxW - kRegW], lmv[kMaxW - kRegW];
  for (int w = 0; w < W; w++) {
    const int lo = pad - 32 * w;
    const uint32_t v0 = lo <= 0 ? 0xFFFFFFFFu : (lo >= 32 ? 0u : (0xFFFFFFFFu << lo));
    if (w < kRegW)
      ppv[w] = v0, pmv[w] = 0;
    else
      lpv[w - kRegW] = v0, lmv[w - kRegW] = 0;
    cs.at((0 * W + w) * F) = v0;  // column 0: D[j][0] = j
    cs.at((0 * W + w) * F + 1) = 0;
  }
  for (uint32_t i = 1; i <= wlen; i++) {
    const uint8_t tc = text_at_dir(text, n, rev, off + i - 1);
    const uint32_t row = ((uint32_t)tc >> sh0) & (msk0 & 0xFFu);
    const uint32_t* e = eq + row * W;
    uint32_t carry = 0, phc = 0, mhc = 0;
#pragma unroll
    for (int w = 0; w < kRegW; w++) {
      if (w >= W) break;
      const uint32_t pv = ppv[w], mv = pmv[w];
      const uint32_t x = e[w] | mv;
      const uint32_t t = x & pv;
      const uint64_t sum = (uint64_t)t + pv + carry;
      const uint32_t u = (uint32_t)sum;
      carry = (uint32_t)(sum >> 32);
      const uint32_t d0 = (u ^ pv) | x;
      const uint32_t ph = mv | ~(d0 | pv);
      const uint32_t mh = pv & d0;
      const uint32_t ph1 = (ph << 1) | phc;
      const uint32_t mh1 = (mh << 1) | mhc;
      phc = ph >> 31;
      mhc = mh >> 31;
      ppv[w] = mh1 | ~(d0 | ph1);
      pmv[w] = ph1 & d0;
      cs.at((i * W + w) * F) = ppv[w];
      cs.at((i * W + w) * F + 1) = pmv[w];
      if (wide) cs.at((i * W + w) * F + 2) = ph, cs.at((i * W + w) * F + 3) = mh;
    }
    for (int w = kRegW; w < W; w++) {
      const uint32_t pv = lpv[w - kRegW], mv = lmv[w - kRegW];
      const uint32_t x = e[w] | mv;
      const uint32_t t = x & pv;
      const uint64_t sum = (uint64_t)t + pv + carry;
      const uint32_t u = (uint32_t)sum;
      carry = (uint32_t)(sum >> 32);
      const uint32_t d0 = (u ^ pv) | x;
      const uint32_t ph = mv | ~(d0 | pv);
      const uint32_t mh = pv & d0;
      const uint32_t ph1 = (ph << 1) | phc;
      const uint32_t mh1 = (mh << 1) | mhc;
      phc = ph >> 31;
      mhc = mh >> 31;
      const uint32_t npv = mh1 | ~(d0 | ph1), nmv = ph1 & d0;
      lpv[w - kRegW] = npv, lmv[w - kRegW] = nmv;
      cs.at((i * W + w) * F) = npv;
      cs.at((i * W + w) * F + 1) = nmv;
      cs.at((i * W + w) * F + 2) = ph;  // W > 4: always the wide layout
      cs.at((i * W + w) * F + 3) = mh;
    }
  }
  trace_walk<P>(text, n, rev, pattern, m, W, off, wlen, end, cs, ops, ops_words, out);
}

// ---------------------------------------------------------------------------
// Overhang (reference src/search.rs:347-356,1274-1282,1693-1710; src/trace.rs:36-53,298-334):
// pattern characters hanging over either end of the text cost alpha each.  The left column of D
// is L(j) = floor(min(j, mo) * alpha) + (j - min(j, mo)), the text is followed by `steps`
// wildcard characters, and an end position o characters beyond the text costs floor(alpha * o)
// on top of the DP value.  f32 arithmetic as in the reference.
SB_HD int overhang_left_cost(int j, float alpha, int mo) {
  const int jm = (mo >= 0 && mo < j) ? mo : j;
  return (int)floorf((float)jm * alpha) + (j - jm);
}
SB_HD int overhang_overshoot_cost(float alpha, uint32_t o) { return o == 0 ? 0 : (int)floorf(alpha * (float)o); }

// Record flags (GpuMatch::failed / TraceOut::failed): bit 0 traceback failed, bit 1 dropped by the
// traced N filter, bits 8-19 pattern_start (left overhang), bits 20-31 right overshoot
// (pattern_end = m - overshoot).
SB_HD uint32_t pack_overhang(uint32_t pattern_start, uint32_t overshoot) {
  return (pattern_start << 8) | (overshoot << 20);
}

// trace_one with overhang: the window has m + k columns whatever the end position, the left
// column is L(j), columns beyond the text are wildcards (src/trace.rs:57-104); an end beyond the
// text first steps diagonally back into it, reaching column 0 with j rows left is a left overhang.
template <int P>
SB_HD void trace_one_ov(const uint8_t* text, uint64_t n, bool rev, const uint8_t* pattern, int m, int k,
                        const uint32_t* eq /*[rows][W]*/, int W, uint32_t sh0, uint32_t msk0, uint64_t end,
                        float alpha, int mo, const ColStore& cs, uint32_t* ops, uint32_t ops_words, TraceOut& out) {
  const int pad = 32 * W - m;
  const uint64_t fill = (uint64_t)m + (uint64_t)k;
  const uint64_t off = end > fill ? end - fill : 0;
  const uint32_t slice = (uint32_t)((end < n ? end : n) - off);
  const uint32_t wlen = (uint32_t)(end - off);
  const uint32_t cols = (uint32_t)fill > wlen ? (uint32_t)fill : wlen;
  // column 0: vertical deltas of L(j)
  for (int w = 0; w < W; w++) cs.at((0 * W + w) * 2) = 0, cs.at((0 * W + w) * 2 + 1) = 0;
  for (int j = 0; j < m; j++)
    if (overhang_left_cost(j + 1, alpha, mo) - overhang_left_cost(j, alpha, mo)) {
      const int b = pad + j;
      cs.at((0 * W + (b >> 5)) * 2) |= 1u << (b & 31);
    }
  for (uint32_t i = 1; i <= cols; i++) {
    const bool wild = i > slice;
    const uint8_t tc = wild ? 0 : text_at_dir(text, n, rev, off + i - 1);
    const uint32_t row = ((uint32_t)tc >> sh0) & (msk0 & 0xFFu);
    const uint32_t* e = eq + row * W;
    uint32_t carry = 0, phc = 0, mhc = 0;
    for (int w = 0; w < W; w++) {
      const uint32_t pv = cs.at(((i - 1) * W + w) * 2), mv = cs.at(((i - 1) * W + w) * 2 + 1);
      const uint32_t x = (wild ? 0xFFFFFFFFu : e[w]) | mv;
      const uint32_t t = x & pv;
      const uint64_t sum = (uint64_t)t + pv + carry;
      const uint32_t u = (uint32_t)sum;
      carry = (uint32_t)(sum >> 32);
      const uint32_t d0 = (u ^ pv) | x;
      const uint32_t ph = mv | ~(d0 | pv);
      const uint32_t mh = pv & d0;
      const uint32_t ph1 = (ph << 1) | phc;
      const uint32_t mh1 = (mh << 1) | mhc;
      phc = ph >> 31;
      mhc = mh >> 31;
      cs.at((i * W + w) * 2) = mh1 | ~(d0 | ph1);
      cs.at((i * W + w) * 2 + 1) = ph1 & d0;
    }
  }
  // D[j][i]: row 0 is 0; otherwise the sum of the column's vertical deltas of rows 1..j
  auto cost = [&](int j, uint32_t i) -> int {
    if (j == 0) return 0;
    int bits = pad + j, v = 0;
    for (int w = 0; w < W && bits > 0; w++) {
      const uint32_t msk = bits >= 32 ? 0xFFFFFFFFu : ((1u << bits) - 1u);
      v += popc32(cs.at((i * W + w) * 2) & msk) - popc32(cs.at((i * W + w) * 2 + 1) & msk);
      bits -= 32;
    }
    return v;
  };
  constexpr uint32_t kLocalOpsWords = 80;
  uint32_t local_ops[kLocalOpsWords];
  uint32_t* const ops_out = ops;
  if (ops_words <= kLocalOpsWords) ops = local_ops;
  for (uint32_t w = 0; w < ops_words; w++) ops[w] = 0;
  int j = m;
  uint32_t i = wlen;
  int g = cost(j, i);
  int total = g;
  uint32_t pattern_start = 0, overshoot = 0;
  out.failed = 0;
  if (i > slice) {  // src/trace.rs:298-309
    overshoot = i - slice;
    total += overhang_overshoot_cost(alpha, overshoot);
    i -= overshoot;
    j -= (int)overshoot;
  }
  uint32_t nops = 0;
  const uint32_t max_ops = ops_words * 16;
  while (j > 0) {
    if (i == 0) {  // src/trace.rs:320-334
      pattern_start = (uint32_t)j;
      g -= overhang_left_cost(j, alpha, mo);
      break;
    }
    uint32_t op;
    if (cost(j - 1, i - 1) == g && trace_match<P>(pattern[j - 1], text_at_dir(text, n, rev, off + i - 1))) {
      op = kOpEq;
      j--, i--;
    } else {
      g -= 1;
      if (cost(j - 1, i - 1) == g) {
        op = kOpX;
        j--, i--;
      } else if (cost(j, i - 1) == g) {
        op = kOpD;
        i--;
      } else if (cost(j - 1, i) == g) {
        op = kOpI;
        j--;
      } else {
        out.failed = 1;
        break;
      }
    }
    if (nops < max_ops) ops[nops >> 4] |= op << ((nops & 15) * 2);
    nops++;
  }
  if (!out.failed && g != 0) out.failed = 1;  // assert_eq!(g, 0), src/trace.rs:390
  if (nops > max_ops) {
    out.failed = 1;
    nops = max_ops;
  }
  for (uint32_t a = 0, b = nops; a + 1 < b; a++, b--) {
    const uint32_t oa = (ops[a >> 4] >> ((a & 15) * 2)) & 3u;
    const uint32_t ob = (ops[(b - 1) >> 4] >> (((b - 1) & 15) * 2)) & 3u;
    ops[a >> 4] = (ops[a >> 4] & ~(3u << ((a & 15) * 2))) | (ob << ((a & 15) * 2));
    ops[(b - 1) >> 4] = (ops[(b - 1) >> 4] & ~(3u << (((b - 1) & 15) * 2))) | (oa << (((b - 1) & 15) * 2));
  }
  if (ops != ops_out)
    for (uint32_t w = 0; w < ops_words; w++) ops_out[w] = ops[w];
  out.cost = total;
  out.text_start = off + i;
  out.text_end = off + slice;
  out.nops = nops;
  out.failed |= pack_overhang(pattern_start, overshoot);
}

}  // namespace sb
