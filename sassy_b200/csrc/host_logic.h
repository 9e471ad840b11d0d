// Host-side logic shared by the CUDA engine (engine.cu) and the CPU-only host
// emulator (emu.cpp): equality-table construction and the row tiling.
#pragma once
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "scan_core.cuh"

namespace sb {

#ifndef SB_THREADS
#define SB_THREADS 128
#endif
constexpr int kScanThreads = SB_THREADS;  // rows (threads) per block; each warp has its own 32-row TMA box
constexpr uint32_t kMaxRowBytes = 16384;
constexpr uint32_t kRowAlign = 128;  // rows start on 128-byte lines
constexpr int kMaxWords = 32;  // patterns up to 32*32 = 1024 characters

// Supported word counts (kernel template instantiations).
inline int round_words(int w) {
  const int opts[] = {1, 2, 3, 4, 6, 8, 16, 32};
  for (int o : opts)
    if (w <= o) return o;
  return -1;
}

struct ProfileParams {
  uint32_t nrows, sh0, msk0;
};

inline bool profile_params(int profile, ProfileParams& p) {
  switch (profile) {
    case kDna: p = {4, 1, 0x03030303u}; return true;
    case kIupac: p = {32, 0, 0x1F1F1F1Fu}; return true;
    case kAscii: p = {256, 0, 0xFFFFFFFFu}; return true;
    default: return false;
  }
}

// Equality table of one query: eq[row][w], bit (pad + j) set iff pattern[j]
// matches a text byte that selects `row`; the low `pad` wildcard bits are all
// ones (cf. reference src/pattern_tiling/tqueries.rs:90-114 for the
// per-pattern peq build; the wildcard padding is ours).
inline void build_eq_table(int profile, const uint8_t* p, int m, int W, uint32_t nrows, uint32_t* tab) {
  const int pad = 32 * W - m;
  memset(tab, 0, (size_t)nrows * W * sizeof(uint32_t));
  for (uint32_t row = 0; row < nrows; row++) {
    uint32_t* e = tab + (size_t)row * W;
    for (int b = 0; b < pad; b++) e[b >> 5] |= 1u << (b & 31);
    for (int j = 0; j < m; j++) {
      bool match;
      switch (profile) {
        case kDna: match = row_matches<kDna>(p[j], (int)row); break;
        case kIupac: match = row_matches<kIupac>(p[j], (int)row); break;
        default: match = row_matches<kAscii>(p[j], (int)row); break;
      }
      if (match) {
        const int b = pad + j;
        e[b >> 5] |= 1u << (b & 31);
      }
    }
  }
}

// Row tiling: rows of `ltot` bytes, one thread per row, kScanThreads rows per
// block, one block per (tile, query).  ltot is as long as possible (the
// per-row warm-up costs nwarm*128 bytes) subject to filling the GPU with a
// whole number of waves when the grid is small.  bpw = resident blocks per wave.
inline ScanGeom choose_geom(uint64_t n, int m, int k, uint32_t nq, int bpw) {
  ScanGeom g;
  g.nwarm = (uint32_t)((m + k + kStageBytes - 1) / kStageBytes);
  if (g.nwarm == 0) g.nwarm = 1;
  const uint64_t min_ltot = ((uint64_t)g.nwarm * kStageBytes + kRowAlign - 1) / kRowAlign * kRowAlign;
  const uint64_t per_tile_max = (uint64_t)kScanThreads * kMaxRowBytes;
  uint64_t tiles = std::max<uint64_t>(1, (n + per_tile_max - 1) / per_tile_max);
  const uint64_t fill = ((uint64_t)bpw + nq - 1) / nq;  // tiles needed to give every SM slot a block
  tiles = std::max(tiles, fill);
  uint64_t blocks = tiles * nq;
  if (blocks < 16ull * bpw) {  // few waves: round up to whole waves
    blocks = (blocks + bpw - 1) / bpw * bpw;
    tiles = (blocks + nq - 1) / nq;
  }
  uint64_t ltot = (n + tiles * kScanThreads - 1) / (tiles * kScanThreads);
  ltot = (ltot + kRowAlign - 1) / kRowAlign * kRowAlign;
  ltot = std::max(ltot, min_ltot);
  ltot = std::min<uint64_t>(ltot, std::max<uint64_t>(kMaxRowBytes, min_ltot));
  g.ltot = (uint32_t)ltot;
  g.rows = (uint32_t)((n + ltot - 1) / ltot);
  g.nstage = g.ltot / kStageBytes;
  return g;
}

inline size_t padded_alloc(uint64_t n) {
  // room for one extra row of any tiling plus alignment slack
  return (size_t)((n + 2ull * kMaxRowBytes + 255ull) & ~255ull);
}

}  // namespace sb
