// The scan kernel: one thread per text row, one block per (row tile, query).
//
// Data movement (B200): the text is a 2-D tensor [rows][ltot] of bytes in HBM.
// Every WARP owns a private ring of kScanStages shared-memory buffers and its
// own mbarriers.  Per pipeline stage lane 0 issues one TMA tiled copy
// (cp.async.bulk.tensor.2d, SASS UTMALDG) of a box [32 rows][64 B] with the
// 64-byte swizzle; lane t then reads its own 64-byte row with four
// conflict-free LDS.128 (16-byte chunk c of row t lives at c ^ ((t >> 1) & 3)).
// Warps never wait for each other (no __syncthreads in the loop): a warp that
// takes the rare exact path only delays itself.  No thread issues a global
// load for text in this variant; the LDG variant (per-thread 16-byte loads, no
// staging) exists as an A/B baseline and as a safety net.
//
// Compute: see scan_core.cuh.  Integer/logic only; no tensor cores.
#include "kernels.cuh"

#include <stdlib.h>

#include <mutex>

#ifndef SB_UNROLL
#define SB_UNROLL 2
#endif

namespace sb {
namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(addr), "r"(parity)
      : "memory");
  return ok;
}

// Out of line: the stage was not there at the first look.
__device__ __noinline__ void mbar_wait_slow(uint32_t addr, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(addr, parity))
    if (++spins > (1u << 22)) __trap();  // a lost TMA must fail loudly, not hang the GPU
}

__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity) {
  if (!mbar_try_wait(addr, parity)) mbar_wait_slow(addr, parity);
}

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tmap, int32_t x, int32_t y,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3}], [%4];"
      :
      : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(x), "r"(y), "r"(smem_u32(bar))
      : "memory");
}

constexpr int kWarpsPerBlock = kScanThreads / 32;
constexpr int kWarpStageBytes = 32 * kStageBytes;                  // one TMA box
constexpr int kWarpRingBytes = kScanStages * kWarpStageBytes;      // per warp
constexpr int kRingBytes = kWarpsPerBlock * kWarpRingBytes;        // per block
constexpr int kChunks = kStageBytes / 16;
#ifndef SB_SEQ_STAGES
#define SB_SEQ_STAGES 3
#endif
constexpr int kSeqStages = SB_SEQ_STAGES;  // ring depth of the contiguous-tile kernels (2 KB tiles)
constexpr int kUnroll = SB_UNROLL;  // 16-byte chunks unrolled per loop trip (narrow patterns)

template <int W>
constexpr int min_blocks() {
  // threads per SM: 1024 (<=64 registers) for W <= 2, 768 for W <= 4, 384 for W <= 8
  return (W <= 2 ? 1024 : (W <= 4 ? 768 : (W <= 8 ? 384 : 128))) / kScanThreads;
}

// The row pipeline shared by the scan and the prefilter kernels.  Copies the block's
// query table to shared memory, then feeds every thread its row, 64 bytes per stage:
//   body(stage_idx, own, chunk) with chunk(c) -> the c-th 16-byte piece of the stage.
// static_tbl: destination of the query table when the kernel keeps it in a static
// __shared__ array (its address is then an immediate in every LDS), else nullptr.
template <bool REV, int VARIANT, class Body>
__device__ __forceinline__ void row_pipeline(const CUtensorMap& tmap, const ScanArgs& a, const uint32_t* table_src,
                                             uint32_t table_words, uint32_t* static_tbl, EqTab& tab, Body&& body) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[kWarpsPerBlock * kScanStages];

  const uint32_t tid = threadIdx.x;
  const uint32_t warp = tid >> 5, lane = tid & 31;
  uint32_t tile = blockIdx.x / a.nq;
  if (a.tile_list) {  // regional fallback: only the listed tiles (block-uniform exit)
    if (tile >= *a.tile_count) return;
    tile = a.tile_list[tile];
  }
  const int64_t row0 = (int64_t)tile * kScanThreads + 32 * warp;  // first row of this warp
  const int64_t row = row0 + lane;

  // shared memory carve-up: [text rings, 1024-aligned (TMA variant only)] [query table]
  uint8_t* ring = smem_raw;
  uint32_t* tbl;
  if (VARIANT == kVariantTma) {
    const uint32_t base = smem_u32(smem_raw);
    ring = smem_raw + (((base + 1023u) & ~1023u) - base);
    tbl = reinterpret_cast<uint32_t*>(ring + kRingBytes);
    ring += warp * kWarpRingBytes;
  } else {
    tbl = reinterpret_cast<uint32_t*>(smem_raw);
  }
  if (static_tbl) tbl = static_tbl;
  for (uint32_t i = tid; i < table_words; i += kScanThreads) tbl[i] = table_src[i];
  tab.p = tbl;
  tab.saddr = smem_u32(tbl);

  const uint32_t total = a.g.nwarm + a.g.nstage;
  uint64_t* wbar = &full_bar[warp * kScanStages];
  const uint32_t wbar_addr = smem_u32(wbar);
  const uint32_t ring_addr = smem_u32(ring);

  // Stage `it` of a row: the nwarm warm-up stages are the END (forward scan) or the START
  // (reverse scan) of the neighbouring row, so in memory the stages of one thread are
  // contiguous: forward index of stage it = start + it * step (see stage_coord).  TMA wants
  // 2-D coordinates: x = (it - nwarm) * 64 is the column inside the own row, a negative x
  // lies in the previous row (forward) / the mirrored column beyond the row lies in the next.
  auto issue = [&](uint32_t it) {  // lane 0 only
    const int32_t x = ((int32_t)it - (int32_t)a.g.nwarm) * kStageBytes;
    int32_t col, r = (int32_t)row0;
    if (!REV) {
      col = x;
      if (x < 0) col += (int32_t)a.g.ltot, r -= 1;
    } else {
      col = (int32_t)a.g.ltot - kStageBytes - x;
      if (x < 0) col -= (int32_t)a.g.ltot, r += 1;
    }
    const uint32_t slot = it % kScanStages;
    uint64_t* bar = &wbar[slot];
    mbar_expect_tx(bar, kWarpStageBytes);
    tma_load_2d(ring + slot * kWarpStageBytes, &tmap, col, r, bar);
  };

  if (VARIANT == kVariantTma) {
    if (lane == 0) {
      for (int s = 0; s < kScanStages; s++) mbar_init(&wbar[s], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
  }
  __syncthreads();  // query table + barriers visible; the only block-wide sync
  if (VARIANT == kVariantTma) {
    if (lane == 0) {
      for (uint32_t it = 0; it < (uint32_t)kScanStages && it < total; it++) issue(it);
    }
  }

  // forward index of this thread's stage 0 and the step between stages
  const int64_t warm_bytes = (int64_t)a.g.nwarm * kStageBytes;
  int64_t sidx = REV ? (row + 1) * (int64_t)a.g.ltot + warm_bytes - kStageBytes : row * (int64_t)a.g.ltot - warm_bytes;
  constexpr int64_t kStep = REV ? -(int64_t)kStageBytes : (int64_t)kStageBytes;
  // CU_TENSOR_MAP_SWIZZLE_64B: chunk ^= (row >> 1) & 3;  SWIZZLE_128B: chunk ^= row & 7
  const uint32_t sw = kStageBytes == 64 ? ((lane >> 1) & 3u) : (lane & 7u);
  const uint32_t lane_buf = ring_addr + lane * kStageBytes;

  for (uint32_t it = 0; it < total; ++it, sidx += kStep) {
    const bool own = it >= a.g.nwarm;
    const uint64_t stage_idx = (uint64_t)sidx;
    if (VARIANT == kVariantTma) {
      const uint32_t st = it % kScanStages;
      mbar_wait(wbar_addr + st * 8u, (it / kScanStages) & 1u);
      const uint32_t buf = lane_buf + st * kWarpStageBytes;
      body(stage_idx, own, [&](int c) {
        uint4 v;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "r"(buf + (((uint32_t)c ^ sw) << 4)));
        return v;
      });
      __syncwarp();  // every lane is done with this warp's ring[st]
      if (lane == 0 && it + kScanStages < total) issue(it + kScanStages);
    } else {
      // rows outside the text: (r < 0 || r >= rows) <=> the stage lies outside [0, rows * ltot)
      const bool valid = sidx >= 0 && sidx < (int64_t)a.g.rows * (int64_t)a.g.ltot;
      const uint4* src = reinterpret_cast<const uint4*>(a.text + (valid ? stage_idx : 0));
      body(stage_idx, own, [&](int c) { return valid ? __ldg(src + c) : make_uint4(0, 0, 0, 0); });
    }
  }
}

template <int W, bool REV, int VARIANT>
__global__ void __launch_bounds__(kScanThreads, min_blocks<W>())
    scan_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ ScanArgs a) {
  const uint32_t q = blockIdx.x % a.nq;  // queries vary fastest: blocks sharing a text tile are co-resident (L2 reuse)
  const uint32_t qs = a.qs_base + q;
  EqTab eqt;
  eqt.rowbytes = a.rowbytes;
  Lane<W> s;
  lane_reset<W>(s, a.m);
  int prev_score = a.m;
  row_pipeline<REV, VARIANT>(
      tmap, a, a.eq + (size_t)q * a.nrows * W, a.nrows * W, nullptr, eqt,
      [&](uint64_t stage_idx, bool own, auto chunk) {
        if (!stage_is_special(a, stage_idx)) {
#pragma unroll(W <= 2 ? kUnroll : 1)
          for (int cc = 0; cc < kChunks; cc++) {
            const int c = REV ? (kChunks - 1 - cc) : cc;
            const uint4 v = chunk(c);
            const uint32_t x[4] = {v.x, v.y, v.z, v.w};
            process16<W, REV, false>(s, prev_score, x, stage_idx + 16u * c, a, eqt, qs, own);
          }
        } else {
#pragma unroll 1
          for (int cc = 0; cc < kChunks; cc++) {
            const int c = REV ? (kChunks - 1 - cc) : cc;
            const uint4 v = chunk(c);
            const uint32_t x[4] = {v.x, v.y, v.z, v.w};
            process16<W, REV, true>(s, prev_score, x, stage_idx + 16u * c, a, eqt, qs, own);
          }
        }
      });
}

// Batches of one-word patterns: two patterns per thread (scan_core.cuh: Lane2).  Block = (row tile,
// pattern pair); a.nq counts PAIRS here, a.eq holds the per-class tables of 2 * a.nq - (odd ? 1 : 0)
// queries (a.nq_odd: the last pair's second pattern is a padding copy of the first).
template <bool REV, int VARIANT>
__global__ void __launch_bounds__(kScanThreads, 1024 / kScanThreads)
    scan2_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ ScanArgs a) {
  __shared__ __align__(16) EqPair pair[256];
  const uint32_t pq = blockIdx.x % a.nq;  // pattern pair
  const uint32_t qa = 2 * pq;
  const bool has_b = !(a.nq_odd && pq + 1 == a.nq);
  const uint32_t qb = has_b ? qa + 1 : qa;
  const uint32_t qs = a.qs_base + qa;
  // expand the two per-class tables to one table indexed by the raw text byte
  for (uint32_t c = threadIdx.x; c < 256; c += kScanThreads) {
    const uint32_t row = (c >> a.sh0) & (a.msk0 & 0xFFu);
    EqPair e;
    e.x = a.eq[(size_t)qa * a.nrows + row];
    e.y = a.eq[(size_t)qb * a.nrows + row];
    pair[c] = e;
  }
  const uint32_t saddr = smem_u32(pair);
  EqTab dummy;
  dummy.rowbytes = 4;
  Lane2 s;
  lane2_reset(s, a.m);
  int prev_a = a.m, prev_b = a.m;
  row_pipeline<REV, VARIANT>(
      tmap, a, nullptr, 0, reinterpret_cast<uint32_t*>(pair), dummy,
      [&](uint64_t stage_idx, bool own, auto chunk) {
        if (!stage_is_special(a, stage_idx)) {
#pragma unroll(kUnroll)
          for (int cc = 0; cc < kChunks; cc++) {
            const int c = REV ? (kChunks - 1 - cc) : cc;
            const uint4 v = chunk(c);
            const uint32_t x[4] = {v.x, v.y, v.z, v.w};
            process16_2<REV, false>(s, prev_a, prev_b, x, stage_idx + 16u * c, a, pair, saddr, qs, has_b, own);
          }
        } else {
#pragma unroll 1
          for (int cc = 0; cc < kChunks; cc++) {
            const int c = REV ? (kChunks - 1 - cc) : cc;
            const uint4 v = chunk(c);
            const uint32_t x[4] = {v.x, v.y, v.z, v.w};
            process16_2<REV, true>(s, prev_a, prev_b, x, stage_idx + 16u * c, a, pair, saddr, qs, has_b, own);
          }
        }
      });
}

// Moves a warp's staged hits to the global list with one global atomic.
__device__ __forceinline__ void flush_hits(const ScanArgs& a, const HitQueue& hq, uint32_t lane) {
  __syncwarp();
  const uint32_t staged = *hq.n;
  const uint32_t cnt = staged < kHitQueueCap ? staged : kHitQueueCap;  // the excess went straight to the list
  if (cnt) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(a.hit_count, (unsigned long long)cnt);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    for (uint32_t i = lane; i < cnt; i += 32)
      if (base + i < a.hit_cap) a.hit_keys[base + i] = hq.q[i];
  }
  __syncwarp();
  if (lane == 0) *hq.n = 0;
  __syncwarp();
}

// Stages one hit (16-byte text chunk at forward index base_idx) in the warp's queue; a full
// queue sends it straight to the global list.
__device__ __forceinline__ void push_hit(const ScanArgs& a, const HitQueue& hq, uint32_t qs, uint64_t base_idx) {
  const uint64_t key = cand_key(qs, base_idx / kHitChars);
  const uint32_t pos = atomicAdd(hq.n, 1u);
  if (pos < kHitQueueCap) {
    hq.q[pos] = key;
  } else {
    const unsigned long long i = atomicAdd(a.hit_count, 1ull);
    if (i < a.hit_cap) a.hit_keys[i] = key;
  }
}

// Prefilter: Shift-And automaton over k+1 exact pieces (scan_core.cuh); emits the text
// words in which a piece occurrence ends.
// PAIR: two characters per automaton step through a class-pair table (Dna profile).
template <int WF, bool REV, int VARIANT, bool PAIR>
__global__ void __launch_bounds__(kScanThreads, (WF <= 4 ? 1024 : 512) / kScanThreads)
    filter_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ ScanArgs a) {
  __shared__ uint64_t hit_q[kWarpsPerBlock][kHitQueueCap];
  __shared__ uint32_t hit_n[kWarpsPerBlock];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t q = blockIdx.x % a.nq;
  const uint32_t qs = a.qs_base + q;
  HitQueue hq;
  hq.q = hit_q[warp];
  hq.n = &hit_n[warp];
  if (lane == 0) hit_n[warp] = 0;  // ordered before any use by the __syncthreads in row_pipeline
  EqTab eqt;
  eqt.rowbytes = (uint32_t)WF * 4u;
  FLane<WF> s;
  flane_reset<WF>(s, a);
  constexpr uint32_t kTabWords = PAIR ? kPairTableWords * WF : 256 * WF;
  __shared__ __align__(16) uint32_t pair_tab[PAIR ? kPairTableWords * WF : 4];
  row_pipeline<REV, VARIANT>(tmap, a, a.feq + (size_t)q * kTabWords, kTabWords, PAIR ? pair_tab : nullptr, eqt,
                             [&](uint64_t stage_idx, bool own, auto chunk) {
                               uint32_t acc[kChunks][2];
#pragma unroll
                               for (int cc = 0; cc < kChunks; cc++) {
                                 const int c = REV ? (kChunks - 1 - cc) : cc;
                                 const uint4 v = chunk(c);
                                 const uint32_t x[4] = {v.x, v.y, v.z, v.w};
                                 if (PAIR)
                                   filter16_pair<WF, REV>(s, x, eqt, acc[c]);
                                 else
                                   filter16<WF, REV>(s, x, eqt, acc[c]);
                               }
                               uint32_t any = 0;
#pragma unroll
                               for (int c = 0; c < kChunks; c++) any |= acc[c][0] | acc[c][1];
                               // One vote per stage; everything below runs only in warps that saw a
                               // piece occurrence in this stage (hits are staged per warp, see HitQueue).
                               if (__any_sync(0xFFFFFFFFu, any != 0)) {
                                 if (any && own) {
#pragma unroll
                                   for (int c = 0; c < kChunks; c++) {
                                     const uint64_t base_idx = stage_idx + (uint64_t)(kHitChars * c);
                                     if (base_idx >= a.n) continue;
                                     if (acc[c][0] | (a.fused ? 0u : acc[c][1])) push_hit(a, hq, qs, base_idx);
                                     if (a.fused && acc[c][1]) push_hit(a, hq, qs + a.nq, base_idx);
                                   }
                                 }
                                 __syncwarp();
                                 if (*hq.n >= kHitQueueCap / 2) flush_hits(a, hq, lane);  // warp-uniform
                               }
                             });
  flush_hits(a, hq, lane);
}

// q-gram bitmap prefilter (scan_core.cuh: qgram16): one forward pass for both strands, cost per
// character independent of m and k.  a.feq = the bitmap (4^Q bits), copied to shared memory;
// a.fused = 1: every hit is reported for the reversed partner slot (qs + a.nq) as well.
template <int Q, int S, int VARIANT>
__global__ void __launch_bounds__(kScanThreads, 1024 / kScanThreads)
    qgram_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ ScanArgs a) {
  __shared__ uint64_t hit_q[kWarpsPerBlock][kHitQueueCap];
  __shared__ uint32_t hit_n[kWarpsPerBlock];
  constexpr uint32_t kTabWords = (1u << (2 * Q)) / 32;
  __shared__ __align__(16) uint32_t bitmap[kTabWords];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t q = blockIdx.x % a.nq;
  const uint32_t qs = a.qs_base + q;
  HitQueue hq;
  hq.q = hit_q[warp];
  hq.n = &hit_n[warp];
  if (lane == 0) hit_n[warp] = 0;
  EqTab eqt;
  eqt.rowbytes = 4u;
  QLane s;
  s.w = 0, s.prev = 0;
  row_pipeline<false, VARIANT>(tmap, a, a.feq + (size_t)q * kTabWords, kTabWords, bitmap, eqt,
                               [&](uint64_t stage_idx, bool own, auto chunk) {
                                 uint32_t acc[kChunks];
#pragma unroll
                                 for (int c = 0; c < kChunks; c++) {
                                   const uint4 v = chunk(c);
                                   const uint32_t x[4] = {v.x, v.y, v.z, v.w};
                                   acc[c] = qgram16<Q, S>(s, x, bitmap);
                                 }
                                 uint32_t any = 0;
#pragma unroll
                                 for (int c = 0; c < kChunks; c++) any |= acc[c];
                                 any &= 1u;
                                 if (__any_sync(0xFFFFFFFFu, any != 0)) {
                                   if (any && own) {
#pragma unroll
                                     for (int c = 0; c < kChunks; c++) {
                                       const uint64_t base_idx = stage_idx + (uint64_t)(kHitChars * c);
                                       if (!(acc[c] & 1u) || base_idx >= a.n) continue;
                                       push_hit(a, hq, qs, base_idx);
                                       if (a.fused) push_hit(a, hq, qs + a.nq, base_idx);
                                     }
                                   }
                                   __syncwarp();
                                   if (*hq.n >= kHitQueueCap / 2) flush_hits(a, hq, lane);  // warp-uniform
                                 }
                               });
  flush_hits(a, hq, lane);
}

// The same filter over CONTIGUOUS memory.  The q-gram window only needs the 16 previous
// characters, so a thread does not have to own a long row: a warp streams a contiguous segment of
// the text in 2 KB tiles (one TMA box of 32 "rows" of 64 consecutive bytes each, 64-byte swizzle,
// the same conflict-free shared-memory reads as the row pipeline), lane t takes bytes
// [64 t, 64 t + 64) of the tile and rebuilds its window from the last 16 bytes of lane t - 1 (one
// shuffle of a 16-byte chunk; lane 0 gets them from lane 31 of the previous tile).  DRAM sees long
// sequential bursts instead of 64-byte pieces of 32 rows that lie kilobytes apart, and there is no
// per-row warm-up stage.  tmap: the text as [ceil(n / 64)][64] bytes.
template <int Q, int S>
__global__ void __launch_bounds__(kScanThreads, 1024 / kScanThreads)
    qgram_seq_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ ScanArgs a,
                     const uint64_t total_tiles, const uint32_t tiles_per_warp) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[kWarpsPerBlock * kSeqStages];
  __shared__ uint64_t hit_q[kWarpsPerBlock][kHitQueueCap];
  __shared__ uint32_t hit_n[kWarpsPerBlock];
  constexpr uint32_t kTabWords = (1u << (2 * Q)) / 32;
  __shared__ __align__(16) uint32_t bitmap[kTabWords];
  // tiles are 32 x 64 bytes: builds with another stage size (tiling experiments) run the row-tiled kernel
  if (kStageBytes != 64) return;
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t qs = a.qs_base;
  HitQueue hq;
  hq.q = hit_q[warp];
  hq.n = &hit_n[warp];
  if (lane == 0) hit_n[warp] = 0;
  for (uint32_t i = tid; i < kTabWords; i += kScanThreads) bitmap[i] = a.feq[i];
  const uint32_t base = smem_u32(smem_raw);
  uint8_t* ring = smem_raw + (((base + 1023u) & ~1023u) - base) + warp * (kSeqStages * kWarpStageBytes);
  uint64_t* wbar = &full_bar[warp * kSeqStages];
  const uint32_t wbar_addr = smem_u32(wbar);
  if (lane == 0) {
    for (int st = 0; st < kSeqStages; st++) mbar_init(&wbar[st], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  const uint64_t gwarp = (uint64_t)blockIdx.x * kWarpsPerBlock + warp;
  const uint64_t tile0 = gwarp * tiles_per_warp;
  uint64_t tile1 = tile0 + tiles_per_warp;
  if (tile1 > total_tiles) tile1 = total_tiles;
  const uint32_t ntiles = tile0 < tile1 ? (uint32_t)(tile1 - tile0) : 0u;
  auto issue = [&](uint32_t it) {  // lane 0 only
    const uint32_t slot = it % kSeqStages;
    mbar_expect_tx(&wbar[slot], kWarpStageBytes);
    tma_load_2d(ring + slot * kWarpStageBytes, &tmap, 0, (int32_t)((tile0 + it) * 32), &wbar[slot]);
  };
  if (lane == 0)
    for (uint32_t it = 0; it < (uint32_t)kSeqStages && it < ntiles; it++) issue(it);

  // the 16 bytes before this warp's segment (for lane 0 of the first tile)
  uint4 carry = make_uint4(0, 0, 0, 0);
  if (lane == 0 && tile0 > 0 && ntiles) carry = __ldg(reinterpret_cast<const uint4*>(a.text + tile0 * 2048 - 16));
  const uint32_t sw = (lane >> 1) & 3u;
  const uint32_t lane_buf = smem_u32(ring) + lane * kStageBytes;
  for (uint32_t it = 0; it < ntiles; it++) {
    const uint32_t st = it % kSeqStages;
    mbar_wait(wbar_addr + st * 8u, (it / kSeqStages) & 1u);
    const uint32_t buf = lane_buf + st * kWarpStageBytes;
    uint4 v[kChunks];
#pragma unroll
    for (int c = 0; c < kChunks; c++)
      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(v[c].x), "=r"(v[c].y), "=r"(v[c].z), "=r"(v[c].w)
                   : "r"(buf + (((uint32_t)c ^ sw) << 4)));
    __syncwarp();  // every lane has read this warp's ring[st]
    if (lane == 0 && it + kSeqStages < ntiles) issue(it + kSeqStages);
    // history: the last chunk of the lane to the left (of the previous tile for lane 0)
    uint4 h;
    h.x = __shfl_up_sync(0xFFFFFFFFu, v[kChunks - 1].x, 1);
    h.y = __shfl_up_sync(0xFFFFFFFFu, v[kChunks - 1].y, 1);
    h.z = __shfl_up_sync(0xFFFFFFFFu, v[kChunks - 1].z, 1);
    h.w = __shfl_up_sync(0xFFFFFFFFu, v[kChunks - 1].w, 1);
    if (lane == 0) h = carry;
    carry.x = __shfl_sync(0xFFFFFFFFu, v[kChunks - 1].x, 31);
    carry.y = __shfl_sync(0xFFFFFFFFu, v[kChunks - 1].y, 31);
    carry.z = __shfl_sync(0xFFFFFFFFu, v[kChunks - 1].z, 31);
    carry.w = __shfl_sync(0xFFFFFFFFu, v[kChunks - 1].w, 31);
    QLane s;
    {
      const uint32_t hist[4] = {h.x, h.y, h.z, h.w};
      qlane_init(s, hist);
    }
    uint32_t acc[kChunks];
#pragma unroll
    for (int c = 0; c < kChunks; c++) {
      const uint32_t x[4] = {v[c].x, v[c].y, v[c].z, v[c].w};
      acc[c] = qgram16<Q, S>(s, x, bitmap);
    }
    uint32_t any = 0;
#pragma unroll
    for (int c = 0; c < kChunks; c++) any |= acc[c];
    any &= 1u;
    if (__any_sync(0xFFFFFFFFu, any != 0)) {
      if (any) {
        const uint64_t stage_idx = (tile0 + it) * 2048ull + (uint64_t)lane * kStageBytes;
#pragma unroll
        for (int c = 0; c < kChunks; c++) {
          const uint64_t base_idx = stage_idx + (uint64_t)(kHitChars * c);
          if (!(acc[c] & 1u) || base_idx >= a.n) continue;
          push_hit(a, hq, qs, base_idx);
          if (a.fused) push_hit(a, hq, qs + 1, base_idx);
        }
      }
      __syncwarp();
      if (*hq.n >= kHitQueueCap / 2) flush_hits(a, hq, lane);  // warp-uniform
    }
  }
  flush_hits(a, hq, lane);
}

template <int W, bool SM>
__device__ __forceinline__ void verify_one(const ScanArgs& a, const uint8_t* __restrict__ rev_flags,
                                           unsigned long long i, uint32_t tab_saddr);

// One thread per prefilter hit: exact recurrences over the hit's neighbourhood
// (same window and emission rule as verify_hit in scan_core.cuh, which the host emulator
// runs).  The window is fetched as aligned 16-byte chunks, one chunk ahead of the
// computation, after an L2 prefetch of the whole window: thousands of resident threads
// each touching 2-3 lines would otherwise evict each other's lines from L1 between two
// consecutive byte loads.
// Refined entries of patterns of 3 .. 7 words: up to this many entries the warp-systolic kernel
// takes them, 4 or 8 lanes per entry (launch_verify launches both kernels, each tests the
// device-side count; beyond, one thread per entry hides its latency by itself).
constexpr unsigned long long kWideFewEntries = 32768;
constexpr uint32_t kVerifyEqWords = 2048;  // shared-memory copy of the equality tables in verify_kernel

template <int W>
__global__ void __launch_bounds__(128)
    verify_kernel(const __grid_constant__ ScanArgs a, const uint8_t* __restrict__ rev_flags) {
  if (a.guard_limit && *a.guard_count > a.guard_limit) return;  // too many hits: the regional pass follows
  unsigned long long nhits = *a.hit_count;
  if (nhits > a.hit_cap) nhits = a.hit_cap;
  if (a.wide_few && nhits <= kWideFewEntries) return;  // verify_wide_kernel has them
  const unsigned long long nthreads = (unsigned long long)gridDim.x * blockDim.x;
  // the equality tables of all query slots in shared memory when they fit (one pattern: 4 .. 256 words)
  __shared__ uint32_t eq_s[kVerifyEqWords];
  const uint32_t tab_words = a.nq * a.nrows * (uint32_t)W;
  if (a.nq <= kVerifyEqWords && tab_words <= kVerifyEqWords) {
    if ((unsigned long long)blockIdx.x * blockDim.x >= nhits) return;  // (block-uniform)
    for (uint32_t i = threadIdx.x; i < tab_words; i += blockDim.x) eq_s[i] = a.eq[i];
    __syncthreads();
    const uint32_t saddr = smem_u32(eq_s);
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < nhits; i += nthreads)
      verify_one<W, true>(a, rev_flags, i, saddr);
    return;
  }
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < nhits; i += nthreads)
    verify_one<W, false>(a, rev_flags, i, 0u);
}

// Exact recurrences over scan-direction positions [w0, end) of one text, emitting every end
// position > emit_from with score <= k under query slot `qs`.  `text` must be 16-byte aligned
// and padded to a multiple of 16 bytes.
// SM: the query's equality table lies in shared memory (`tab`; PRMT + IMAD + LDS per look-up instead of
// a 64-bit address computation + LDG), else it is read through `eq`.
template <int W, bool SM = false>
__device__ __forceinline__ void scan_window(const ScanArgs& a, const uint32_t* __restrict__ eq, uint32_t qs, bool rev,
                                            const uint8_t* __restrict__ text, int64_t n, int64_t w0, int64_t end,
                                            int64_t emit_from, EqTab tab = EqTab()) {
  if (end <= w0) return;
  auto step = [&](Lane<W>& st, uint32_t pre, int b) {
    if (SM) {
      uint32_t e[W];
      load_eq<W>(e, tab, pre, b);
      myers_step<W>(st, e);
    } else {
      myers_step<W>(st, eq + ((pre >> (8 * b)) & 0xFFu) * W);
    }
  };
  // forward byte range [lo, hi) of the window and its aligned 16-byte chunks
  const int64_t lo = rev ? n - end : w0, hi = rev ? n - w0 : end;
  const int64_t c_lo = lo >> 4, c_hi = (hi + 15) >> 4;  // chunk indices [c_lo, c_hi)
  const uint4* __restrict__ chunks = reinterpret_cast<const uint4*>(text);
  for (int64_t c = c_lo; c < c_hi && c < c_lo + 16; c += 2)  // one prefetch per 32-byte sector
    asm volatile("prefetch.global.L2 [%0];" ::"l"(chunks + c));

  Lane<W> s;
  lane_reset<W>(s, a.m);
  const int64_t nchunks = c_hi - c_lo;
  // 32-bit scan-direction offsets relative to w0 keep the per-character bookkeeping to one
  // add and one unsigned compare; words and bytes of a chunk are put into scan order once
  // per chunk (SEL / PRMT) so that the unrolled loops index registers statically.
  const uint32_t wlen = (uint32_t)(end - w0);
  const int32_t emit_rel = (int32_t)(emit_from - w0);
  const uint32_t byte_order = rev ? 0x0123u : 0x3210u;
  int64_t c = rev ? c_hi - 1 : c_lo;  // chunks in scan order
  uint4 cur = __ldg(chunks + c);
  int prev = -1;  // score before the current text word once the emit zone has been entered
  for (int64_t t = 0; t < nchunks; t++) {
    const int64_t cn = rev ? c - 1 : c + 1;
    uint4 nxt = make_uint4(0, 0, 0, 0);
    if (t + 1 < nchunks) nxt = __ldg(chunks + cn);
    // scan-direction offset of the chunk's first character in scan order
    int32_t srel0 = rev ? (int32_t)((n - 1 - ((c << 4) + 15)) - w0) : (int32_t)((c << 4) - w0);
    uint32_t q0 = rev ? cur.w : cur.x, q1 = rev ? cur.z : cur.y, q2 = rev ? cur.y : cur.z, q3 = rev ? cur.x : cur.w;
#pragma unroll 1
    for (int j = 0; j < 4; j++, srel0 += 4) {
      const uint32_t wv = __byte_perm(q0, 0u, byte_order);  // the word's characters in scan order
      q0 = q1, q1 = q2, q2 = q3;
      const uint32_t pre = (wv >> a.sh0) & a.msk0;
      // all four characters inside the window?  (one unsigned compare: srel0 in [0, wlen - 4])
      const bool inside = wlen >= 4u && (uint32_t)srel0 <= wlen - 4u;
      if (inside && srel0 + 4 <= emit_rel) {  // warm-up: no position of this word can be reported
#pragma unroll
        for (int b = 0; b < 4; b++) step(s, pre, b);
      } else if (inside && srel0 >= emit_rel) {
        // emit zone: the score moves by at most 1 per character, so a position of this word can
        // only be <= k if score_before + score_after <= 2k + 4 (as fast_group in scan_core.cuh)
        if (prev < 0) prev = lane_score<W>(s);
        const Lane<W> saved = s;
#pragma unroll
        for (int b = 0; b < 4; b++) step(s, pre, b);
        const int score = lane_score<W>(s);
        if (prev + score <= 2 * a.k + 4) {
          s = saved;
#pragma unroll 1
          for (int b = 0; b < 4; b++) {
            step(s, pre, b);
            const int sc = lane_score<W>(s);
            if (sc <= a.k) emit_candidate(a, qs, (uint64_t)(w0 + srel0 + b) + 1, sc);
          }
        }
        prev = score;
      } else {  // a word straddling the window or the start of the emit zone: character by character
#pragma unroll 1
        for (int b = 0; b < 4; b++) {
          const int32_t srel = srel0 + b;
          if ((uint32_t)srel < wlen) {
            step(s, pre, b);
            if (srel >= emit_rel) {
              const int sc = lane_score<W>(s);
              if (sc <= a.k) emit_candidate(a, qs, (uint64_t)(w0 + srel) + 1, sc);
            }
          }
        }
        prev = -1;
      }
    }
    cur = nxt;
    c = cn;
  }
}

template <int W, bool SM>
__device__ __forceinline__ void verify_one(const ScanArgs& a, const uint8_t* __restrict__ rev_flags,
                                           unsigned long long i, uint32_t tab_saddr) {
  const uint64_t key = a.hit_keys[i];
  const uint32_t qs = key_qs(key);
  const bool rev = rev_flags[qs] != 0;
  const uint32_t* __restrict__ eq = a.eq + (size_t)qs * a.nrows * W;

  const int64_t n = (int64_t)a.n;
  int64_t w0, end, emit_from;
  if (a.hit_exact) {  // refined hit: nominal end position
    hit_window_exact(a, key_pos(key), a.hit_span[i], w0, end, emit_from);
  } else {
    const int64_t base = (int64_t)(key_pos(key) * kHitChars);
    if (hit_in_dense_tile(a, (uint64_t)base)) return;  // its tile is scanned whole
    const int64_t g0 = rev ? n - kHitChars - base : base;
    const int64_t span = (int64_t)a.m + (int64_t)a.k;
    w0 = g0 - span;
    if (w0 < 0) w0 = 0;
    end = g0 + kHitChars + span + (rev ? (int64_t)a.rev_lead : 0);
    if (end > n) end = n;
    emit_from = g0 < 0 ? 0 : g0;
  }
  EqTab tab;
  tab.p = nullptr;
  tab.rowbytes = (uint32_t)W * 4u;
  tab.saddr = tab_saddr + qs * a.nrows * (uint32_t)W * 4u;
  scan_window<W, SM>(a, eq, qs, rev, a.text, n, w0, end, emit_from, tab);
}

// search_texts / search_many: many short texts, one thread per (text, query) pair runs the
// exact recurrences over its whole text (the v1 boundary conditions hold per text: fresh state
// at the text start).  Slot = text index * nq + query index.
template <int W>
__global__ void __launch_bounds__(128)
    texts_kernel(const __grid_constant__ ScanArgs a, const __grid_constant__ TextsArgs t) {
  const unsigned long long total = (unsigned long long)t.ntexts * t.nq;
  const unsigned long long nthreads = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long slot = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; slot < total;
       slot += nthreads) {
    const uint32_t ti = (uint32_t)(slot / t.nq);
    const uint32_t q = (uint32_t)(slot - (unsigned long long)ti * t.nq);
    const int64_t n = (int64_t)t.lens[ti];
    if (n == 0) continue;
    const int64_t span = (int64_t)a.m + (int64_t)a.k;
    if (t.overhang) {  // the first m+k end positions (and those beyond the text) come from the edge kernel
      if (n > span)
        scan_window<W>(a, a.eq + (size_t)q * a.nrows * W, (uint32_t)slot, t.rev_flags[q] != 0, t.base + t.offs[ti],
                       n, 0, n, span);
      continue;
    }
    if (t.include_pos0 && a.m <= a.k) emit_candidate(a, (uint32_t)slot, 0, a.m);
    const int64_t stop = (t.prefix && (int64_t)t.prefix < n) ? (int64_t)t.prefix : n;
    scan_window<W>(a, a.eq + (size_t)q * a.nrows * W, (uint32_t)slot, t.rev_flags[q] != 0, t.base + t.offs[ti], n, 0,
                   stop, 0);
  }
}

__global__ void __launch_bounds__(256)
    concat_remap_kernel(const uint64_t* __restrict__ raw_keys, const uint32_t* __restrict__ raw_cost,
                        const unsigned long long* __restrict__ raw_count, uint64_t raw_cap,
                        const __grid_constant__ ScanArgs out, const __grid_constant__ TextsArgs t, uint64_t total,
                        uint64_t skip) {
  unsigned long long n = *raw_count;
  if (n > raw_cap) n = raw_cap;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const uint64_t key = raw_keys[i];
    const uint32_t q = key_qs(key);
    uint32_t ti;
    uint64_t local;
    if (!concat_locate(key_pos(key), t.rev_flags[q] != 0, total, t.offs, t.lens, t.ntexts, ti, local)) continue;
    if (local <= skip) continue;
    emit_candidate(out, ti * t.nq + q, local, (int)raw_cost[i]);
  }
}

// Raises (never lowers) a kernel's dynamic shared-memory limit; one tracker per kernel
// instantiation and device, because the attribute lives in the device's context and
// cudaFuncSetAttribute *sets* the limit rather than maximising it.
// One searcher per host thread is the documented usage (reference bin/grep.rs:488-498), so the
// trackers and the occupancy / configuration caches below are shared between threads: every
// access holds this lock (a few dozen nanoseconds per launch).
std::mutex& config_mutex() {
  static std::mutex m;
  return m;
}

template <class Kern>
cudaError_t ensure_smem(Kern kern, size_t smem, size_t (&set_dev)[64]) {
  std::lock_guard<std::mutex> lock(config_mutex());
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  size_t& cur = set_dev[dev];
  if (smem > cur) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cur = smem;
  }
  return cudaSuccess;
}

template <int WL>
size_t (&wide_smem_tracker())[64] {  // one per kernel: the attribute only ever grows
  static size_t t[64] = {};
  return t;
}
template <int W, bool REV, int VARIANT>
size_t (&scan_smem_tracker())[64] {
  static size_t t[64] = {};
  return t;
}
template <int WF, bool REV, int VARIANT, bool PAIR>
size_t (&filter_smem_tracker())[64] {
  static size_t t[64] = {};
  return t;
}

template <int W, bool REV, int VARIANT>
cudaError_t launch_one(const CUtensorMap* tmap, const ScanArgs& a, size_t smem, cudaStream_t stream) {
  auto kern = scan_kernel<W, REV, VARIANT>;
  {
    cudaError_t e = ensure_smem(kern, smem, scan_smem_tracker<W, REV, VARIANT>());
    if (e != cudaSuccess) return e;
  }
  const uint64_t tiles = ((uint64_t)a.g.rows + kScanThreads - 1) / kScanThreads;
  const uint64_t blocks = tiles * a.nq;
  if (blocks == 0) return cudaSuccess;
  if (blocks > 0x7FFFFFFFull) return cudaErrorInvalidConfiguration;
  CUtensorMap dummy;
  if (!tmap) {
    memset(&dummy, 0, sizeof dummy);
    tmap = &dummy;
  }
  kern<<<(unsigned)blocks, kScanThreads, smem, stream>>>(*tmap, a);
  return cudaGetLastError();
}

template <int W, bool REV, int VARIANT>
int occupancy_one(size_t smem) {
  auto kern = scan_kernel<W, REV, VARIANT>;
  if (ensure_smem(kern, smem, scan_smem_tracker<W, REV, VARIANT>()) != cudaSuccess) return 1;
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kScanThreads, smem) != cudaSuccess) return 1;
  return nb > 0 ? nb : 1;
}

template <int WF, bool REV, int VARIANT, bool PAIR>
cudaError_t launch_filter_one(const CUtensorMap* tmap, const ScanArgs& a, size_t smem, cudaStream_t stream) {
  auto kern = filter_kernel<WF, REV, VARIANT, PAIR>;
  {
    cudaError_t e = ensure_smem(kern, smem, filter_smem_tracker<WF, REV, VARIANT, PAIR>());
    if (e != cudaSuccess) return e;
  }
  const uint64_t tiles = ((uint64_t)a.g.rows + kScanThreads - 1) / kScanThreads;
  const uint64_t blocks = tiles * a.nq;
  if (blocks == 0) return cudaSuccess;
  if (blocks > 0x7FFFFFFFull) return cudaErrorInvalidConfiguration;
  CUtensorMap dummy;
  if (!tmap) {
    memset(&dummy, 0, sizeof dummy);
    tmap = &dummy;
  }
  kern<<<(unsigned)blocks, kScanThreads, smem, stream>>>(*tmap, a);
  return cudaGetLastError();
}

template <int WF, bool REV, int VARIANT, bool PAIR>
int filter_occupancy_one(size_t smem) {
  auto kern = filter_kernel<WF, REV, VARIANT, PAIR>;
  if (ensure_smem(kern, smem, filter_smem_tracker<WF, REV, VARIANT, PAIR>()) != cudaSuccess) return 1;
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kScanThreads, smem) != cudaSuccess) return 1;
  return nb > 0 ? nb : 1;
}

}  // namespace

namespace {

int filter_occupancy(int WF, int variant, bool pair, size_t smem) {
#define SB_OCC(WW)                                                                                              \
  if (pair)                                                                                                     \
    return variant == kVariantTma ? filter_occupancy_one<WW, false, kVariantTma, true>(smem)                    \
                                  : filter_occupancy_one<WW, false, kVariantLdg, true>(smem);                   \
  return variant == kVariantTma ? filter_occupancy_one<WW, false, kVariantTma, false>(smem)                     \
                                : filter_occupancy_one<WW, false, kVariantLdg, false>(smem);
  // 8 words: byte-indexed table only (the pair table is used up to 4 words)
#define SB_OCC8                                                                                                 \
  return variant == kVariantTma ? filter_occupancy_one<8, false, kVariantTma, false>(smem)                      \
                                : filter_occupancy_one<8, false, kVariantLdg, false>(smem);
  switch (WF) {
    case 1: SB_OCC(1)
    case 2: SB_OCC(2)
    case 4: SB_OCC(4)
    case 8: SB_OCC8
    default: return 1;
  }
#undef SB_OCC
#undef SB_OCC8
}

// Resident prefilter blocks per SM.  Measured on B200 (profiles/r01b_filter_residency.md): the
// TMA-fed kernel loses a third of its throughput when a 9th block (36 warps, 72 stages in
// flight) becomes resident, and the one-word pair automaton is fastest with 6.
int filter_target_bps(int WF, bool pair) {
  static const int env = [] {
    const char* e = getenv("SASSY_B200_FILTER_BPS");
    const int x = e ? atoi(e) : 0;
    return x >= 1 && x <= 16 ? x : 0;
  }();
  if (env) return env;
  return (WF == 1 && pair) ? 6 : 8;
}

struct FilterConfig {
  size_t smem = 0;  // dynamic shared memory per block, padded until at most the target is resident
  int bps = 0;
};

FilterConfig filter_config(int WF, int variant, bool pair) {
  static FilterConfig cache[9][2][2];  // (all devices of a box are the same part)
  static std::mutex mu;                // not config_mutex(): filter_occupancy takes that one
  if (WF < 1 || WF > 8) return FilterConfig();
  std::lock_guard<std::mutex> lock(mu);
  FilterConfig& c = cache[WF][variant & 1][pair ? 1 : 0];
  if (c.bps) return c;
  size_t smem = (size_t)256 * WF * sizeof(uint32_t);
  if (variant == kVariantTma) smem += 1024 + (size_t)kRingBytes;
  int occ = filter_occupancy(WF, variant, pair, smem);
  if (variant == kVariantTma) {
    const int target = filter_target_bps(WF, pair);
    while (occ > target && smem + 1024 <= 200 * 1024) {
      smem += 1024;
      occ = filter_occupancy(WF, variant, pair, smem);
    }
  }
  c.smem = smem;
  c.bps = occ;
  return c;
}

}  // namespace

int filter_blocks_per_sm(int WF, int variant, bool pair) { return filter_config(WF, variant, pair).bps; }

cudaError_t launch_filter(int WF, bool rev, int variant, bool pair, const CUtensorMap* tmap, const ScanArgs& a,
                          cudaStream_t stream) {
  const size_t smem = filter_config(WF, variant, pair).smem;
#define SB_FCALL2(WW, PP)                                                                         \
  if (variant == kVariantTma)                                                                     \
    return rev ? launch_filter_one<WW, true, kVariantTma, PP>(tmap, a, smem, stream)              \
               : launch_filter_one<WW, false, kVariantTma, PP>(tmap, a, smem, stream);            \
  else                                                                                            \
    return rev ? launch_filter_one<WW, true, kVariantLdg, PP>(tmap, a, smem, stream)              \
               : launch_filter_one<WW, false, kVariantLdg, PP>(tmap, a, smem, stream);
#define SB_FCALL(WW)       \
  if (pair) {              \
    SB_FCALL2(WW, true)    \
  } else {                 \
    SB_FCALL2(WW, false)   \
  }
  switch (WF) {
    case 1: SB_FCALL(1)
    case 2: SB_FCALL(2)
    case 4: SB_FCALL(4)
    case 8:
      if (pair) return cudaErrorInvalidValue;
      SB_FCALL2(8, false)
    default: return cudaErrorInvalidValue;
  }
#undef SB_FCALL
#undef SB_FCALL2
}

namespace {

template <int Q, int S, int VARIANT>
size_t (&qgram_smem_tracker())[64] {
  static size_t t[64] = {};
  return t;
}

template <int Q, int S, int VARIANT>
cudaError_t launch_qgram_one(const CUtensorMap* tmap, const ScanArgs& a, size_t smem, cudaStream_t stream,
                             int* occupancy) {
  auto kern = qgram_kernel<Q, S, VARIANT>;
  {
    cudaError_t e = ensure_smem(kern, smem, qgram_smem_tracker<Q, S, VARIANT>());
    if (e != cudaSuccess) return e;
  }
  if (occupancy) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kScanThreads, smem) != cudaSuccess) nb = 1;
    *occupancy = nb > 0 ? nb : 1;
    return cudaSuccess;
  }
  const uint64_t tiles = ((uint64_t)a.g.rows + kScanThreads - 1) / kScanThreads;
  const uint64_t blocks = tiles * a.nq;
  if (blocks == 0) return cudaSuccess;
  if (blocks > 0x7FFFFFFFull) return cudaErrorInvalidConfiguration;
  CUtensorMap dummy;
  if (!tmap) {
    memset(&dummy, 0, sizeof dummy);
    tmap = &dummy;
  }
  kern<<<(unsigned)blocks, kScanThreads, smem, stream>>>(*tmap, a);
  return cudaGetLastError();
}

cudaError_t qgram_dispatch(int Q, int S, int variant, const CUtensorMap* tmap, const ScanArgs& a, size_t smem,
                           cudaStream_t stream, int* occupancy) {
#define SB_QCALL(QQ, SS)                                                                            \
  if (Q == QQ && S == SS)                                                                           \
    return variant == kVariantTma ? launch_qgram_one<QQ, SS, kVariantTma>(tmap, a, smem, stream, occupancy) \
                                  : launch_qgram_one<QQ, SS, kVariantLdg>(tmap, a, smem, stream, occupancy);
  SB_QCALL(8, 4) SB_QCALL(8, 8) SB_QCALL(8, 16)
  SB_QCALL(7, 4) SB_QCALL(6, 4)
#undef SB_QCALL
  return cudaErrorInvalidValue;
}

// Resident blocks per SM of the q-gram kernel: the TMA-fed row pipeline is fastest well below
// the register-limited residency (profiles/r01b_filter_residency.md); the dynamic shared memory is
// padded until at most the target is resident.
FilterConfig qgram_config(int Q, int S, int variant) {
  static FilterConfig cache[9][17][2];
  static std::mutex mu;
  if (Q < 6 || Q > 8 || S < 4 || S > 16) return FilterConfig();
  std::lock_guard<std::mutex> lock(mu);
  FilterConfig& c = cache[Q][S][variant & 1];
  if (c.bps) return c;
  static const int target = [] {
    const char* e = getenv("SASSY_B200_QGRAM_BPS");
    const int x = e ? atoi(e) : 0;
    return x >= 1 && x <= 16 ? x : 6;
  }();
  ScanArgs dummy;
  memset(&dummy, 0, sizeof dummy);
  size_t smem = variant == kVariantTma ? 1024 + (size_t)kRingBytes : 0;
  int occ = 1;
  qgram_dispatch(Q, S, variant, nullptr, dummy, smem, nullptr, &occ);
  if (variant == kVariantTma)
    while (occ > target && smem + 1024 <= 200 * 1024) {
      smem += 1024;
      qgram_dispatch(Q, S, variant, nullptr, dummy, smem, nullptr, &occ);
    }
  c.smem = smem;
  c.bps = occ;
  return c;
}

}  // namespace

int qgram_blocks_per_sm(int Q, int S, int variant) { return qgram_config(Q, S, variant).bps; }

namespace {

template <int Q, int S>
cudaError_t launch_qgram_seq_one(const CUtensorMap* tmap, const ScanArgs& a, int sm_count, cudaStream_t stream) {
  auto kern = qgram_seq_kernel<Q, S>;
  static size_t tracker[64] = {};
  static const int target = [] {  // SASSY_B200_QSEQ_BPS: cap on resident blocks per SM (0 = whatever fits)
    const char* e = getenv("SASSY_B200_QSEQ_BPS");
    const int x = e ? atoi(e) : 0;
    return x >= 1 && x <= 16 ? x : 0;
  }();
  static const int waves = [] {
    const char* e = getenv("SASSY_B200_QSEQ_WAVES");
    const int x = e ? atoi(e) : 0;
    return x >= 1 && x <= 64 ? x : 4;
  }();
  // occupancy queries cost tens of microseconds: once per instantiation (all devices are the same part)
  static std::mutex mu;
  static size_t cfg_smem = 0;
  static int cfg_nb = 0;
  size_t smem;
  int nb;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (!cfg_nb) {
      size_t sm = 1024 + (size_t)kWarpsPerBlock * kSeqStages * kWarpStageBytes;
      int occ = 1;
      for (;;) {
        cudaError_t e = ensure_smem(kern, sm, tracker);
        if (e != cudaSuccess) return e;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kScanThreads, sm) != cudaSuccess || occ < 1) occ = 1;
        if (!target || occ <= target || sm + 2048 > 200 * 1024) break;
        sm += 2048;  // pad until at most `target` blocks are resident
      }
      cfg_smem = sm, cfg_nb = occ;
    }
    smem = cfg_smem, nb = cfg_nb;
  }
  {
    cudaError_t e = ensure_smem(kern, smem, tracker);  // another device of the box
    if (e != cudaSuccess) return e;
  }
  const uint64_t total_tiles = (a.n + 2047) / 2048;
  if (total_tiles == 0) return cudaSuccess;
  // `waves` segments per resident warp: long sequential segments, a short tail
  const uint64_t warps = (uint64_t)sm_count * nb * kWarpsPerBlock * waves;
  uint64_t tpw = (total_tiles + warps - 1) / warps;
  if (tpw < 8) tpw = 8;
  const uint64_t blocks = (total_tiles + tpw * kWarpsPerBlock - 1) / (tpw * kWarpsPerBlock);
  if (blocks > 0x7FFFFFFFull) return cudaErrorInvalidConfiguration;
  kern<<<(unsigned)blocks, kScanThreads, smem, stream>>>(*tmap, a, total_tiles, (uint32_t)tpw);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_qgram_seq(int Q, int S, const CUtensorMap* tmap, const ScanArgs& a, int sm_count,
                             cudaStream_t stream) {
#define SB_QCALL(QQ, SS) \
  if (Q == QQ && S == SS) return launch_qgram_seq_one<QQ, SS>(tmap, a, sm_count, stream);
  SB_QCALL(8, 4) SB_QCALL(8, 8) SB_QCALL(8, 16)
  SB_QCALL(7, 4) SB_QCALL(6, 4)
#undef SB_QCALL
  return cudaErrorInvalidValue;
}

cudaError_t launch_qgram(int Q, int S, int variant, const CUtensorMap* tmap, const ScanArgs& a, cudaStream_t stream) {
  return qgram_dispatch(Q, S, variant, tmap, a, qgram_config(Q, S, variant).smem, stream, nullptr);
}

namespace {

// One word of the recurrences with explicit carries (the multi-word form of myers_step): used by
// the warp-systolic kernels, where word w of a pattern lives in lane w.
//   cin / cout: bit 0 = carry of the addition, bit 1 = Ph carry, bit 2 = Mh carry
__device__ __forceinline__ void myers_word(uint32_t& pv, uint32_t& mv, uint32_t eq, uint32_t cin, uint32_t& cout,
                                           uint32_t& ph_out, uint32_t& mh_out) {
  const uint32_t x = eq | mv;
  const uint32_t t = x & pv;
  const uint64_t sum = (uint64_t)t + pv + (cin & 1u);
  const uint32_t u = (uint32_t)sum;
  const uint32_t d0 = (u ^ pv) | x;
  const uint32_t ph = mv | ~(d0 | pv);
  const uint32_t mh = pv & d0;
  const uint32_t ph1 = (ph << 1) | ((cin >> 1) & 1u);
  const uint32_t mh1 = (mh << 1) | ((cin >> 2) & 1u);
  cout = (uint32_t)(sum >> 32) | ((ph >> 31) << 1) | ((mh >> 31) << 2);
  pv = mh1 | ~(d0 | ph1);
  mv = ph1 & d0;
  ph_out = ph;
  mh_out = mh;
}

constexpr int kWideWarps = 4;        // warps per block of the systolic kernels
constexpr int kWideWindow = 2304;    // >= 2 (m + k) + 16 + piece length for m + k <= 1100

// Re-scan of the prefilter's hits for patterns of many words: ONE WARP per hit, word w of the
// pattern in lane w, the lanes skewed by one character (lane w works on character t - w at step
// t), so that the carries of the addition and of the two shifts travel to the next lane with one
// shuffle per step and all lanes are busy: a window of L characters takes L + W steps instead of
// L x W word-steps of a single thread.  Lane W-1 follows the score through the horizontal delta of
// the pattern's last row and emits the candidates.
// WL = words per lane: word w lives in lane w / WL (patterns of up to 32 * WL words: WL = 1 up to 1024
// characters, 2 up to 2048, 4 up to 4096); a lane runs its WL words one after the other, the carries
// leave it after the last one.  wcap = bytes of window staging per warp (dynamic shared memory).
template <int WL>
__global__ void __launch_bounds__(32 * kWideWarps)
    verify_wide_kernel(const __grid_constant__ ScanArgs a, const uint8_t* __restrict__ rev_flags, const int W,
                       const int32_t wcap, const uint32_t G) {
  // G = lanes per entry (a power of two >= W / WL): a warp re-scans 32 / G entries side by side, so
  // the instructions of a step are shared by that many entries (4 words: 8 entries per warp)
  extern __shared__ uint8_t win_dyn[];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t gl = lane & (G - 1);  // lane inside its group
  const uint32_t grp = lane / G, ngrp = 32u / G;
  uint8_t* const win = win_dyn + ((size_t)warp * ngrp + grp) * (size_t)wcap;
  if (a.guard_limit && *a.guard_count > a.guard_limit) return;  // too many hits: the regional pass follows
  unsigned long long nhits = *a.hit_count;
  if (nhits > a.hit_cap) nhits = a.hit_cap;
  // 3 .. 7 words: the lane groups pay off while the entries are few (latency of the word chain: 4x
  // shorter per step); long lists belong to the one-thread-per-entry kernel, which is launched
  // next to this one and applies the opposite test
  if (a.wide_few && nhits > kWideFewEntries) return;
  const unsigned long long nwarps = (unsigned long long)gridDim.x * kWideWarps;
  const int pad = 32 * W - a.m;
  const int nl = W / WL;               // lanes of a group that hold words
  const int w_first = (int)gl * WL;    // first word of this lane
  const bool word_lane = (int)gl < nl;
  const int64_t n = (int64_t)a.n;
  for (unsigned long long base = ((unsigned long long)blockIdx.x * kWideWarps + warp) * ngrp; base < nhits;
       base += nwarps * ngrp) {
    // every lane of a group derives the same window; a group without an entry idles (the warp stays
    // converged: the carries travel by shuffle)
    const unsigned long long h = base + grp;
    bool valid = h < nhits;
    uint32_t qs = 0;
    bool rev = false;
    int64_t w0 = 0, end = 0, emit_from = 0;
    if (valid) {
      const uint64_t key = a.hit_keys[h];
      qs = key_qs(key);
      rev = rev_flags[qs] != 0;
      if (a.hit_exact) {
        hit_window_exact(a, key_pos(key), a.hit_span ? a.hit_span[h] : 0u, w0, end, emit_from);
      } else {
        const int64_t hb = (int64_t)(key_pos(key) * kHitChars);
        if (hit_in_dense_tile(a, (uint64_t)hb)) valid = false;  // its tile is scanned whole
        const int64_t g0 = rev ? n - kHitChars - hb : hb;
        const int64_t span = (int64_t)a.m + (int64_t)a.k;
        w0 = g0 - span;
        if (w0 < 0) w0 = 0;
        end = g0 + kHitChars + span + (rev ? (int64_t)a.rev_lead : 0);
        if (end > n) end = n;
        emit_from = g0 < 0 ? 0 : g0;
      }
      if (end <= w0) valid = false;
    }
    const uint32_t* __restrict__ eq = a.eq + (size_t)qs * a.nrows * W;
    const int32_t L = valid ? (int32_t)(end - w0) : 0;
    const int32_t Lc = L < wcap ? L : wcap;  // (launch_verify sizes wcap for the longest possible window)
    // stage the group's window in scan order
    __syncwarp();
    for (int32_t i = (int32_t)gl; i < Lc; i += (int32_t)G)
      win[i] = text_at_dir(a.text, (uint64_t)n, rev, (uint64_t)(w0 + i));
    __syncwarp();
    const int32_t steps = (int32_t)__reduce_max_sync(0xFFFFFFFFu, Lc > 0 ? Lc + nl - 1 : 0);
    uint32_t pv[WL], mv[WL];
#pragma unroll
    for (int j = 0; j < WL; j++) {
      pv[j] = mv[j] = 0;
      if (word_lane) {
        const int lo = pad - 32 * (w_first + j);
        pv[j] = lo <= 0 ? 0xFFFFFFFFu : (lo >= 32 ? 0u : (0xFFFFFFFFu << lo));
      }
    }
    uint32_t carry = 0;  // carries handed over by the previous lane for the character of this step
    int score = a.m;
    const int32_t emit_rel = (int32_t)(emit_from - w0);
    // the equality words of a step do not depend on the carries: they are fetched one step ahead, so
    // the chain between two steps is the word steps and the shuffle only
    auto fetch = [&](int32_t idx, uint32_t (&e)[WL]) {
#pragma unroll
      for (int j = 0; j < WL; j++) e[j] = 0u;
      if (!word_lane || idx < 0 || idx >= Lc) return;
      const uint32_t row = ((uint32_t)win[idx] >> a.sh0) & (a.msk0 & 0xFFu);
#pragma unroll
      for (int j = 0; j < WL; j++) e[j] = __ldg(eq + row * W + w_first + j);
    };
    uint32_t eq_next[WL];
    fetch(-(int32_t)gl, eq_next);
    for (int32_t t = 0; t < steps; t++) {
      const int32_t idx = t - (int32_t)gl;
      uint32_t eq_cur[WL];
#pragma unroll
      for (int j = 0; j < WL; j++) eq_cur[j] = eq_next[j];
      fetch(idx + 1, eq_next);
      uint32_t cout = 0;
      if (word_lane && idx >= 0 && idx < Lc) {
        uint32_t c = carry, ph = 0, mh = 0;
#pragma unroll
        for (int j = 0; j < WL; j++) {
          myers_word(pv[j], mv[j], eq_cur[j], c, cout, ph, mh);
          c = cout;
        }
        if ((int)gl == nl - 1) {
          score += (int)(ph >> 31) - (int)(mh >> 31);
          if (idx >= emit_rel && score <= a.k) emit_candidate(a, qs, (uint64_t)(w0 + idx) + 1, score);
        }
      }
      carry = __shfl_up_sync(0xFFFFFFFFu, cout, 1);
      if (gl == 0) carry = 0;
    }
  }
}

// Entries that cover a whole text for the re-scan kernels (patterns of more than 32 words have no
// row-tiled scan kernel: the "full scan" is a re-scan of consecutive stretches of `stride` end
// positions, each with its own m + k characters of warm-up).
__global__ void __launch_bounds__(256)
    cover_kernel(uint64_t* __restrict__ keys, uint32_t* __restrict__ spans, unsigned long long* __restrict__ count,
                 uint32_t nq, uint64_t n, uint64_t stride, uint32_t k) {
  const uint64_t per = (n + stride - 1) / stride;
  const uint64_t total = per * nq;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t q = (uint32_t)(i / per);
    const uint64_t j = i - (uint64_t)q * per;
    // an entry reports end positions e0 - k .. e0 + span + k: e0 = j * stride + k + 1 and
    // span = stride - 1 - 2k make that j * stride + 1 .. (j + 1) * stride, a partition (stride > 2k + 1)
    const uint64_t e0 = j * stride + k + 1;
    keys[i] = cand_key(q, e0);
    spans[i] = (uint32_t)(stride - 1 - 2 * (uint64_t)k);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *count = total;
}

}  // namespace

namespace {

__global__ void __launch_bounds__(256)
    refine_kernel(const __grid_constant__ ScanArgs a, const uint8_t* __restrict__ rev_flags, uint64_t* __restrict__ out,
                  uint32_t* __restrict__ out_span, unsigned long long* out_count) {
  if (a.guard_limit && *a.guard_count > a.guard_limit) return;  // too many hits: the regional pass follows
  unsigned long long nhits = *a.hit_count;
  if (nhits > a.hit_cap) nhits = a.hit_cap;
  const unsigned long long nthreads = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < nhits; i += nthreads) {
    const uint64_t key = a.hit_keys[i];
    const uint32_t qs = key_qs(key);
    if (hit_in_dense_tile(a, key_pos(key) * kHitChars)) continue;
    int64_t lo, hi;
    if (!refine_hit(a, qs, rev_flags[qs] != 0, key_pos(key) * kHitChars, lo, hi)) continue;
    // survivors are rare (a whole share of the pattern behind the hit): one atomic each
    const unsigned long long at = atomicAdd(out_count, 1ull);
    if (at < a.hit_cap) {
      out[at] = cand_key(qs, (uint64_t)lo);
      out_span[at] = (uint32_t)(hi - lo);
    }
  }
}

}  // namespace

namespace {

// Hits per tile (a prefilter hit = a 16-byte chunk; all query slots together).
__global__ void __launch_bounds__(256)
    tile_hist_kernel(const __grid_constant__ ScanArgs a, uint32_t* __restrict__ counts, unsigned long long min_hits) {
  unsigned long long nhits = *a.hit_count;
  if (nhits < min_hits) return;  // few hits overall: no tile can be dense enough to matter
  if (nhits > a.hit_cap) nhits = a.hit_cap;
  const unsigned long long nthreads = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < nhits; i += nthreads)
    atomicAdd(&counts[(key_pos(a.hit_keys[i]) * kHitChars) / a.tile_bytes], 1u);
}

// dense[t] = 1 and t appended to the list when the tile's hits exceed `max_hits_per_tile`.
__global__ void __launch_bounds__(256)
    tile_mark_kernel(const uint32_t* __restrict__ counts, uint32_t ntiles, uint32_t max_hits_per_tile,
                     uint8_t* __restrict__ dense, uint32_t* __restrict__ list, uint32_t* __restrict__ list_count) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntiles) return;
  const bool d = counts[t] > max_hits_per_tile;
  dense[t] = d ? 1 : 0;
  if (d) list[atomicAdd(list_count, 1u)] = t;
}

}  // namespace

cudaError_t launch_tile_marks(const ScanArgs& a, uint32_t ntiles, uint32_t* counts, unsigned long long min_hits,
                              uint32_t max_hits_per_tile, uint8_t* dense, uint32_t* list, uint32_t* list_count,
                              cudaStream_t stream) {
  tile_hist_kernel<<<148 * 4, 256, 0, stream>>>(a, counts, min_hits);
  tile_mark_kernel<<<(ntiles + 255) / 256, 256, 0, stream>>>(counts, ntiles, max_hits_per_tile, dense, list, list_count);
  return cudaGetLastError();
}

cudaError_t launch_refine(const ScanArgs& a, const uint8_t* rev_flags, uint64_t* out, uint32_t* out_span,
                          unsigned long long* out_count, cudaStream_t stream) {
  refine_kernel<<<148 * 8, 256, 0, stream>>>(a, rev_flags, out, out_span, out_count);
  return cudaGetLastError();
}

cudaError_t launch_cover(uint64_t* keys, uint32_t* spans, unsigned long long* count, uint32_t nq, uint64_t n,
                         uint64_t stride, uint32_t k, cudaStream_t stream) {
  cover_kernel<<<148 * 2, 256, 0, stream>>>(keys, spans, count, nq, n, stride, k);
  return cudaGetLastError();
}

cudaError_t launch_verify(int W, const ScanArgs& a, const uint8_t* rev_flags, cudaStream_t stream) {
  const unsigned threads = 128;
  const unsigned blocks = 148 * 12;  // grid-stride over the device-side hit count
  // many words: one warp per hit (systolic), when the window fits the staging buffer
  // (refined entries: warm-up m + k, 2k + 1 end positions, and a span of at most 16 + m on repetitive text)
  const int64_t window = a.hit_exact ? 2 * (int64_t)a.m + 3 * (int64_t)a.k + 18
                                     : 2 * ((int64_t)a.m + a.k) + kHitChars + (int64_t)a.rev_lead;
  // lanes per entry of the warp-systolic kernel: the smallest power of two that holds the words
  uint32_t G = 32;
  if (W <= 16) G = 16;
  if (W <= 8) G = 8;
  if (W <= 4) G = 4;
  // window staging per entry: the longest window (<= 2304 bytes up to 32 words)
  auto window_cap = [&]() -> int32_t {
    if (W > 32) {
      // warm-up m + k, 2k + 1 end positions, the span (refined hits: at most 16 + m; cover: the stride)
      const int64_t longest = a.hit_exact ? (int64_t)a.m + 3 * (int64_t)a.k + 2 + std::max<int64_t>(a.max_span, a.m + 16)
                                          : window;
      return (int32_t)((longest + 64 + 127) / 128 * 128);
    }
    return (int32_t)((window + 127) / 128 * 128);
  };
  // few refined entries of 3 .. 7 words: both kernels, the device-side count decides which one works
  const bool few_route = a.hit_exact && W >= 3 && W < 8 && window <= kWideWindow;
  if (few_route) {
    const int32_t wcap = window_cap();
    const size_t smem = (size_t)kWideWarps * (32u / G) * (size_t)wcap;
    cudaError_t e = ensure_smem(verify_wide_kernel<1>, smem, wide_smem_tracker<1>());
    if (e != cudaSuccess) return e;
    ScanArgs b = a;
    b.wide_few = 1;
    verify_wide_kernel<1><<<148 * 8, 32 * kWideWarps, smem, stream>>>(b, rev_flags, W, wcap, G);
    switch (W) {
#define SB_VCALL(WW) case WW: verify_kernel<WW><<<blocks, threads, 0, stream>>>(b, rev_flags); break;
      SB_VCALL(3) SB_VCALL(4) SB_VCALL(6)
#undef SB_VCALL
      default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
  }
  if (W >= 8 && (window <= kWideWindow || W > 32)) {
    const unsigned wblocks = 148 * 8;
    const int32_t wcap = window_cap();
    const size_t smem = (size_t)kWideWarps * (32u / G) * (size_t)wcap;
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaSuccess;
    if (W <= 32) {
      e = ensure_smem(verify_wide_kernel<1>, smem, wide_smem_tracker<1>());
      if (e == cudaSuccess) verify_wide_kernel<1><<<wblocks, 32 * kWideWarps, smem, stream>>>(a, rev_flags, W, wcap, G);
    } else if (W == 64) {
      e = ensure_smem(verify_wide_kernel<2>, smem, wide_smem_tracker<2>());
      if (e == cudaSuccess) verify_wide_kernel<2><<<wblocks, 32 * kWideWarps, smem, stream>>>(a, rev_flags, W, wcap, 32u);
    } else if (W == 128) {
      e = ensure_smem(verify_wide_kernel<4>, smem, wide_smem_tracker<4>());
      if (e == cudaSuccess) verify_wide_kernel<4><<<wblocks, 32 * kWideWarps, smem, stream>>>(a, rev_flags, W, wcap, 32u);
    } else {
      return cudaErrorInvalidValue;
    }
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
  }
  switch (W) {
#define SB_VCALL(WW) case WW: verify_kernel<WW><<<blocks, threads, 0, stream>>>(a, rev_flags); break;
    SB_VCALL(1) SB_VCALL(2) SB_VCALL(3) SB_VCALL(4) SB_VCALL(6) SB_VCALL(8) SB_VCALL(16) SB_VCALL(32)
#undef SB_VCALL
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

namespace {

template <int W>
__global__ void __launch_bounds__(128)
    overhang_edges_kernel(const __grid_constant__ ScanArgs a, const __grid_constant__ OverhangArgs o) {
  const unsigned long long total = 2ull * o.nslots;
  const unsigned long long nthreads = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long id = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += nthreads) {
    const uint32_t slot = (uint32_t)(id >> 1);
    const bool right = (id & 1) != 0;
    const uint8_t* text;
    uint64_t n;
    uint32_t q;
    text_of_slot(o.text, slot, text, n, q);
    if (n + o.steps == 0) continue;  // max_pos == 0: nothing is reported (src/search.rs:1314-1316)
    const bool rev = o.rev_flags[q] != 0;
    const uint32_t* __restrict__ eq = a.eq + (size_t)q * a.nrows * W;
    const uint64_t span = (uint64_t)a.m + (uint64_t)a.k;
    Lane<W> init;
#pragma unroll
    for (int w = 0; w < W; w++) init.pv[w] = o.init_pv[w], init.mv[w] = 0;
    uint32_t wild[W];
#pragma unroll
    for (int w = 0; w < W; w++) wild[w] = 0xFFFFFFFFu;
    if (!right) {
      // end positions 0 .. min(n, m+k) from the overhang left column
      if (o.left_total <= a.k) emit_candidate(a, slot, 0, o.left_total);
      Lane<W> s = init;
      const uint64_t lim = n < span ? n : span;
      for (uint64_t i = 0; i < lim; i++) {
        const uint8_t c = text_at_dir(text, n, rev, i);
        myers_step<W>(s, eq + (((uint32_t)c >> a.sh0) & (a.msk0 & 0xFFu)) * W);
        const int sc = lane_score<W>(s);
        if (sc <= a.k) emit_candidate(a, slot, i + 1, sc);
      }
    } else if (o.steps > 0) {
      // end positions n+1 .. n+steps: wildcard columns, + floor(alpha * overshoot)
      const uint64_t w0 = n > span ? n - span : 0;
      Lane<W> s;
      if (w0 == 0)
        s = init;
      else
        lane_reset<W>(s, a.m);
      for (uint64_t i = w0; i < n; i++) {
        const uint8_t c = text_at_dir(text, n, rev, i);
        myers_step<W>(s, eq + (((uint32_t)c >> a.sh0) & (a.msk0 & 0xFFu)) * W);
      }
      for (uint32_t ov = 1; ov <= o.steps; ov++) {
        myers_step<W>(s, wild);
        const int sc = lane_score<W>(s) + overhang_overshoot_cost(o.alpha, ov);
        if (sc <= a.k) emit_candidate(a, slot, n + ov, sc);
      }
    }
  }
}

}  // namespace

cudaError_t launch_overhang_edges(int W, const ScanArgs& a, const OverhangArgs& o, cudaStream_t stream) {
  const unsigned long long total = 2ull * o.nslots;
  if (total == 0) return cudaSuccess;
  const unsigned threads = 128;
  const unsigned blocks = (unsigned)std::min<unsigned long long>((total + threads - 1) / threads, 148ull * 16);
  switch (W) {
#define SB_OCALL(WW) case WW: overhang_edges_kernel<WW><<<blocks, threads, 0, stream>>>(a, o); break;
    SB_OCALL(1) SB_OCALL(2) SB_OCALL(3) SB_OCALL(4) SB_OCALL(6) SB_OCALL(8) SB_OCALL(16) SB_OCALL(32)
#undef SB_OCALL
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t launch_concat_remap(const uint64_t* raw_keys, const uint32_t* raw_cost, const unsigned long long* raw_count,
                                uint64_t raw_cap, const ScanArgs& out, const TextsArgs& t, uint64_t total,
                                uint64_t skip, cudaStream_t stream) {
  concat_remap_kernel<<<148 * 4, 256, 0, stream>>>(raw_keys, raw_cost, raw_count, raw_cap, out, t, total, skip);
  return cudaGetLastError();
}

cudaError_t launch_texts(int W, const ScanArgs& a, const TextsArgs& t, cudaStream_t stream) {
  const unsigned long long total = (unsigned long long)t.ntexts * t.nq;
  if (total == 0) return cudaSuccess;
  const unsigned threads = 128;
  const unsigned blocks = (unsigned)std::min<unsigned long long>((total + threads - 1) / threads, 148ull * 16);
  switch (W) {
#define SB_TCALL(WW) case WW: texts_kernel<WW><<<blocks, threads, 0, stream>>>(a, t); break;
    SB_TCALL(1) SB_TCALL(2) SB_TCALL(3) SB_TCALL(4) SB_TCALL(6) SB_TCALL(8) SB_TCALL(16) SB_TCALL(32)
#undef SB_TCALL
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

namespace {

template <bool REV, int VARIANT>
cudaError_t launch_scan2_one(const CUtensorMap* tmap, const ScanArgs& a, cudaStream_t stream) {
  auto kern = scan2_kernel<REV, VARIANT>;
  static size_t tracker[64] = {};
  const size_t smem = VARIANT == kVariantTma ? 1024 + (size_t)kRingBytes : 0;
  {
    cudaError_t e = ensure_smem(kern, smem, tracker);
    if (e != cudaSuccess) return e;
  }
  const uint64_t tiles = ((uint64_t)a.g.rows + kScanThreads - 1) / kScanThreads;
  const uint64_t blocks = tiles * a.nq;
  if (blocks == 0) return cudaSuccess;
  if (blocks > 0x7FFFFFFFull) return cudaErrorInvalidConfiguration;
  CUtensorMap dummy;
  if (!tmap) {
    memset(&dummy, 0, sizeof dummy);
    tmap = &dummy;
  }
  kern<<<(unsigned)blocks, kScanThreads, smem, stream>>>(*tmap, a);
  return cudaGetLastError();
}

}  // namespace

// a.nq = number of PATTERNS here; pairs are formed inside.
cudaError_t launch_scan2(bool rev, int variant, const CUtensorMap* tmap, const ScanArgs& a0, cudaStream_t stream) {
  ScanArgs a = a0;
  a.nq_odd = a0.nq & 1u;
  a.nq = (a0.nq + 1) / 2;
  if (variant == kVariantTma)
    return rev ? launch_scan2_one<true, kVariantTma>(tmap, a, stream) : launch_scan2_one<false, kVariantTma>(tmap, a, stream);
  return rev ? launch_scan2_one<true, kVariantLdg>(tmap, a, stream) : launch_scan2_one<false, kVariantLdg>(tmap, a, stream);
}

size_t scan_smem_bytes(int W, int variant, uint32_t nrows) {
  size_t eq = (size_t)nrows * W * sizeof(uint32_t);
  if (variant == kVariantTma) return 1024 + (size_t)kRingBytes + eq;
  return eq;
}

#define SB_DISPATCH_W(W_, CALL)            \
  switch (W_) {                            \
    case 1: CALL(1); break;                \
    case 2: CALL(2); break;                \
    case 3: CALL(3); break;                \
    case 4: CALL(4); break;                \
    case 6: CALL(6); break;                \
    case 8: CALL(8); break;                \
    case 16: CALL(16); break;              \
    case 32: CALL(32); break;              \
    default: break;                        \
  }

cudaError_t launch_scan(int W, bool rev, int variant, const CUtensorMap* tmap, const ScanArgs& a,
                        cudaStream_t stream) {
  const size_t smem = scan_smem_bytes(W, variant, a.nrows);
  cudaError_t e = cudaErrorInvalidValue;
#define SB_CALL(WW)                                                                             \
  if (variant == kVariantTma)                                                                   \
    e = rev ? launch_one<WW, true, kVariantTma>(tmap, a, smem, stream)                          \
            : launch_one<WW, false, kVariantTma>(tmap, a, smem, stream);                        \
  else                                                                                          \
    e = rev ? launch_one<WW, true, kVariantLdg>(tmap, a, smem, stream)                          \
            : launch_one<WW, false, kVariantLdg>(tmap, a, smem, stream);
  SB_DISPATCH_W(W, SB_CALL)
#undef SB_CALL
  return e;
}

int scan_blocks_per_sm(int W, bool rev, int variant, uint32_t nrows) {
  static int cache[33][2][2][3] = {};  // occupancy queries cost tens of microseconds: ask once per shape
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  const int rb = nrows <= 4 ? 0 : (nrows <= 32 ? 1 : 2);
  int* slot = (W >= 1 && W <= 32) ? &cache[W][rev ? 1 : 0][variant & 1][rb] : nullptr;
  if (slot && *slot) return *slot;
  const size_t smem = scan_smem_bytes(W, variant, nrows);
  int nb = 1;
#define SB_CALL(WW)                                                                    \
  if (variant == kVariantTma)                                                          \
    nb = rev ? occupancy_one<WW, true, kVariantTma>(smem) : occupancy_one<WW, false, kVariantTma>(smem); \
  else                                                                                 \
    nb = rev ? occupancy_one<WW, true, kVariantLdg>(smem) : occupancy_one<WW, false, kVariantLdg>(smem);
  SB_DISPATCH_W(W, SB_CALL)
#undef SB_CALL
  if (slot) *slot = nb;
  return nb;
}

}  // namespace sb
