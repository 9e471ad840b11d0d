// Kernels that run after the scan on the (small) candidate list:
//  * minima_kernel : local-minima selection on the sorted candidates
//  * trace_kernel  : one thread per selected end position -> Match + CIGAR ops
// Their per-thread logic lives in scan_core.cuh (shared with the host emulator).
#include "kernels.cuh"

#include <cub/block/block_radix_sort.cuh>
#include <cub/block/block_scan.cuh>

namespace sb {
namespace {

template <bool FILTER>
__global__ void minima_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ cost, uint64_t n,
                              uint8_t* __restrict__ flags, bool all_minima, const __grid_constant__ EndFilter f) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool keep = select_candidate(keys, cost, i, n, all_minima);
  if (FILTER) keep = keep && end_filter_pass(f, keys[i]);
  flags[i] = keep ? 1 : 0;
}

// (cost, rightmost end) packed so that the minimum is the best match of a slot
__device__ __forceinline__ unsigned long long best_pack(uint64_t key, uint32_t cost) {
  return ((unsigned long long)cost << kPosBits) | (((1ull << kPosBits) - 1) - key_pos(key));
}

__global__ void best_reduce_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ cost, uint64_t n,
                                   const uint8_t* __restrict__ flags, unsigned long long* __restrict__ best) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !flags[i]) return;
  atomicMin(&best[key_qs(keys[i])], best_pack(keys[i], cost[i]));
}

__global__ void best_flag_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ cost, uint64_t n,
                                 uint8_t* __restrict__ flags, const unsigned long long* __restrict__ best) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !flags[i]) return;
  if (best[key_qs(keys[i])] != best_pack(keys[i], cost[i])) flags[i] = 0;
}

// Whole post-processing of a SMALL candidate list in one block: sort by key, keep the first
// copy of every position, apply the local-minima rule, compact.  Reads the candidate count
// on the device, so the host can queue it (and the traceback) right behind the scan without
// a round trip; lists above kSmallCandidates raise `big` and are handled by the general path
// (device radix sort + selection kernel + stream compaction).
constexpr int kSmallThreads = 256;
constexpr int kSmallItems = kSmallCandidates / kSmallThreads;

// ITEMS keys per thread: the list is sorted with ITEMS = 1 when it has at most kSmallThreads
// entries (the usual case: a handful of planted or chance matches) and with kSmallItems else.
template <int ITEMS>
__device__ __forceinline__ void post_small_body(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ cost,
                                                unsigned long long n, uint64_t* __restrict__ sel_keys,
                                                unsigned long long* __restrict__ nsel, bool all_minima, int end_bit,
                                                uint64_t* s_keys, uint32_t* s_cost, void* sort_tmp, void* scan_tmp) {
  using Sort = cub::BlockRadixSort<uint64_t, kSmallThreads, ITEMS, uint32_t>;
  using Scan = cub::BlockScan<uint32_t, kSmallThreads>;
  const uint64_t pad_key = 1ull << end_bit;  // above every real key: padding sorts to the end
  uint64_t k[ITEMS];
  uint32_t v[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    const uint32_t idx = threadIdx.x * ITEMS + i;
    k[i] = idx < n ? keys[idx] : pad_key;
    v[i] = idx < n ? cost[idx] : 0u;
  }
  Sort(*reinterpret_cast<typename Sort::TempStorage*>(sort_tmp)).Sort(k, v, 0, end_bit + 1);
  __syncthreads();
#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    const uint32_t idx = threadIdx.x * ITEMS + i;
    s_keys[idx] = k[i];
    s_cost[idx] = v[i];
  }
  __syncthreads();
  uint32_t flag[ITEMS], pos[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    const uint32_t idx = threadIdx.x * ITEMS + i;
    flag[i] = idx < n && select_candidate(s_keys, s_cost, idx, n, all_minima) ? 1u : 0u;
  }
  uint32_t total;
  Scan(*reinterpret_cast<typename Scan::TempStorage*>(scan_tmp)).ExclusiveSum(flag, pos, total);
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
    if (flag[i]) sel_keys[pos[i]] = k[i];
  if (threadIdx.x == 0) *nsel = total;
}

__global__ void __launch_bounds__(kSmallThreads)
    post_small_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ cost,
                      const unsigned long long* __restrict__ cand_count, uint64_t cand_cap,
                      uint64_t* __restrict__ sel_keys, unsigned long long* __restrict__ nsel,
                      unsigned long long* __restrict__ big, bool all_minima, int end_bit) {
  using SortBig = cub::BlockRadixSort<uint64_t, kSmallThreads, kSmallItems, uint32_t>;
  using SortOne = cub::BlockRadixSort<uint64_t, kSmallThreads, 1, uint32_t>;
  using Scan = cub::BlockScan<uint32_t, kSmallThreads>;
  __shared__ union {
    typename SortBig::TempStorage sort_big;
    typename SortOne::TempStorage sort_one;
  } sort_tmp;
  __shared__ uint64_t s_keys[kSmallCandidates];
  __shared__ uint32_t s_cost[kSmallCandidates];
  __shared__ typename Scan::TempStorage scan_tmp;
  const unsigned long long n = *cand_count;
  if (n > (unsigned long long)kSmallCandidates || n > cand_cap) {
    if (threadIdx.x == 0) {
      *big = 1;
      *nsel = 0;
    }
    return;
  }
  if (threadIdx.x == 0) *big = 0;
  if (n <= (unsigned long long)kSmallThreads)
    post_small_body<1>(keys, cost, n, sel_keys, nsel, all_minima, end_bit, s_keys, s_cost, &sort_tmp, &scan_tmp);
  else
    post_small_body<kSmallItems>(keys, cost, n, sel_keys, nsel, all_minima, end_bit, s_keys, s_cost, &sort_tmp,
                                 &scan_tmp);
}

// Grid-stride over the slice [first, first + count) of the selected candidates; when
// t.count_dev is set the slice end is additionally clipped by the device-side count
// (number of selected candidates written by the stream compaction), so the host need
// not read it back before the launch.
template <int P>
__global__ void trace_kernel(const __grid_constant__ TraceArgs t) {
  extern __shared__ uint32_t trace_smem[];  // column store of the block when t.smem_cols is set
  uint64_t count = t.count;
  if (t.count_dev) {
    const unsigned long long total = *t.count_dev;
    count = total > t.first ? (total - t.first < count ? total - t.first : count) : 0;
  }
  const uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t li = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; li < count; li += nthreads) {
  const uint64_t gi = t.first + li;
  const uint64_t key = t.keys[gi];
  const uint32_t qs = key_qs(key);
  const uint64_t end = key_pos(key);
  const uint8_t* text;
  uint64_t n;
  uint32_t q;
  text_of_slot(t.text, qs, text, n, q);
  const bool rev = t.rev_flags[q] != 0;
  GpuMatch gm;
  gm.qs = qs;
  if (t.costs) {  // without_trace: end position and cost only (reference src/search.rs:1464-1475)
    gm.text_start = ~0ull;
    gm.text_end = end < n ? end : n;
    gm.cost = (int32_t)t.costs[gi];
    gm.nops = 0;
    gm.failed = pack_overhang(0, end > n ? (uint32_t)(end - n) : 0u);  // pattern_end = m - overshoot
  } else {
    ColStore cs;
    if (t.smem_cols) {
      cs.base = trace_smem + threadIdx.x;
      cs.stride = blockDim.x;
    } else {
      cs.base = t.scratch + (li % nthreads);
      cs.stride = nthreads;
    }
    TraceOut out;
    if (t.alpha >= 0.f)
      trace_one_ov<P>(text, n, rev, t.patterns + (size_t)q * t.m, t.m, t.k, t.eq + (size_t)q * t.nrows * t.W, t.W,
                      t.sh0, t.msk0, end, t.alpha, t.max_overhang, cs, t.ops + gi * t.ops_words, t.ops_words, out);
    else
      trace_one<P>(text, n, rev, t.patterns + (size_t)q * t.m, t.m, t.k, t.eq + (size_t)q * t.nrows * t.W, t.W,
                   t.sh0, t.msk0, end, cs, t.ops + gi * t.ops_words, t.ops_words, out);
    gm.text_start = out.text_start;
    gm.text_end = out.text_end;
    gm.cost = out.cost;
    gm.nops = out.nops;
    gm.failed = out.failed;
    if (t.max_n_frac >= 0.f &&
        !n_fraction_ok(text, n, rev, out.text_start, out.text_end < n ? out.text_end : n, t.max_n_frac, 0))
      gm.failed |= 2u;
  }
  t.out[gi] = gm;
  }
}

// One word of the recurrences with explicit carries (as myers_word in scan_kernels.cu).
__device__ __forceinline__ void trace_word(uint32_t& pv, uint32_t& mv, uint32_t eq, uint32_t cin, uint32_t& cout,
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

constexpr int kTraceWideWarps = 4;
constexpr int kTraceWideWindow = 1280;  // characters of a traceback window staged in shared memory
constexpr int kTraceWidePattern = 1024; // pattern characters staged next to it

// Traceback for patterns of many words (W >= 8): ONE WARP per match.  The column store is filled
// systolically -- word w in lane w, lane w one column behind lane w-1, carries by shuffle -- in
// m + k + W steps instead of (m + k) x W word-steps of one thread; lane 0 then walks the path
// (single-bit look-ups in the wide layout).
// WL = words per lane (1 up to 32 words, 2 / 4 for patterns of 64 / 128 words): a lane runs its WL
// words one after the other, the carries leave it after the last one.
template <int P, int WL>
__global__ void __launch_bounds__(32 * kTraceWideWarps) trace_wide_kernel(const __grid_constant__ TraceArgs t) {
  __shared__ uint8_t win[kTraceWideWarps][kTraceWideWindow];
  __shared__ uint8_t pat_s[kTraceWideWarps][kTraceWidePattern];
  // column stores of the block's warps when they fit (launch_trace sets smem_cols): the walk of
  // lane 0 reads two or three words per step, each a round trip to L2 otherwise
  extern __shared__ __align__(16) uint32_t trace_cols[];
  uint64_t count = t.count;
  if (t.count_dev) {
    const unsigned long long total = *t.count_dev;
    count = total > t.first ? (total - t.first < count ? total - t.first : count) : 0;
  }
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t nwarps = (uint64_t)gridDim.x * kTraceWideWarps;
  const int W = t.W, m = t.m, k = t.k;
  const int pad = 32 * W - m;
  const uint32_t nl = (uint32_t)(W / WL);      // active lanes
  const uint32_t w_first = lane * (uint32_t)WL;  // first word of this lane
  constexpr int F = 4;
  for (uint64_t li = (uint64_t)blockIdx.x * kTraceWideWarps + (threadIdx.x >> 5); li < count; li += nwarps) {
    const uint64_t gi = t.first + li;
    const uint64_t key = t.keys[gi];
    const uint32_t qs = key_qs(key);
    const uint64_t end = key_pos(key);
    const uint8_t* text;
    uint64_t n;
    uint32_t q;
    text_of_slot(t.text, qs, text, n, q);
    const bool rev = t.rev_flags[q] != 0;
    const uint32_t* __restrict__ eq = t.eq + (size_t)q * t.nrows * W;
    // contiguous column store per match: the 4 fields of a word are 16 adjacent bytes, the words of a
    // column adjacent lines, so the walk touches one new 128-byte line per step
    ColStore cs;
    cs.base = t.smem_cols ? trace_cols + (threadIdx.x >> 5) * trace_words_per_match(m, k, W)
                          : t.scratch + (li % nwarps) * trace_words_per_match(m, k, W);
    cs.stride = 1;
    const uint64_t fill = (uint64_t)m + (uint64_t)k;
    const uint64_t off = end > fill ? end - fill : 0;
    const uint32_t wlen = (uint32_t)(end - off);
    // stage the window's characters (scan order): the fill reads one per lane and step
    const bool staged = wlen <= (uint32_t)kTraceWideWindow;
    uint8_t* wbuf = win[threadIdx.x >> 5];
    // the pattern too (the walk compares one pattern character per step)
    const uint8_t* pat = t.patterns + (size_t)q * m;
    uint8_t* pbuf = pat_s[threadIdx.x >> 5];
    __syncwarp();
    if (staged)
      for (uint32_t i = lane; i < wlen; i += 32) wbuf[i] = text_at_dir(text, n, rev, off + i);
    if (m <= kTraceWidePattern) {
      for (int i = (int)lane; i < m; i += 32) pbuf[i] = pat[i];
      pat = pbuf;
    }
    __syncwarp();
    uint32_t pv[WL], mv[WL];
#pragma unroll
    for (int j = 0; j < WL; j++) {
      pv[j] = mv[j] = 0;
      if (lane < nl) {
        const int lo = pad - 32 * (int)(w_first + j);
        pv[j] = lo <= 0 ? 0xFFFFFFFFu : (lo >= 32 ? 0u : (0xFFFFFFFFu << lo));
        cs.at((0 * W + w_first + j) * F) = pv[j];  // column 0: D[j][0] = j
        cs.at((0 * W + w_first + j) * F + 1) = 0;
      }
    }
    uint32_t carry = 0;
    // equality words of the next step fetched one step ahead (they do not depend on the carries)
    auto fetch = [&](int32_t c, uint32_t (&e)[WL]) {
#pragma unroll
      for (int j = 0; j < WL; j++) e[j] = 0u;
      if (lane >= nl || c < 0 || c >= (int32_t)wlen) return;
      const uint8_t tc = staged ? wbuf[c] : text_at_dir(text, n, rev, off + (uint64_t)c);
      const uint32_t row = ((uint32_t)tc >> t.sh0) & (t.msk0 & 0xFFu);
#pragma unroll
      for (int j = 0; j < WL; j++) e[j] = __ldg(eq + row * W + w_first + j);
    };
    uint32_t eq_next[WL];
    fetch(-(int32_t)lane, eq_next);
    for (uint32_t step = 0; step < wlen + nl - 1; step++) {
      const int32_t c = (int32_t)step - (int32_t)lane;  // this lane's column is c + 1
      uint32_t eq_cur[WL];
#pragma unroll
      for (int j = 0; j < WL; j++) eq_cur[j] = eq_next[j];
      fetch(c + 1, eq_next);
      uint32_t cout = 0;
      if (lane < nl && c >= 0 && c < (int32_t)wlen) {
        uint32_t cin = carry;
#pragma unroll
        for (int j = 0; j < WL; j++) {
          uint32_t ph, mh;
          trace_word(pv[j], mv[j], eq_cur[j], cin, cout, ph, mh);
          cin = cout;
          const uint32_t slot = (((uint32_t)c + 1) * W + w_first + j) * F;
          *reinterpret_cast<uint4*>(&cs.at(slot)) = make_uint4(pv[j], mv[j], ph, mh);  // 16-byte aligned: slot % 4 == 0
        }
      }
      carry = __shfl_up_sync(0xFFFFFFFFu, cout, 1);
      if (lane == 0) carry = 0;
    }
    __threadfence_block();
    __syncwarp();  // the column store written by all lanes is read by lane 0
    if (lane == 0) {
      TraceOut out;
      trace_walk<P>(text, n, rev, pat, m, W, off, wlen, end, cs, t.ops + gi * t.ops_words, t.ops_words, out,
                    staged ? wbuf : nullptr);
      GpuMatch gm;
      gm.qs = qs;
      gm.text_start = out.text_start;
      gm.text_end = out.text_end;
      gm.cost = out.cost;
      gm.nops = out.nops;
      gm.failed = out.failed;
      if (t.max_n_frac >= 0.f &&
          !n_fraction_ok(text, n, rev, out.text_start, out.text_end < n ? out.text_end : n, t.max_n_frac, 0))
        gm.failed |= 2u;
      t.out[gi] = gm;
    }
    __syncwarp();
  }
}

}  // namespace

// Wide path (one warp per match): patterns of >= 3 words, traced (not cost-only), no overhang.  One
// thread needs ~120 us for a 100-character pattern (430 dependent word-steps, then a walk whose
// every step reads the column store), a warp ~20 us.
static bool trace_is_wide(const TraceArgs& t) { return t.W >= 3 && !t.costs && !(t.alpha >= 0.f); }

cudaError_t launch_minima(const uint64_t* keys, const uint32_t* cost, uint64_t n, uint8_t* flags, bool all_minima,
                          const EndFilter* filter, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const unsigned threads = 256;
  const uint64_t blocks = (n + threads - 1) / threads;
  if (filter) {
    minima_kernel<true><<<(unsigned)blocks, threads, 0, stream>>>(keys, cost, n, flags, all_minima, *filter);
  } else {
    EndFilter none;
    memset(&none, 0, sizeof none);
    minima_kernel<false><<<(unsigned)blocks, threads, 0, stream>>>(keys, cost, n, flags, all_minima, none);
  }
  return cudaGetLastError();
}

cudaError_t launch_best(const uint64_t* keys, const uint32_t* cost, uint64_t n, uint8_t* flags,
                        unsigned long long* best, uint32_t nslots, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(best, 0xFF, (size_t)nslots * sizeof(unsigned long long), stream);
  if (e != cudaSuccess) return e;
  const unsigned threads = 256;
  const uint64_t blocks = (n + threads - 1) / threads;
  best_reduce_kernel<<<(unsigned)blocks, threads, 0, stream>>>(keys, cost, n, flags, best);
  best_flag_kernel<<<(unsigned)blocks, threads, 0, stream>>>(keys, cost, n, flags, best);
  return cudaGetLastError();
}

cudaError_t launch_post_small(const uint64_t* keys, const uint32_t* cost, const unsigned long long* cand_count,
                              uint64_t cand_cap, uint64_t* sel_keys, unsigned long long* nsel,
                              unsigned long long* big, bool all_minima, int end_bit, cudaStream_t stream) {
  post_small_kernel<<<1, kSmallThreads, 0, stream>>>(keys, cost, cand_count, cand_cap, sel_keys, nsel, big,
                                                     all_minima, end_bit);
  return cudaGetLastError();
}

uint64_t trace_threads(uint64_t count) {
  const uint64_t threads = 128;
  uint64_t blocks = (count + threads - 1) / threads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks == 0) blocks = 1;
  return blocks * threads;
}

// matches in flight of the one-warp-per-match kernel (each needs a column store)
static uint64_t trace_wide_warps(uint64_t count);
uint64_t trace_slots(uint64_t count, int W, bool cost_only, bool overhang) {
  return (W >= 3 && !cost_only && !overhang) ? trace_wide_warps(count) : trace_threads(count);
}
static uint64_t trace_wide_warps(uint64_t count) {
  uint64_t blocks = (count + kTraceWideWarps - 1) / kTraceWideWarps;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks == 0) blocks = 1;
  return blocks * kTraceWideWarps;
}

cudaError_t launch_trace(const TraceArgs& t0, cudaStream_t stream) {
  if (t0.count == 0) return cudaSuccess;
  TraceArgs t = t0;
  if (trace_is_wide(t)) {
    // the scratch holds trace_threads(count) column stores, the warps in flight need fewer
    const unsigned blocks = (unsigned)(trace_wide_warps(t.count) / kTraceWideWarps);
    // column stores in shared memory when the block's four fit 36 KB (e.g. m = 100, k = 8: 7 KB each)
    const size_t cols_bytes = (size_t)trace_words_per_match(t.m, t.k, t.W) * sizeof(uint32_t) * kTraceWideWarps;
    t.smem_cols = cols_bytes <= 36 * 1024 ? 1 : 0;
    const size_t dyn = t.smem_cols ? cols_bytes : 0;
#define SB_TW(PP)                                                                                     \
  if (t.W <= 32)                                                                                      \
    trace_wide_kernel<PP, 1><<<blocks, 32 * kTraceWideWarps, dyn, stream>>>(t);                       \
  else if (t.W == 64)                                                                                 \
    trace_wide_kernel<PP, 2><<<blocks, 32 * kTraceWideWarps, dyn, stream>>>(t);                       \
  else if (t.W == 128)                                                                                \
    trace_wide_kernel<PP, 4><<<blocks, 32 * kTraceWideWarps, dyn, stream>>>(t);                       \
  else                                                                                                \
    return cudaErrorInvalidValue;
    switch (t.profile) {
      case kDna: SB_TW(kDna) break;
      case kIupac: SB_TW(kIupac) break;
      case kAscii: SB_TW(kAscii) break;
      default: return cudaErrorInvalidValue;
    }
#undef SB_TW
    return cudaGetLastError();
  }
  const unsigned threads = 128;
  const uint64_t blocks = trace_threads(t.count) / threads;
  // the per-match column store ((m+k+1) columns x W words x 2) lives in shared memory when a
  // block's share fits the default 48 KB: every column is read back by the greedy walk
  const size_t smem = t.costs ? 0 : (size_t)trace_words_per_match(t.m, t.k, t.W) * sizeof(uint32_t) * threads;
  t.smem_cols = (smem > 0 && smem <= 48 * 1024) ? 1 : 0;
  const size_t dyn = t.smem_cols ? smem : 0;
  switch (t.profile) {
    case kDna: trace_kernel<kDna><<<(unsigned)blocks, threads, dyn, stream>>>(t); break;
    case kIupac: trace_kernel<kIupac><<<(unsigned)blocks, threads, dyn, stream>>>(t); break;
    case kAscii: trace_kernel<kAscii><<<(unsigned)blocks, threads, dyn, stream>>>(t); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace sb
