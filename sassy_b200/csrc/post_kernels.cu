// Kernels that run after the scan on the (small) candidate list:
//  * minima_kernel : local-minima selection on the sorted candidates
//  * trace_kernel  : one thread per selected end position -> Match + CIGAR ops
// Their per-thread logic lives in scan_core.cuh (shared with the host emulator).
#include "kernels.cuh"

#include <cub/block/block_radix_sort.cuh>
#include <cub/block/block_scan.cuh>

namespace sb {
namespace {

__global__ void minima_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ cost, uint64_t n,
                              uint8_t* __restrict__ flags, bool all_minima) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  flags[i] = select_candidate(keys, cost, i, n, all_minima) ? 1 : 0;
}

// Whole post-processing of a SMALL candidate list in one block: sort by key, keep the first
// copy of every position, apply the local-minima rule, compact.  Reads the candidate count
// on the device, so the host can queue it (and the traceback) right behind the scan without
// a round trip; lists above kSmallCandidates raise `big` and are handled by the general path
// (device radix sort + selection kernel + stream compaction).
constexpr int kSmallThreads = 256;
constexpr int kSmallItems = kSmallCandidates / kSmallThreads;

__global__ void __launch_bounds__(kSmallThreads)
    post_small_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ cost,
                      const unsigned long long* __restrict__ cand_count, uint64_t cand_cap,
                      uint64_t* __restrict__ sel_keys, unsigned long long* __restrict__ nsel,
                      unsigned long long* __restrict__ big, bool all_minima, int end_bit) {
  using Sort = cub::BlockRadixSort<uint64_t, kSmallThreads, kSmallItems, uint32_t>;
  using Scan = cub::BlockScan<uint32_t, kSmallThreads>;
  __shared__ union {
    typename Sort::TempStorage sort;
    struct {
      uint64_t keys[kSmallCandidates];
      uint32_t cost[kSmallCandidates];
    } sorted;
  } sm;
  __shared__ typename Scan::TempStorage scan_tmp;
  const unsigned long long n = *cand_count;
  if (n > (unsigned long long)kSmallCandidates || n > cand_cap) {
    if (threadIdx.x == 0) {
      *big = 1;
      *nsel = 0;
    }
    return;
  }
  uint64_t k[kSmallItems];
  uint32_t v[kSmallItems];
#pragma unroll
  for (int i = 0; i < kSmallItems; i++) {
    const uint32_t idx = threadIdx.x * kSmallItems + i;
    k[i] = idx < n ? keys[idx] : ~0ull;  // padding sorts to the end
    v[i] = idx < n ? cost[idx] : 0u;
  }
  Sort(sm.sort).Sort(k, v, 0, 64);
  (void)end_bit;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kSmallItems; i++) {
    const uint32_t idx = threadIdx.x * kSmallItems + i;
    sm.sorted.keys[idx] = k[i];
    sm.sorted.cost[idx] = v[i];
  }
  __syncthreads();
  uint32_t flag[kSmallItems], pos[kSmallItems];
#pragma unroll
  for (int i = 0; i < kSmallItems; i++) {
    const uint32_t idx = threadIdx.x * kSmallItems + i;
    flag[i] = idx < n && select_candidate(sm.sorted.keys, sm.sorted.cost, idx, n, all_minima) ? 1u : 0u;
  }
  uint32_t total;
  Scan(scan_tmp).ExclusiveSum(flag, pos, total);
#pragma unroll
  for (int i = 0; i < kSmallItems; i++)
    if (flag[i]) sel_keys[pos[i]] = k[i];
  if (threadIdx.x == 0) {
    *nsel = total;
    *big = 0;
  }
}

// Grid-stride over the slice [first, first + count) of the selected candidates; when
// t.count_dev is set the slice end is additionally clipped by the device-side count
// (number of selected candidates written by the stream compaction), so the host need
// not read it back before the launch.
template <int P>
__global__ void trace_kernel(const __grid_constant__ TraceArgs t) {
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
  const bool rev = t.rev_flags[qs] != 0;
  ColStore cs;
  cs.base = t.scratch + (li % nthreads);
  cs.stride = nthreads;
  TraceOut out;
  trace_one<P>(t.text, t.n, rev, t.patterns + (size_t)qs * t.m, t.m, t.k,
               t.eq + (size_t)qs * t.nrows * t.W, t.W, t.sh0, t.msk0, end, cs,
               t.ops + gi * t.ops_words, t.ops_words, out);
  GpuMatch gm;
  gm.text_start = out.text_start;
  gm.text_end = out.text_end;
  gm.qs = qs;
  gm.cost = out.cost;
  gm.nops = out.nops;
  gm.failed = out.failed;
  t.out[gi] = gm;
  }
}

}  // namespace

cudaError_t launch_minima(const uint64_t* keys, const uint32_t* cost, uint64_t n, uint8_t* flags, bool all_minima,
                          cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const unsigned threads = 256;
  const uint64_t blocks = (n + threads - 1) / threads;
  minima_kernel<<<(unsigned)blocks, threads, 0, stream>>>(keys, cost, n, flags, all_minima);
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

cudaError_t launch_trace(const TraceArgs& t, cudaStream_t stream) {
  if (t.count == 0) return cudaSuccess;
  const unsigned threads = 128;
  const uint64_t blocks = trace_threads(t.count) / threads;
  switch (t.profile) {
    case kDna: trace_kernel<kDna><<<(unsigned)blocks, threads, 0, stream>>>(t); break;
    case kIupac: trace_kernel<kIupac><<<(unsigned)blocks, threads, 0, stream>>>(t); break;
    case kAscii: trace_kernel<kAscii><<<(unsigned)blocks, threads, 0, stream>>>(t); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace sb
