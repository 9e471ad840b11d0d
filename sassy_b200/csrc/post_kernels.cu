// Kernels that run after the scan on the (small) candidate list:
//  * minima_kernel : local-minima selection on the sorted candidates
//  * trace_kernel  : one thread per selected end position -> Match + CIGAR ops
// Their per-thread logic lives in scan_core.cuh (shared with the host emulator).
#include "kernels.cuh"

namespace sb {
namespace {

__global__ void minima_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ cost, uint64_t n,
                              uint8_t* __restrict__ flags, bool all_minima) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  flags[i] = select_candidate(keys, cost, i, n, all_minima) ? 1 : 0;
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
