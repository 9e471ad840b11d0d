// Host->device transport of Dna texts.
//
// For a single pattern the scan runs at several TB/s while a pinned host->device copy
// runs at ~55 GB/s, so a search over a host text is bounded by PCIe.  The Dna profile
// only ever looks at bits 1-2 of a text byte ((c >> 1) & 3, reference
// src/profiles/dna.rs:19-23,26-40) and compares case-insensitively in the traceback
// (dna.rs:48-50), so an ACGT text (any case) is fully described by 2 bits per character.
// Host threads squeeze the text to 2 bits per character into pinned memory while earlier
// chunks are already in flight over PCIe; a streaming kernel expands it back to one
// canonical upper-case byte per character in HBM, and everything downstream (scan,
// traceback) runs unchanged on bytes.  A text holding any byte outside ACGTacgt is sent
// unpacked instead, so out-of-alphabet behaviour stays what it is on the byte path.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__linux__)
#include <pthread.h>
#include <sched.h>
#endif

#include "dna_pack.h"
#include "transport.h"

namespace sb {

// ---------------------------------------------------------------------------
// host side

namespace {

bool pack_range(const uint8_t* src, uint8_t* dst, size_t n, bool stream) { return dna_pack(src, dst, n, -1, stream); }

}  // namespace

struct PackPool::Impl {
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv_job, cv_done;
  // current job
  const uint8_t* src = nullptr;
  uint8_t* dst = nullptr;
  size_t n = 0, chunk = 0, nchunks = 0, ring = 0;
  // unclaimed chunks [lo, hi): lo in the low 32 bits (workers take from the front), hi in the
  // high 32 bits (the caller takes from the back)
  std::atomic<uint64_t> range{0};
  std::atomic<size_t> released{0};
  std::unique_ptr<std::atomic<uint8_t>[]> done;  // per chunk: 0 pending, 1 ok, 2 saw a foreign byte
  uint64_t generation = 0;
  size_t active = 0;
  bool quit = false;

  void worker() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_job.wait(lk, [&] { return quit || generation != seen; });
        if (quit) return;
        seen = generation;
      }
      run_chunks();
      {
        std::lock_guard<std::mutex> lk(mu);
        if (--active == 0) cv_done.notify_all();
      }
    }
  }

  bool take_front(size_t& c) {
    uint64_t r = range.load(std::memory_order_relaxed);
    for (;;) {
      const uint32_t lo = (uint32_t)r, hi = (uint32_t)(r >> 32);
      if (lo >= hi) return false;
      if (range.compare_exchange_weak(r, (r & 0xFFFFFFFF00000000ull) | (uint64_t)(lo + 1), std::memory_order_relaxed)) {
        c = lo;
        return true;
      }
    }
  }

  void run_chunks() {
    size_t c;
    while (take_front(c)) {
      if (ring) {
        // the previous user of this slot must have been copied out; cancel() empties the range and
        // releases everything
        while (c >= released.load(std::memory_order_acquire) + ring)
          std::this_thread::sleep_for(std::chrono::microseconds(10));
      }
      const size_t off = c * chunk;
      const size_t len = off + chunk <= n ? chunk : n - off;
      const size_t slot = ring ? c % ring : c;
      const bool ok = pack_range(src + off, dst + slot * (chunk >> 2), len, /*stream=*/ring == 0);
      done[c].store(ok ? 1 : 2, std::memory_order_release);
    }
  }
};

PackPool::PackPool(int threads) : impl_(new Impl) {
  if (threads < 1) threads = 1;
  // the calling thread only feeds the copy engine, so all `threads` are workers; they are
  // spread over the allowed CPUs up front (freshly created threads otherwise share a CPU for
  // a long time on the virtualised hosts this runs on)
  std::vector<int> cpus;
#if defined(__linux__)
  cpu_set_t allowed;
  if (sched_getaffinity(0, sizeof allowed, &allowed) == 0)
    for (int c = 0; c < CPU_SETSIZE; c++)
      if (CPU_ISSET(c, &allowed)) cpus.push_back(c);
#endif
  for (int i = 0; i < threads; i++) {
    impl_->workers.emplace_back([this] { impl_->worker(); });
#if defined(__linux__)
    if (!cpus.empty()) {
      cpu_set_t one;
      CPU_ZERO(&one);
      CPU_SET(cpus[i % cpus.size()], &one);
      pthread_setaffinity_np(impl_->workers.back().native_handle(), sizeof one, &one);
    }
#endif
  }
}

PackPool::~PackPool() {
  {
    std::lock_guard<std::mutex> lk(impl_->mu);
    impl_->quit = true;
  }
  impl_->cv_job.notify_all();
  for (auto& t : impl_->workers) t.join();
  delete impl_;
}

int PackPool::threads() const { return (int)impl_->workers.size(); }

void PackPool::start(const uint8_t* src, uint8_t* dst, size_t n, size_t chunk, size_t ring) {
  Impl& s = *impl_;
  std::lock_guard<std::mutex> lk(s.mu);
  s.src = src, s.dst = dst, s.n = n, s.chunk = chunk, s.ring = ring;
  s.nchunks = (n + chunk - 1) / chunk;
  s.range.store((uint64_t)s.nchunks << 32);
  s.released.store(0);
  s.done.reset(new std::atomic<uint8_t>[s.nchunks ? s.nchunks : 1]);
  for (size_t c = 0; c < s.nchunks; c++) s.done[c].store(0);
  s.active = s.workers.size();
  s.generation++;
  s.cv_job.notify_all();
}

size_t PackPool::chunks() const { return impl_->nchunks; }

size_t PackPool::packed_end() const { return (size_t)(impl_->range.load(std::memory_order_acquire) >> 32); }

int PackPool::chunk_state(size_t c) const { return impl_->done[c].load(std::memory_order_acquire); }

bool PackPool::wait_chunk(size_t c) {
  uint8_t v;
  // sleep rather than spin: the workers own every core
  while ((v = impl_->done[c].load(std::memory_order_acquire)) == 0)
    std::this_thread::sleep_for(std::chrono::microseconds(30));
  return v == 1;
}

bool PackPool::claim_tail(size_t* c) {
  uint64_t r = impl_->range.load(std::memory_order_relaxed);
  for (;;) {
    const uint32_t lo = (uint32_t)r, hi = (uint32_t)(r >> 32);
    if (lo >= hi) return false;
    if (impl_->range.compare_exchange_weak(r, ((uint64_t)(hi - 1) << 32) | lo, std::memory_order_acq_rel)) {
      *c = hi - 1;
      return true;
    }
  }
}

void PackPool::release(size_t upto) {
  if (upto > impl_->released.load(std::memory_order_relaxed)) impl_->released.store(upto, std::memory_order_release);
}

void PackPool::cancel() {
  uint64_t r = impl_->range.load(std::memory_order_relaxed);
  for (;;) {
    const uint32_t lo = (uint32_t)r;
    if (impl_->range.compare_exchange_weak(r, ((uint64_t)lo << 32) | lo, std::memory_order_acq_rel)) break;
  }
  impl_->released.store(~(size_t)0 >> 1, std::memory_order_release);
}

void PackPool::finish() {
  std::unique_lock<std::mutex> lk(impl_->mu);
  impl_->cv_done.wait(lk, [&] { return impl_->active == 0; });
}

// ---------------------------------------------------------------------------
// device side

namespace {

// 16 packed bytes (64 characters: 8 groups of two plane bytes) per thread -> four 16-byte stores
// of canonical bytes (dna_pack.h: dna_unpack4).
__global__ void unpack_dna_kernel(const uint4* __restrict__ packed, uint4* __restrict__ out, size_t n_vec) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_vec) return;
  const uint4 p = __ldg(packed + i);
  const uint32_t w[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const uint32_t g0 = w[j] & 0xFFFFu, g1 = w[j] >> 16;  // characters 16 j .. + 7 and + 8 .. + 15
    out[i * 4 + j] = make_uint4(dna_unpack4(g0, 0), dna_unpack4(g0, 1), dna_unpack4(g1, 0), dna_unpack4(g1, 1));
  }
}

}  // namespace

cudaError_t launch_unpack_dna(const uint8_t* packed, uint8_t* out, size_t n_chars, cudaStream_t stream) {
  const size_t n_vec = (n_chars + 63) / 64;  // whole 64-character groups; the text buffer is padded
  if (n_vec == 0) return cudaSuccess;
  const unsigned threads = 256;
  const size_t blocks = (n_vec + threads - 1) / threads;
  unpack_dna_kernel<<<(unsigned)blocks, threads, 0, stream>>>(reinterpret_cast<const uint4*>(packed),
                                                              reinterpret_cast<uint4*>(out), n_vec);
  return cudaGetLastError();
}

}  // namespace sb
