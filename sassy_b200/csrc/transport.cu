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

#if defined(__x86_64__)
#include <immintrin.h>
#endif
#if defined(__linux__)
#include <pthread.h>
#include <sched.h>
#endif

#include "transport.h"

namespace sb {

// ---------------------------------------------------------------------------
// host side

namespace {

// 4 characters -> 1 byte, character i at bits 2i..2i+1, code = (c >> 1) & 3.
// Returns false if a byte outside ACGTacgt was seen (the output is still written).
bool pack_scalar(const uint8_t* src, uint8_t* dst, size_t n) {
  bool ok = true;
  size_t i = 0;
  for (; i + 4 <= n; i += 4) {
    uint8_t b = 0;
    for (int j = 0; j < 4; j++) {
      const uint8_t c = src[i + j], u = c & 0xDF;
      ok &= (u == 'A') | (u == 'C') | (u == 'G') | (u == 'T');
      b |= (uint8_t)(((c >> 1) & 3) << (2 * j));
    }
    dst[i >> 2] = b;
  }
  if (i < n) {
    uint8_t b = 0;
    for (int j = 0; i + j < n; j++) {
      const uint8_t c = src[i + j], u = c & 0xDF;
      ok &= (u == 'A') | (u == 'C') | (u == 'G') | (u == 'T');
      b |= (uint8_t)(((c >> 1) & 3) << (2 * j));
    }
    dst[i >> 2] = b;
  }
  return ok;
}

#if defined(__x86_64__)
// 32 characters per iteration: codes = (v >> 1) & 3, then two multiply-adds fold four
// codes into one byte per 32-bit lane (c0 + 4 c1 + 16 c2 + 64 c3), a byte shuffle gathers them.
__attribute__((target("avx2"))) bool pack_avx2(const uint8_t* src, uint8_t* dst, size_t n) {
  const __m256i up = _mm256_set1_epi8((char)0xDF), three = _mm256_set1_epi8(3);
  const __m256i cA = _mm256_set1_epi8('A'), cC = _mm256_set1_epi8('C'), cG = _mm256_set1_epi8('G'),
                cT = _mm256_set1_epi8('T');
  const __m256i m1 = _mm256_set1_epi16(0x0401), m2 = _mm256_set1_epi32(0x00100001);
  const __m256i gather = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,  //
                                          0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
  __m256i all_ok = _mm256_set1_epi8((char)0xFF);
  size_t i = 0;
  for (; i + 32 <= n; i += 32) {
    const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
    const __m256i u = _mm256_and_si256(v, up);
    const __m256i ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(u, cA), _mm256_cmpeq_epi8(u, cC)),
                                       _mm256_or_si256(_mm256_cmpeq_epi8(u, cG), _mm256_cmpeq_epi8(u, cT)));
    all_ok = _mm256_and_si256(all_ok, ok);
    const __m256i codes = _mm256_and_si256(_mm256_srli_epi16(v, 1), three);
    const __m256i pairs = _mm256_maddubs_epi16(codes, m1);  // c0 + 4 c1 per 16-bit lane
    const __m256i quads = _mm256_madd_epi16(pairs, m2);     // + 16 (c2 + 4 c3) per 32-bit lane
    const __m256i g = _mm256_shuffle_epi8(quads, gather);
    const uint32_t lo = (uint32_t)_mm256_cvtsi256_si32(g);
    const uint32_t hi = (uint32_t)_mm256_extract_epi32(g, 4);
    const uint64_t packed = (uint64_t)lo | ((uint64_t)hi << 32);
    memcpy(dst + (i >> 2), &packed, 8);
  }
  bool ok = _mm256_movemask_epi8(all_ok) == -1;
  if (i < n) ok &= pack_scalar(src + i, dst + (i >> 2), n - i);
  return ok;
}

// 64 characters per iteration; VPMOVDB truncates the sixteen 32-bit lanes to 16 bytes.
__attribute__((target("avx512f,avx512bw"))) bool pack_avx512(const uint8_t* src, uint8_t* dst, size_t n) {
  const __m512i up = _mm512_set1_epi8((char)0xDF), three = _mm512_set1_epi8(3);
  const __m512i cA = _mm512_set1_epi8('A'), cC = _mm512_set1_epi8('C'), cG = _mm512_set1_epi8('G'),
                cT = _mm512_set1_epi8('T');
  const __m512i m1 = _mm512_set1_epi16(0x0401), m2 = _mm512_set1_epi32(0x00100001);
  __mmask64 all_ok = ~(__mmask64)0;
  size_t i = 0;
  for (; i + 64 <= n; i += 64) {
    _mm_prefetch(reinterpret_cast<const char*>(src + i + 1024), _MM_HINT_NTA);
    const __m512i v = _mm512_loadu_si512(src + i);
    const __m512i u = _mm512_and_si512(v, up);
    all_ok &= _mm512_cmpeq_epi8_mask(u, cA) | _mm512_cmpeq_epi8_mask(u, cC) | _mm512_cmpeq_epi8_mask(u, cG) |
              _mm512_cmpeq_epi8_mask(u, cT);
    const __m512i codes = _mm512_and_si512(_mm512_srli_epi16(v, 1), three);
    const __m512i pairs = _mm512_maddubs_epi16(codes, m1);
    const __m512i quads = _mm512_madd_epi16(pairs, m2);
    // dst chunks start on 16-byte boundaries: streaming store, no read-for-ownership
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + (i >> 2)), _mm512_cvtepi32_epi8(quads));
  }
  _mm_sfence();
  bool ok = all_ok == ~(__mmask64)0;
  if (i < n) ok &= pack_scalar(src + i, dst + (i >> 2), n - i);
  return ok;
}
#endif

bool pack_range(const uint8_t* src, uint8_t* dst, size_t n) {
#if defined(__x86_64__)
  static const int level = __builtin_cpu_supports("avx512bw") ? 2 : (__builtin_cpu_supports("avx2") ? 1 : 0);
  if (level == 2) return pack_avx512(src, dst, n);
  if (level == 1) return pack_avx2(src, dst, n);
#endif
  return pack_scalar(src, dst, n);
}

}  // namespace

struct PackPool::Impl {
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv_job, cv_done;
  // current job
  const uint8_t* src = nullptr;
  uint8_t* dst = nullptr;
  size_t n = 0, chunk = 0, nchunks = 0;
  std::atomic<size_t> next{0};
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

  void run_chunks() {
    for (;;) {
      const size_t c = next.fetch_add(1, std::memory_order_relaxed);
      if (c >= nchunks) return;
      const size_t off = c * chunk;
      const size_t len = off + chunk <= n ? chunk : n - off;
      const bool ok = pack_range(src + off, dst + (off >> 2), len);
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

void PackPool::start(const uint8_t* src, uint8_t* dst, size_t n, size_t chunk) {
  Impl& s = *impl_;
  std::lock_guard<std::mutex> lk(s.mu);
  s.src = src, s.dst = dst, s.n = n, s.chunk = chunk;
  s.nchunks = (n + chunk - 1) / chunk;
  s.next.store(0);
  s.done.reset(new std::atomic<uint8_t>[s.nchunks ? s.nchunks : 1]);
  for (size_t c = 0; c < s.nchunks; c++) s.done[c].store(0);
  s.active = s.workers.size();
  s.generation++;
  s.cv_job.notify_all();
}

size_t PackPool::chunks() const { return impl_->nchunks; }

bool PackPool::wait_chunk(size_t c) {
  uint8_t v;
  // sleep rather than spin: the workers own every core
  while ((v = impl_->done[c].load(std::memory_order_acquire)) == 0)
    std::this_thread::sleep_for(std::chrono::microseconds(30));
  return v == 1;
}

void PackPool::finish() {
  std::unique_lock<std::mutex> lk(impl_->mu);
  impl_->cv_done.wait(lk, [&] { return impl_->active == 0; });
}

// ---------------------------------------------------------------------------
// device side

namespace {

// 16 packed bytes (64 characters) per thread -> four 16-byte stores of canonical bytes.
// Code -> byte through PRMT on the 4-byte table "ACTG" (A=0,C=1,T=2,G=3 as (c>>1)&3).
__global__ void unpack_dna_kernel(const uint4* __restrict__ packed, uint4* __restrict__ out, size_t n_vec) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_vec) return;
  const uint4 p = __ldg(packed + i);
  const uint32_t w[4] = {p.x, p.y, p.z, p.w};
  const uint32_t table = 0x47544341u;  // 'A','C','T','G'
#pragma unroll
  for (int j = 0; j < 4; j++) {
    uint32_t o[4];
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const uint32_t byte = (w[j] >> (8 * b)) & 0xFFu;
      // spread the four 2-bit codes to the four selector nibbles
      const uint32_t sel = (byte & 3u) | ((byte & 0xCu) << 2) | ((byte & 0x30u) << 4) | ((byte & 0xC0u) << 6);
      o[b] = __byte_perm(table, 0u, sel);
    }
    out[i * 4 + j] = make_uint4(o[0], o[1], o[2], o[3]);
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
