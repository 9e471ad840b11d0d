// Host-side transport microbenchmark (run on the GPU box): how fast do the packing threads read
// a 3 GB text, how fast is a plain pinned copy, and how do the two share the host memory system?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/pack_bench.cu sassy_b200/csrc/transport.cu -o /tmp/pack_bench
#include <cuda_runtime.h>
#include <immintrin.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <thread>
#include <vector>

#include "../sassy_b200/csrc/transport.h"

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

__attribute__((target("avx2"))) static uint64_t read_sum(const uint8_t* p, size_t n) {
  __m256i acc = _mm256_setzero_si256();
  for (size_t i = 0; i + 32 <= n; i += 32) acc = _mm256_xor_si256(acc, _mm256_loadu_si256((const __m256i*)(p + i)));
  uint64_t t[4];
  memcpy(t, &acc, 32);
  return t[0] ^ t[1] ^ t[2] ^ t[3];
}

int main(int argc, char** argv) {
  const size_t n = argc > 1 ? strtoull(argv[1], 0, 10) : 3000000000ull;
  uint8_t *text, *packed, *dev;
  cudaHostAlloc((void**)&text, n, cudaHostAllocDefault);
  cudaHostAlloc((void**)&packed, n / 4 + 64, cudaHostAllocDefault);
  cudaMalloc((void**)&dev, n);
  {
    std::vector<std::thread> th;
    const int T = 16;
    for (int t = 0; t < T; t++)
      th.emplace_back([=] {
        uint64_t s = 88172645463325252ull + t;
        for (size_t i = n / T * t; i < (t == T - 1 ? n : n / T * (t + 1)); i++) {
          s ^= s << 13, s ^= s >> 7, s ^= s << 17;
          text[i] = "ACGT"[s & 3];
        }
      });
    for (auto& t : th) t.join();
  }
  printf("hardware_concurrency %u\n", std::thread::hardware_concurrency());
  // plain pinned copy
  for (int rep = 0; rep < 3; rep++) {
    const double t0 = now();
    cudaMemcpy(dev, text, n, cudaMemcpyHostToDevice);
    printf("h2d pinned: %.1f GB/s\n", n / (now() - t0) / 1e9);
  }
  // read-only bandwidth per thread count
  for (int T : {1, 2, 4, 8, 12, 16, 24, 32}) {
    volatile uint64_t sink = 0;
    double best = 1e9;
    for (int rep = 0; rep < 3; rep++) {
      std::vector<std::thread> th;
      const double t0 = now();
      for (int t = 0; t < T; t++) th.emplace_back([&, t] { sink = sink ^ read_sum(text + n / T * t, n / T); });
      for (auto& x : th) x.join();
      best = std::min(best, now() - t0);
    }
    printf("read-only %2d threads: %.1f GB/s\n", T, n / best / 1e9);
  }
  // packing alone, and packing while a pinned copy of the same size runs
  for (int T : {4, 8, 12, 16, 24, 32}) {
    sb::PackPool pool(T);
    for (int with_copy = 0; with_copy < 2; with_copy++) {
      double best = 1e9, best_copy = 0;
      for (int rep = 0; rep < 3; rep++) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0), cudaEventCreate(&e1);
        const double t0 = now();
        pool.start(text, packed, n / 64 * 64, 8ull << 20);
        if (with_copy) {
          cudaEventRecord(e0);
          cudaMemcpyAsync(dev, text, n / 2, cudaMemcpyHostToDevice);
          cudaEventRecord(e1);
        }
        pool.wait_chunk(pool.chunks() - 1);
        pool.finish();
        const double dt = now() - t0;
        if (with_copy) {
          cudaEventSynchronize(e1);
          float ms;
          cudaEventElapsedTime(&ms, e0, e1);
          best_copy = n / 2 / (ms * 1e-3) / 1e9;
        }
        if (dt < best) best = dt;
      }
      printf("pack %2d threads%s: %.1f GB/s", T, with_copy ? " + concurrent h2d" : "", n / best / 1e9);
      if (with_copy) printf("  (h2d meanwhile %.1f GB/s)", best_copy);
      printf("\n");
    }
  }
  return 0;
}
