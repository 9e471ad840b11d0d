// Host->device transport encoding for Dna texts (2 bits per character), see transport.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace sb {

// A small persistent thread pool that packs a text chunk by chunk; the caller consumes the
// chunks in order (to feed the copy engine) while later chunks are still being packed.
class PackPool {
 public:
  explicit PackPool(int threads);
  ~PackPool();
  PackPool(const PackPool&) = delete;
  PackPool& operator=(const PackPool&) = delete;

  int threads() const;
  // Packs src[0..n) into dst[0..ceil(n/4)) in chunks of `chunk` characters (multiple of 64).
  void start(const uint8_t* src, uint8_t* dst, size_t n, size_t chunk);
  size_t chunks() const;
  // Blocks until chunk c is packed; false if it held a byte outside ACGTacgt.
  bool wait_chunk(size_t c);
  // Blocks until every worker is idle again (required before the next start()).
  void finish();

 private:
  struct Impl;
  Impl* impl_;
};

// Expands ceil(n_chars/64)*16 packed bytes to canonical upper-case bytes (whole groups of 64).
cudaError_t launch_unpack_dna(const uint8_t* packed, uint8_t* out, size_t n_chars, cudaStream_t stream);

}  // namespace sb
