// Host->device transport encoding for Dna texts (2 bits per character), see transport.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace sb {

// A small persistent thread pool that packs a text chunk by chunk; the caller consumes the
// chunks in order (to feed the copy engine) while later chunks are still being packed.
//
// Staging ring: with `ring` > 0 chunk c is written to slot c % ring of a small buffer that stays in
// the host's caches (plain stores; the copy engine reads the lines from the cache, so the packed
// text never travels to DRAM and back); a worker waits until the caller has release()d the previous
// user of its slot.  With ring == 0 the staging buffer holds the whole packed text (streaming stores).
//
// Work sharing with the copy engine: the workers take chunks from the front of the text, the
// caller may take chunks from the back with claim_tail() and send them as plain bytes while the
// next packed chunk is not ready yet -- the split between packed and plain bytes follows the
// speed of the host cores by itself.
class PackPool {
 public:
  explicit PackPool(int threads);
  ~PackPool();
  PackPool(const PackPool&) = delete;
  PackPool& operator=(const PackPool&) = delete;

  int threads() const;
  // Packs src[0..n) in chunks of `chunk` characters (multiple of 256); chunk c goes to
  // dst + (ring ? c % ring : c) * chunk / 4.
  void start(const uint8_t* src, uint8_t* dst, size_t n, size_t chunk, size_t ring = 0);
  size_t chunks() const;      // chunks of the whole text
  size_t packed_end() const;  // chunks [0, packed_end) are packed by the workers (shrinks with claim_tail)
  // 0: chunk c is not packed yet, 1: packed, 2: packed and held a byte outside ACGTacgt
  int chunk_state(size_t c) const;
  // Blocks until chunk c is packed; false if it held a byte outside ACGTacgt.
  bool wait_chunk(size_t c);
  // The caller takes the last chunk no worker has started: false when none is left.
  bool claim_tail(size_t* c);
  // Chunks [0, upto) have left their ring slots.
  void release(size_t upto);
  // No further chunks are started (used when a foreign byte was found).
  void cancel();
  // Blocks until every worker is idle again (required before the next start()).
  void finish();

 private:
  struct Impl;
  Impl* impl_;
};

// Expands ceil(n_chars/64)*16 packed bytes to canonical upper-case bytes (whole groups of 64).
cudaError_t launch_unpack_dna(const uint8_t* packed, uint8_t* out, size_t n_chars, cudaStream_t stream);

}  // namespace sb
