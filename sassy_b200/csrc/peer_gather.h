// Multi-GPU gather of match records over peer memory (see peer_gather.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "scan_core.cuh"

namespace sb {

constexpr int kMaxPeers = 16;

class PeerGather {
 public:
  // cap = records per rank and step that fit a slot; max_ops_words = 32-bit words of packed
  // CIGAR ops per record ((m + k + 1 + 15) / 16 of the longest pattern searched).
  PeerGather(int device, int world, int rank, size_t cap, size_t max_ops_words);
  ~PeerGather();
  PeerGather(const PeerGather&) = delete;
  PeerGather& operator=(const PeerGather&) = delete;

  int world() const { return world_; }
  int rank() const { return rank_; }
  size_t cap() const { return cap_; }
  size_t max_ops_words() const { return max_ops_words_; }

  void export_handle(uint8_t out[64]) const;  // CUDA IPC handle of this rank's receive buffer
  void connect(const uint8_t* handles);       // world x 64 bytes, in rank order

  // Device pointer where the traceback of the NEXT exchange has to leave this rank's records
  // (GpuMatch[cap]); the packed ops follow at *ops_offset bytes, `ops_words` words per record.
  uint8_t* local_records(size_t* ops_offset);
  // Queues push + collect on `stream`.  d_counts = the engine's device counters
  // ([0] candidates, [1] selected, [2] prefilter hits, [3] list too long for the fast tail);
  // the push marks the slot `overflow` when the records in the slot are not this step's
  // complete result.  After the stream is synchronised ok() / slot() describe the step.
  cudaError_t exchange(const unsigned long long* d_counts, unsigned long long cand_cap, unsigned long long hit_cap,
                       uint32_t ops_words, bool force_overflow, unsigned long long text_n, unsigned long long user,
                       cudaStream_t stream);
  // Pipelined mode (all ranks alike, before the first exchange): exchange() of step s first collects
  // step s - 1 -- whose records the peers pushed a whole search ago, so the wait for the slowest
  // rank disappears from the step -- and then pushes step s; ok() / slot() then describe step s - 1.
  // flush() collects the last pushed step.  The ranks may drift one step apart (not more: pushing
  // step s needs every peer's flag of step s - 1), which is what the two slot parities allow.
  void set_pipelined(bool on) { pipelined_ = on; }
  bool pipelined() const { return pipelined_; }
  bool has_result() const { return collected_step_ > 0; }  // a collected step is in the host mirror
  cudaError_t flush(cudaStream_t stream);
  bool ok() const;         // every rank delivered a complete result for the collected step
  bool timed_out() const;  // some rank did not arrive within the time-out
  // After a time-out the step counters of the ranks may differ: no further exchange is possible.
  void mark_broken() { broken_ = true; }
  bool broken() const { return broken_; }
  struct Slot {
    unsigned long long count, text_n, user;
    uint32_t ops_words;
    const GpuMatch* records;
    const uint32_t* ops;
  };
  Slot slot(int r) const;  // host view of rank r's records of the last step

 private:
  int device_, world_, rank_;
  size_t cap_, max_ops_words_;
  size_t slot_bytes_ = 0, flags_off_ = 0, bytes_ = 0;
  uint8_t* local_ = nullptr;
  uint8_t* host_ = nullptr;
  uint8_t* peer_[kMaxPeers];
  bool connected_ = false;
  unsigned long long step_ = 0;
  unsigned long long timeout_ns_ = 120ull * 1000 * 1000 * 1000;  // SASSY_B200_GATHER_TIMEOUT_S
  bool broken_ = false;
  bool pipelined_ = false;
  unsigned long long collected_step_ = 0;  // step whose records the host mirror holds (after a sync)
  cudaError_t collect(unsigned long long step, cudaStream_t stream);
};

}  // namespace sb
