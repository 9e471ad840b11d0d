// Multi-GPU match gather over peer memory (NVLink / NVSwitch), fused behind the traceback.
//
// The reference has no distributed layer; its multi-pattern / multi-text fan-out ends in the
// concatenation of per-task match lists (src/search.rs:531-603,1519-1549).  Here one process
// drives one GPU, every rank searches its shard, and the match records of all ranks have to
// reach every rank.  Instead of copying the records to the host, back to the device and through
// an NCCL all-gather, the rank's post-processing leaves them in its own slot of a receive
// buffer and push_kernel stores them straight into the same slot of every PEER's receive
// buffer (plain st.global on CUDA-IPC mapped peer pointers), followed by a system-scope release
// of a step flag; collect_kernel acquires the flags of all ranks and moves the used prefix of
// every slot into pinned host memory.  No collective library call on the data path, one host
// synchronisation per search.
//
// Receive buffer of a rank (device memory, CUDA IPC exported):
//   [parity 0|1][source rank 0..world) : slot = header (64 B) | cap records (32 B) | cap x max_ops_words ops
//   flags[2][world] (u64): step number of the last completed push of that source rank
// Steps alternate between the two parities, so a rank that runs one step ahead never overwrites
// data a slower peer is still collecting (it cannot run two steps ahead: finishing step s+1
// needs the peer's flag of step s+1, which the peer sends after it has left step s).
#include "peer_gather.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <stdexcept>

namespace sb {

namespace {

struct SlotHeader {
  unsigned long long count;
  unsigned long long text_n;
  uint32_t ops_words;
  uint32_t overflow;
  unsigned long long step;
  unsigned long long user;  // caller-defined (e.g. number of patterns of the source rank)
  unsigned long long pad[3];
};
static_assert(sizeof(SlotHeader) == 64, "slot header is 64 bytes");

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

struct PushArgs {
  uint8_t* peer[kMaxPeers];  // receive buffer of every rank, as mapped in this process
  uint32_t world, rank;
  uint64_t slot_off;         // offset of slot [parity][rank] (same in every receive buffer)
  uint64_t flag_off;         // offset of flags[parity][rank]
  uint64_t cap, max_ops_words;
  uint32_t ops_words;
  uint32_t force_overflow;
  unsigned long long step, text_n, user;
  const unsigned long long* counts;  // [0] candidates [1] selected [2] prefilter hits [3] big
  unsigned long long cand_cap, hit_cap;
};

// One block per destination rank: stores the used prefix of this rank's slot into the same slot
// of the destination's receive buffer, then the header, then releases the step flag there.
__global__ void __launch_bounds__(256) push_kernel(const PushArgs a) {
  const uint32_t dst_rank = blockIdx.x;
  const unsigned long long nsel = a.counts[1];
  const bool overflow = a.force_overflow || a.counts[3] != 0 || a.counts[0] > a.cand_cap ||
                        a.counts[2] > a.hit_cap || nsel > a.cap;
  const unsigned long long count = overflow ? 0 : nsel;
  const uint8_t* src = a.peer[a.rank] + a.slot_off;
  uint8_t* dst = a.peer[dst_rank] + a.slot_off;
  if (dst_rank != a.rank) {
    const uint4* s4 = reinterpret_cast<const uint4*>(src + sizeof(SlotHeader));
    uint4* d4 = reinterpret_cast<uint4*>(dst + sizeof(SlotHeader));
    for (unsigned long long i = threadIdx.x; i < count * 2; i += blockDim.x) d4[i] = s4[i];  // 32-byte records
    const uint64_t ops_off = sizeof(SlotHeader) + a.cap * 32;
    const uint32_t* so = reinterpret_cast<const uint32_t*>(src + ops_off);
    uint32_t* dop = reinterpret_cast<uint32_t*>(dst + ops_off);
    for (unsigned long long i = threadIdx.x; i < count * a.ops_words; i += blockDim.x) dop[i] = so[i];
  }
  if (threadIdx.x == 0) {
    SlotHeader h;
    memset(&h, 0, sizeof h);
    h.count = count;
    h.text_n = a.text_n;
    h.ops_words = a.ops_words;
    h.overflow = overflow ? 1u : 0u;
    h.step = a.step;
    h.user = a.user;
    *reinterpret_cast<SlotHeader*>(dst) = h;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0)
    st_release_sys(reinterpret_cast<unsigned long long*>(a.peer[dst_rank] + a.flag_off), a.step);
}

struct CollectArgs {
  const uint8_t* local;  // this rank's receive buffer
  uint8_t* host;         // pinned host mirror: world slots back to back
  uint32_t world;
  uint64_t slot_bytes, parity_off, flags_off;  // flags_off: flags[parity][0]
  uint64_t cap;
  unsigned long long step;
  unsigned long long timeout_ns;
};

// One block per source rank: waits for that rank's step flag, then moves header and used prefix
// of its slot into pinned host memory.
__global__ void __launch_bounds__(256) collect_kernel(const CollectArgs a) {
  const uint32_t r = blockIdx.x;
  __shared__ int timed_out;
  if (threadIdx.x == 0) {
    timed_out = 0;
    const unsigned long long* flag = reinterpret_cast<const unsigned long long*>(a.local + a.flags_off) + r;
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_sys(flag) < a.step) {
      if (globaltimer_ns() - t0 > a.timeout_ns) {
        timed_out = 1;  // a peer never arrived: report instead of hanging the GPU
        break;
      }
      __nanosleep(200);
    }
  }
  __syncthreads();
  const uint8_t* slot = a.local + a.parity_off + (uint64_t)r * a.slot_bytes;
  uint8_t* out = a.host + (uint64_t)r * a.slot_bytes;
  SlotHeader h;
  if (timed_out) {
    memset(&h, 0, sizeof h);
    h.overflow = 2;  // time-out marker
    h.step = a.step;
  } else {
    const uint4* hp = reinterpret_cast<const uint4*>(slot);
    uint4* hq = reinterpret_cast<uint4*>(&h);
#pragma unroll
    for (int i = 0; i < 4; i++) hq[i] = __ldcg(hp + i);
  }
  const unsigned long long count = h.count <= a.cap ? h.count : 0;
  const uint4* s4 = reinterpret_cast<const uint4*>(slot + sizeof(SlotHeader));
  uint4* d4 = reinterpret_cast<uint4*>(out + sizeof(SlotHeader));
  for (unsigned long long i = threadIdx.x; i < count * 2; i += blockDim.x) d4[i] = __ldcg(s4 + i);
  const uint64_t ops_off = sizeof(SlotHeader) + a.cap * 32;
  const uint32_t* so = reinterpret_cast<const uint32_t*>(slot + ops_off);
  uint32_t* dop = reinterpret_cast<uint32_t*>(out + ops_off);
  for (unsigned long long i = threadIdx.x; i < count * h.ops_words; i += blockDim.x) dop[i] = __ldcg(so + i);
  if (threadIdx.x == 0) *reinterpret_cast<SlotHeader*>(out) = h;
}

void check(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

}  // namespace

PeerGather::PeerGather(int device, int world, int rank, size_t cap, size_t max_ops_words)
    : device_(device), world_(world), rank_(rank), cap_(cap), max_ops_words_(max_ops_words) {
  if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world) throw std::invalid_argument("bad world/rank");
  if (cap == 0) throw std::invalid_argument("capacity must be positive");
  check(cudaSetDevice(device_), "cudaSetDevice");
  // Ranks may legitimately be far apart (a multi-GB host text being staged, a candidate-buffer
  // retry on one shard, first-launch module load): the wait for a peer's step flag is long by
  // default and configurable.  A time-out means a rank is gone, not slow: the search fails on the
  // ranks that waited and the PeerGather must not be used again (Engine drops it).
  if (const char* e = getenv("SASSY_B200_GATHER_TIMEOUT_S")) {
    const double s = atof(e);
    if (s > 0) timeout_ns_ = (unsigned long long)(s * 1e9);
  }
  slot_bytes_ = (sizeof(SlotHeader) + cap_ * 32 + cap_ * max_ops_words_ * 4 + 255) & ~(size_t)255;
  flags_off_ = 2 * (size_t)world_ * slot_bytes_;
  bytes_ = flags_off_ + 2 * (size_t)world_ * sizeof(unsigned long long);
  check(cudaMalloc((void**)&local_, bytes_), "cudaMalloc(receive buffer)");
  check(cudaMemset(local_, 0, bytes_), "cudaMemset");
  check(cudaHostAlloc((void**)&host_, (size_t)world_ * slot_bytes_, cudaHostAllocDefault), "cudaHostAlloc");
  memset(host_, 0, (size_t)world_ * slot_bytes_);
  for (int r = 0; r < kMaxPeers; r++) peer_[r] = nullptr;
  peer_[rank_] = local_;
  check(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
}

PeerGather::~PeerGather() {
  cudaSetDevice(device_);
  cudaDeviceSynchronize();
  for (int r = 0; r < world_; r++)
    if (r != rank_ && peer_[r]) cudaIpcCloseMemHandle(peer_[r]);
  if (local_) cudaFree(local_);
  if (host_) cudaFreeHost(host_);
}

void PeerGather::export_handle(uint8_t out[64]) const {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  cudaIpcMemHandle_t h;
  check(cudaSetDevice(device_), "cudaSetDevice");
  check(cudaIpcGetMemHandle(&h, local_), "cudaIpcGetMemHandle");
  memcpy(out, &h, 64);
}

void PeerGather::connect(const uint8_t* handles) {
  check(cudaSetDevice(device_), "cudaSetDevice");
  for (int r = 0; r < world_; r++) {
    if (r == rank_) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)r * 64, 64);
    void* p = nullptr;
    check(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle (peer receive buffer)");
    peer_[r] = static_cast<uint8_t*>(p);
  }
  connected_ = true;
}

uint8_t* PeerGather::local_records(size_t* ops_offset) {
  // own slot of the parity the NEXT exchange uses
  const unsigned long long next = step_ + 1;
  uint8_t* slot = local_ + (next & 1) * (size_t)world_ * slot_bytes_ + (size_t)rank_ * slot_bytes_;
  if (ops_offset) *ops_offset = cap_ * 32;
  return slot + sizeof(SlotHeader);
}

cudaError_t PeerGather::collect(unsigned long long step, cudaStream_t stream) {
  const size_t parity_off = (step & 1) * (size_t)world_ * slot_bytes_;
  CollectArgs c;
  memset(&c, 0, sizeof c);
  c.local = local_;
  c.host = host_;
  c.world = (uint32_t)world_;
  c.slot_bytes = slot_bytes_;
  c.parity_off = parity_off;
  c.flags_off = flags_off_ + (step & 1) * (size_t)world_ * sizeof(unsigned long long);
  c.cap = cap_;
  c.step = step;
  c.timeout_ns = timeout_ns_;
  collect_kernel<<<world_, 256, 0, stream>>>(c);
  collected_step_ = step;
  return cudaGetLastError();
}

cudaError_t PeerGather::flush(cudaStream_t stream) {
  if (!pipelined_ || step_ == 0 || collected_step_ == step_) return cudaSuccess;
  return collect(step_, stream);
}

cudaError_t PeerGather::exchange(const unsigned long long* d_counts, unsigned long long cand_cap,
                                 unsigned long long hit_cap, uint32_t ops_words, bool force_overflow,
                                 unsigned long long text_n, unsigned long long user, cudaStream_t stream) {
  if (!connected_ && world_ > 1) return cudaErrorNotReady;
  if (pipelined_ && step_ >= 1 && collected_step_ != step_) {
    // the previous step: its flags were released a whole search ago
    cudaError_t e = collect(step_, stream);
    if (e != cudaSuccess) return e;
  }
  step_++;
  const size_t parity_off = (step_ & 1) * (size_t)world_ * slot_bytes_;
  PushArgs p;
  memset(&p, 0, sizeof p);
  for (int r = 0; r < world_; r++) p.peer[r] = peer_[r];
  p.world = (uint32_t)world_;
  p.rank = (uint32_t)rank_;
  p.slot_off = parity_off + (size_t)rank_ * slot_bytes_;
  p.flag_off = flags_off_ + ((step_ & 1) * (size_t)world_ + (size_t)rank_) * sizeof(unsigned long long);
  p.cap = cap_;
  p.max_ops_words = max_ops_words_;
  p.ops_words = ops_words;
  p.force_overflow = (force_overflow || ops_words > max_ops_words_) ? 1u : 0u;
  p.step = step_;
  p.text_n = text_n;
  p.user = user;
  p.counts = d_counts;
  p.cand_cap = cand_cap;
  p.hit_cap = hit_cap;
  push_kernel<<<world_, 256, 0, stream>>>(p);
  if (pipelined_) return cudaGetLastError();
  return collect(step_, stream);
}

bool PeerGather::ok() const {
  if (collected_step_ == 0) return false;
  for (int r = 0; r < world_; r++) {
    const SlotHeader* h = reinterpret_cast<const SlotHeader*>(host_ + (size_t)r * slot_bytes_);
    if (h->overflow || h->step != collected_step_) return false;
  }
  return true;
}

bool PeerGather::timed_out() const {
  for (int r = 0; r < world_; r++)
    if (reinterpret_cast<const SlotHeader*>(host_ + (size_t)r * slot_bytes_)->overflow == 2) return true;
  return false;
}

PeerGather::Slot PeerGather::slot(int r) const {
  const uint8_t* base = host_ + (size_t)r * slot_bytes_;
  const SlotHeader* h = reinterpret_cast<const SlotHeader*>(base);
  Slot s;
  s.count = h->count;
  s.text_n = h->text_n;
  s.ops_words = h->ops_words;
  s.user = h->user;
  s.records = reinterpret_cast<const GpuMatch*>(base + sizeof(SlotHeader));
  s.ops = reinterpret_cast<const uint32_t*>(base + sizeof(SlotHeader) + cap_ * 32);
  return s;
}

}  // namespace sb
