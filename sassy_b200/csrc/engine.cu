#include "engine.h"

#if defined(__linux__)
#include <sys/prctl.h>
#endif
#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <cub/cub.cuh>
#include <chrono>
#include <thread>

namespace sb {

#define SB_CUDA(expr)                                                                         \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      char buf__[512];                                                                        \
      snprintf(buf__, sizeof buf__, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
               __FILE__, __LINE__);                                                           \
      throw CudaError(buf__);                                                                 \
    }                                                                                         \
  } while (0)

void DevBuf::ensure(size_t bytes) {
  if (bytes <= cap) return;
  release();
  size_t want = std::max(bytes, (size_t)256);
  SB_CUDA(cudaMalloc(&p, want));
  cap = want;
}

void DevBuf::release() {
  if (p) cudaFree(p);
  p = nullptr;
  cap = 0;
}

Engine::Engine(int profile, int device) : profile_(profile), device_(device), variant_(kVariantTma) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    throw CudaError(std::string("sassy_b200 needs a CUDA device (no CPU fallback exists): ") +
                    cudaGetErrorString(e));
  if (device < 0 || device >= ndev) throw CudaError("invalid CUDA device index");
  SB_CUDA(cudaSetDevice(device_));
  cudaDeviceProp prop;
  SB_CUDA(cudaGetDeviceProperties(&prop, device_));
  if (prop.major < 10)
    throw CudaError(std::string("sassy_b200 is built for sm_100a (B200); found ") + prop.name);
  sm_count_ = prop.multiProcessorCount;
  SB_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
  for (auto& ev : ev_) SB_CUDA(cudaEventCreate(&ev));
  ProfileParams pp;
  if (!profile_params(profile_, pp)) throw CudaError("unknown profile");
  nrows_ = pp.nrows, sh0_ = pp.sh0, msk0_ = pp.msk0;
  const char* v = getenv("SASSY_B200_VARIANT");
  if (v && !strcmp(v, "ldg")) variant_ = kVariantLdg;
  const char* tm = getenv("SASSY_B200_TRANSPORT");
  if (tm && !strcmp(tm, "bytes")) transport_mode_ = 0;
  const char* fm = getenv("SASSY_B200_FILTER");
  if (fm && !strcmp(fm, "off")) filter_mode_ = 0;
  if (fm && !strcmp(fm, "force")) filter_mode_ = 2;
  const char* pm = getenv("SASSY_B200_PAIR_MAX_WORDS");
  if (pm) pair_max_words_ = atoi(pm);
  const char* qm = getenv("SASSY_B200_QGRAM");
  if (qm && !strcmp(qm, "0")) qgram_mode_ = 0;
  const char* qsq = getenv("SASSY_B200_QGRAM_SEQ");
  if (qsq && !strcmp(qsq, "0")) qgram_seq_ = false;
  const char* qq = getenv("SASSY_B200_QGRAM_MIN_Q");
  if (qq) qgram_min_q_ = std::max(6, std::min(8, atoi(qq)));
  const char* s2 = getenv("SASSY_B200_SCAN2");
  if (s2 && !strcmp(s2, "0")) scan2_ = false;
  const char* rf = getenv("SASSY_B200_REFINE");
  if (rf) refine_mode_ = std::max(0, std::min(2, atoi(rf)));
  const char* rb = getenv("SASSY_B200_FILTER_ROW_BYTES");
  if (rb) filter_row_bytes_ = atoi(rb);
  const char* fs = getenv("SASSY_B200_FUSE_STRANDS");
  if (fs && !strcmp(fs, "0")) fuse_strands_ = false;
  cudaDriverEntryPointQueryResult qres;
  void* fn = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
      qres == cudaDriverEntryPointSuccess)
    encode_tiled_ = fn;
  if (!encode_tiled_ && variant_ == kVariantTma) throw CudaError("cuTensorMapEncodeTiled not available");
}

namespace {
struct HostTimerState {
  double acc[HostTimers::kCount] = {};
  unsigned long long n[HostTimers::kCount] = {};
  bool on = false;
  HostTimerState() {
    const char* e = getenv("SASSY_B200_HOST_TIMING");
    on = e && atoi(e) != 0;
  }
  ~HostTimerState() {
    if (!on) return;
    static const char* names[HostTimers::kCount] = {"entry -> parameters uploaded", "launches + wait for the GPU",
                                                    "after the last synchronisation", "merge of gathered records",
                                                    "whole C call (sharded search)", "many texts: staging (host copies)",
                                                    "many texts: upload + kernels + tail", "many texts: records -> matches"};
    for (int i = 0; i < HostTimers::kCount; i++)
      if (n[i]) fprintf(stderr, "[sassy_b200 host timing] %-34s %9.1f us avg over %llu calls\n", names[i], acc[i] / n[i], n[i]);
  }
};
HostTimerState g_host_timers;
}  // namespace
bool HostTimers::on() { return g_host_timers.on; }
void HostTimers::add(int which, double us) {
  g_host_timers.acc[which] += us;
  g_host_timers.n[which]++;
}
double HostTimers::now_us() {
  return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

Engine::~Engine() {
  cudaSetDevice(device_);
  if (stream_) cudaStreamSynchronize(stream_);
  for (DevBuf* b : {&keys_, &cost_, &keys2_, &cost2_, &flags_, &sel_, &cubtmp_, &scratch_, &ops_, &out_, &hits_,
                    &d_stage_, &best_, &sel_cost_, &d_texts_, &hits2_, &tiles_})
    b->release();
  if (staged_.d) cudaFree(staged_.d);
  if (h_stage_) cudaFreeHost(h_stage_);
  if (h_pack_) cudaFreeHost(h_pack_);
  for (cudaEvent_t e : copy_ev_) cudaEventDestroy(e);
  if (unpack_stream_) cudaStreamDestroy(unpack_stream_);
  if (h_small_) cudaFreeHost(h_small_);
  if (h_texts_) cudaFreeHost(h_texts_);
  sel_small_.release();
  d_pack_.release();
  delete pool_;
  for (auto& ev : ev_)
    if (ev) cudaEventDestroy(ev);
  if (stream_) cudaStreamDestroy(stream_);
}

DeviceText* Engine::upload_text(const uint8_t* host, uint64_t n) {
  SB_CUDA(cudaSetDevice(device_));
  if (n >= (1ull << kPosBits)) throw CapacityError("text longer than 2^40 bytes is not supported");
  DeviceText* t = new DeviceText;
  t->n = n;
  t->alloc = padded_alloc(n);
  try {
    SB_CUDA(cudaMalloc((void**)&t->d, t->alloc));
    send_text(t->d, t->alloc, host, n);
    SB_CUDA(cudaStreamSynchronize(stream_));
    SB_CUDA(cudaEventElapsedTime(&transfer_ms_, ev_[5], ev_[6]));
    transfer_pending_ = false;
  } catch (...) {
    if (t->d) cudaFree(t->d);
    delete t;
    throw;
  }
  return t;
}

DeviceText* Engine::adopt_device_text(const void* dptr, uint64_t n) {
  SB_CUDA(cudaSetDevice(device_));
  if (n >= (1ull << kPosBits)) throw CapacityError("text longer than 2^40 bytes is not supported");
  DeviceText* t = new DeviceText;
  t->n = n;
  t->alloc = padded_alloc(n);
  try {
    SB_CUDA(cudaMalloc((void**)&t->d, t->alloc));
    if (n) SB_CUDA(cudaMemcpyAsync(t->d, dptr, n, cudaMemcpyDeviceToDevice, stream_));
    SB_CUDA(cudaMemsetAsync(t->d + n, 0, t->alloc - n, stream_));
    SB_CUDA(cudaStreamSynchronize(stream_));
  } catch (...) {
    if (t->d) cudaFree(t->d);
    delete t;
    throw;
  }
  return t;
}

void Engine::free_text(DeviceText* t) {
  if (!t) return;
  cudaSetDevice(device_);
  if (t->d && t->owned) cudaFree(t->d);
  delete t;
}

DeviceText* Engine::stage_text(const uint8_t* host, uint64_t n) {
  SB_CUDA(cudaSetDevice(device_));
  if (n >= (1ull << kPosBits)) throw CapacityError("text longer than 2^40 bytes is not supported");
  const size_t need = padded_alloc(n);
  if (need > staged_.alloc) {
    if (staged_.d) cudaFree(staged_.d);
    staged_.d = nullptr;
    staged_.alloc = 0;
    SB_CUDA(cudaMalloc((void**)&staged_.d, need));
    staged_.alloc = need;
  }
  staged_.n = n;
  send_text(staged_.d, staged_.alloc, host, n);
  return &staged_;
}

// Text transfer.  Large Dna texts go over PCIe at 2 bits per character (transport.cu): the
// pool packs chunk by chunk into pinned memory, every finished chunk is handed to the copy
// engine at once, and a streaming kernel expands the text in HBM.  Everything else, and any
// text holding a byte outside ACGTacgt, is copied as bytes.  The tail padding is zeroed.
void Engine::send_text(uint8_t* dst, size_t dst_alloc, const uint8_t* host, uint64_t n) {
  SB_CUDA(cudaEventRecord(ev_[5], stream_));
  transfer_pending_ = true;
  transfer_packed_ = false;
  const size_t pad = std::min<size_t>(dst_alloc - n, 2ull * kMaxRowBytes + 256);
  bool sent = false;
  if (transport_mode_ == 1 && profile_ == kDna && n >= (8ull << 20)) {
    if (!pool_) {
      // one process per GPU (torchrun): the host cores are shared by LOCAL_WORLD_SIZE packing pools
      int nt = (int)std::thread::hardware_concurrency();
      const char* e = getenv("SASSY_B200_PACK_THREADS");
      const char* lw = getenv("LOCAL_WORLD_SIZE");
      if (e)
        nt = atoi(e);
      else if (lw && atoi(lw) > 1)
        nt = std::max(2, nt / atoi(lw));
      nt = std::max(1, std::min(nt, 64));
      pool_ = new PackPool(nt);
    }
    // Chunks of 2 Mi characters (512 KiB packed).  The workers pack from the front of the text into
    // a staging ring that stays in the host's caches; this thread hands every packed chunk to the
    // copy engine, and -- when the source is pinned and the next chunk is not ready while the copy
    // queue runs dry -- takes a chunk from the BACK of the text and sends it as plain bytes, so the
    // split between packed and plain bytes follows the speed of the host cores by itself.
    // SASSY_B200_PACK_CHUNK (characters) / SASSY_B200_PACK_RING (slots, 0 = one staging buffer
    // for the whole text, streaming stores) are tuning knobs.
    static const size_t chunk = [] {
      const char* e = getenv("SASSY_B200_PACK_CHUNK");
      size_t c = e ? (size_t)strtoull(e, nullptr, 10) : (size_t)2 << 20;
      return std::max<size_t>(c / 4096 * 4096, 65536);
    }();
    static const size_t merge = [] {  // packed chunks per copy, at most
      const char* e = getenv("SASSY_B200_PACK_MERGE");
      return std::max<size_t>(1, e ? (size_t)strtoull(e, nullptr, 10) : (size_t)8);
    }();
    static const size_t ring = [] {
      const char* e = getenv("SASSY_B200_PACK_RING");
      return e ? (size_t)strtoull(e, nullptr, 10) : (size_t)48;
    }();
    cudaPointerAttributes at;
    const bool pinned = cudaPointerGetAttributes(&at, host) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError();  // unregistered memory is not an error here
    const size_t nchunks = (size_t)((n + chunk - 1) / chunk);
    const size_t slots = ring ? std::min(ring, nchunks) : nchunks;
    const size_t stage_bytes = slots * (chunk / 4);
    if (stage_bytes > h_pack_cap_) {
      if (h_pack_) cudaFreeHost(h_pack_);
      h_pack_ = nullptr;
      h_pack_cap_ = 0;
      SB_CUDA(cudaHostAlloc((void**)&h_pack_, stage_bytes, cudaHostAllocDefault));
      h_pack_cap_ = stage_bytes;
    }
    d_pack_.ensure(nchunks * (chunk / 4));
    // The expansion runs on a second stream, group by group behind the copies, so that only the
    // last group's expansion is left when the last byte has crossed PCIe.
    if (!unpack_stream_) SB_CUDA(cudaStreamCreateWithFlags(&unpack_stream_, cudaStreamNonBlocking));
    constexpr size_t kCopyEvents = 64;  // copies in flight that are tracked
    while (copy_ev_.size() < kCopyEvents + 2) {
      cudaEvent_t e;
      SB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      copy_ev_.push_back(e);
    }
    const size_t group = std::max<size_t>(1, ((size_t)64 << 20) / chunk);  // chunks per expansion launch
    uint64_t issued = 0, completed = 0;     // copies (packed or plain), in stream order
    uint64_t bytes_issued = 0, bytes_completed = 0;
    uint64_t copy_len[kCopyEvents];
    // plain chunks are added while less than this is queued (~110 us of PCIe work: longer than a
    // sleep of this thread, so the copy engine does not run dry)
    const uint64_t queue_low = 6ull << 20;
    std::vector<uint64_t> seq(slots ? slots : 1, 0);  // copy number of the chunk that last used a staging slot
    size_t sent_chunks = 0, released = 0, expanded_chunks = 0;
    uint64_t plain_bytes = 0;
    bool clean = true, expanded = false;
    auto poll = [&]() {
      while (completed < issued) {
        const cudaError_t q = cudaEventQuery(copy_ev_[completed % kCopyEvents]);
        if (q == cudaErrorNotReady) {
          cudaGetLastError();  // "not ready" is an answer, not an error to be found by a later check
          break;
        }
        SB_CUDA(q);
        bytes_completed += copy_len[completed++ % kCopyEvents];
      }
      while (released < sent_chunks && seq[released % slots] < completed) released++;
      if (ring) pool_->release(released);
    };
    auto issue = [&](void* to, const void* from, size_t len) {
      if (issued - completed >= kCopyEvents) {  // the oldest tracked copy frees its event
        SB_CUDA(cudaEventSynchronize(copy_ev_[completed % kCopyEvents]));
        poll();
      }
      SB_CUDA(cudaMemcpyAsync(to, from, len, cudaMemcpyHostToDevice, stream_));
      SB_CUDA(cudaEventRecord(copy_ev_[issued % kCopyEvents], stream_));
      copy_len[issued % kCopyEvents] = len;
      bytes_issued += len;
      issued++;
    };
    auto expand_upto = [&](size_t upto_chunk) {  // chunks [expanded_chunks, upto_chunk)
      if (upto_chunk <= expanded_chunks) return;
      const uint64_t char0 = (uint64_t)expanded_chunks * chunk;
      const uint64_t chars = std::min<uint64_t>((uint64_t)upto_chunk * chunk, n) - char0;
      cudaEvent_t e = copy_ev_[kCopyEvents];
      SB_CUDA(cudaEventRecord(e, stream_));
      SB_CUDA(cudaStreamWaitEvent(unpack_stream_, e, 0));
      SB_CUDA(launch_unpack_dna(d_pack_.as<uint8_t>() + char0 / 4, dst + char0, chars, unpack_stream_));
      expanded_chunks = upto_chunk;
      expanded = true;
    };
#if defined(__linux__)
    // short sleeps of this thread must be short: the default timer slack adds 50 us to each
    const int slack = prctl(PR_GET_TIMERSLACK);
    prctl(PR_SET_TIMERSLACK, 1000UL);
#endif
    pool_->start(host, h_pack_, n, chunk, ring ? slots : 0);
    try {
    for (;;) {
      poll();
      const size_t packed_end = pool_->packed_end();
      if (sent_chunks >= packed_end) break;
      const int st = pool_->chunk_state(sent_chunks);
      if (st != 0) {
        if (st != 1) {
          clean = false;
          break;
        }
        // consecutive packed chunks that sit back to back in the staging buffer leave in ONE copy
        // (a copy of 512 KiB pays ~4 us of set-up on top of 9.5 us of transfer)
        const size_t c = sent_chunks;
        size_t cnt = 1;
        while (cnt < merge && c + cnt < packed_end && (c + cnt) % slots != 0 && (c + cnt) % group != 0 &&
               pool_->chunk_state(c + cnt) == 1)
          cnt++;
        const uint64_t chars = std::min<uint64_t>((uint64_t)cnt * chunk, n - (uint64_t)c * chunk);
        const size_t len = (size_t)((chars + 63) / 64 * 16);
        issue(d_pack_.as<uint8_t>() + c * (chunk / 4), h_pack_ + (c % slots) * (chunk / 4), len);
        for (size_t j = 0; j < cnt; j++) seq[(c + j) % slots] = issued - 1;
        sent_chunks += cnt;
        if (sent_chunks % group == 0) expand_upto(sent_chunks);
        continue;
      }
      size_t t;
      if (pinned && bytes_issued - bytes_completed < queue_low && pool_->claim_tail(&t)) {
        const uint64_t off = (uint64_t)t * chunk;
        const uint64_t len = std::min<uint64_t>(chunk, n - off);
        issue(dst + off, host + off, (size_t)len);
        plain_bytes += len;
        continue;
      }
      std::this_thread::sleep_for(std::chrono::microseconds(20));
    }
    } catch (...) {  // a failed CUDA call: the workers must be idle before the pool is used again
      pool_->cancel();
      pool_->finish();
      throw;
    }
    if (!clean) pool_->cancel();
    pool_->finish();
#if defined(__linux__)
    if (slack > 0) prctl(PR_SET_TIMERSLACK, (unsigned long)slack);
#endif
    if (clean) expand_upto(sent_chunks);
    if (expanded) {  // everything later in stream_ (also the byte copies of a fallback) follows the expansion
      cudaEvent_t e = copy_ev_[kCopyEvents + 1];
      SB_CUDA(cudaEventRecord(e, unpack_stream_));
      SB_CUDA(cudaStreamWaitEvent(stream_, e, 0));
    }
    if (clean) {
      // the expansion wrote whole groups of 64: clear what lies beyond the text
      SB_CUDA(cudaMemsetAsync(dst + n, 0, pad, stream_));
      const uint64_t packed_chars = std::min<uint64_t>((uint64_t)sent_chunks * chunk, n);
      transfer_packed_ = sent_chunks > 0;
      transfer_bytes_ = (packed_chars + 63) / 64 * 16 + plain_bytes;
      sent = true;
    }
  }
  if (!sent) {
    transfer_bytes_ = n;
    // Copy in slices so that pinned sources stream at full PCIe rate.
    const uint64_t slice = 256ull << 20;
    for (uint64_t off = 0; off < n; off += slice) {
      const uint64_t len = std::min(slice, n - off);
      SB_CUDA(cudaMemcpyAsync(dst + off, host + off, len, cudaMemcpyHostToDevice, stream_));
    }
    SB_CUDA(cudaMemsetAsync(dst + n, 0, pad, stream_));
  }
  SB_CUDA(cudaEventRecord(ev_[6], stream_));
}

// All small per-search inputs travel in ONE host->device copy from a pinned staging
// buffer: [4 counters][equality tables][query bytes][direction flags][prefilter tables].
void Engine::upload_params(const std::vector<Query>& queries, int m, int W, const FilterPlan& fp, bool pair,
                           bool fused, const QgramPlan* qp) {
  const size_t nq = queries.size();
  const size_t eq_bytes = nq * nrows_ * W * sizeof(uint32_t);
  const size_t pat_bytes = nq * (size_t)m;
  const int WT = fused ? 2 * fp.WF : fp.WF;  // automaton words per filter table
  // q-gram route: ONE bitmap for all queries (pattern and reversed partner), then the confirm codes
  const size_t ntab = qp ? 1 : (fused ? nq / 2 : nq);
  const size_t tab_words = qp ? qp->table_words()
                              : (!fp.enabled ? 0 : (pair ? (size_t)kPairTableWords * WT : (size_t)256 * WT));
  // piece records of the exact hit refinement (Dna): [query][piece][kConfWords], behind the tables
  const int conf_pieces = profile_ != kDna ? 0 : (qp ? qp->npieces : (fp.enabled ? fp.npieces : 0));
  const size_t conf_words = nq * (size_t)conf_pieces * kConfWords;
  auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
  off_counts_ = 0;
  off_eq_ = align(8 * sizeof(unsigned long long));
  off_pat_ = off_eq_ + align(eq_bytes);
  off_rev_ = off_pat_ + align(pat_bytes);
  off_feq_ = off_rev_ + align(nq);
  off_qconf_ = off_feq_ + align(ntab * tab_words * sizeof(uint32_t));
  const size_t total = off_qconf_ + align(conf_words * sizeof(uint32_t));
  if (total > stage_cap_) {
    if (h_stage_) cudaFreeHost(h_stage_);
    h_stage_ = nullptr;
    stage_cap_ = 0;
    SB_CUDA(cudaHostAlloc((void**)&h_stage_, total * 2, cudaHostAllocDefault));
    stage_cap_ = total * 2;
  }
  d_stage_.ensure(stage_cap_);
  memset(h_stage_ + off_counts_, 0, 8 * sizeof(unsigned long long));
  uint32_t* h_eq = reinterpret_cast<uint32_t*>(h_stage_ + off_eq_);
  uint32_t* h_feq = reinterpret_cast<uint32_t*>(h_stage_ + off_feq_);
  for (size_t q = 0; q < nq; q++) {
    memcpy(h_stage_ + off_pat_ + q * m, queries[q].bytes, m);
    h_stage_[off_rev_ + q] = queries[q].rev ? 1 : 0;
    build_eq_table(profile_, queries[q].bytes, m, W, nrows_, h_eq + q * nrows_ * W);
    uint32_t* h_conf = reinterpret_cast<uint32_t*>(h_stage_ + off_qconf_) + q * (size_t)conf_pieces * kConfWords;
    if (qp) {
      if (q == 0) memset(h_feq, 0, tab_words * sizeof(uint32_t));
      add_qgram_entries(*qp, queries[q].bytes, queries[q].rev, h_feq);
      if (conf_pieces) build_qgram_confirm(*qp, queries[q].bytes, queries[q].rev, h_conf);
    } else if (fp.enabled && conf_pieces) {
      // reversed queries: matched back to front by the forward pass when the strands share it
      // (fused), by their own right-to-left pass otherwise
      build_filter_confirm(fp, queries[q].bytes, queries[q].rev, queries[q].rev && !fused, h_conf);
    }
    if (!qp && fp.enabled && q < ntab) {
      const uint8_t* partner = fused ? queries[q + ntab].bytes : nullptr;  // the reversed partner query
      if (pair)
        build_pair_table(fp, queries[q].bytes, h_feq + q * tab_words, partner);
      else
        build_filter_table(profile_, fp, queries[q].bytes, h_feq + q * tab_words, partner);
    }
  }
  conf_pieces_ = conf_pieces;
  SB_CUDA(cudaMemcpyAsync(d_stage_.p, h_stage_, total, cudaMemcpyHostToDevice, stream_));
}

void Engine::make_tensor_map(CUtensorMap* map, const DeviceText& text, const ScanGeom& g) const {
  auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(encode_tiled_);
  const cuuint64_t dims[2] = {g.ltot, std::max<uint32_t>(g.rows, 1)};
  const cuuint64_t strides[1] = {g.ltot};
  const cuuint32_t box[2] = {kStageBytes, 32};  // one box per warp per stage
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, text.d, dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE,
                      kStageBytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[128];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    throw CudaError(buf);
  }
}

void Engine::overhang_args(const SearchOpts& opts, int m, int k, int W, OverhangArgs& o) const {
  memset(&o, 0, sizeof o);
  const int pad = 32 * W - m;
  for (int j = 0; j < m; j++)
    if (overhang_left_cost(j + 1, opts.alpha, opts.max_overhang) - overhang_left_cost(j, opts.alpha, opts.max_overhang)) {
      const int b = pad + j;
      o.init_pv[b >> 5] |= 1u << (b & 31);
    }
  o.left_total = overhang_left_cost(m, opts.alpha, opts.max_overhang);
  // get_overhang_steps (src/search.rs:347-356): min(m, ceil((k + alpha) / alpha), max_overhang);
  // Rust's `as usize` turns NaN (k = 0, alpha = 0) into 0 and saturates +inf
  const float r = ceilf(((float)k + opts.alpha) / opts.alpha);
  uint64_t steps = std::isnan(r) ? 0 : (r >= 1e18f ? ~0ull : (uint64_t)r);
  steps = std::min<uint64_t>(steps, (uint64_t)m);
  if (opts.max_overhang >= 0) steps = std::min<uint64_t>(steps, (uint64_t)opts.max_overhang);
  o.steps = (uint32_t)steps;
  o.alpha = opts.alpha;
}

// Candidates (unsorted, in keys_/cost_) -> device radix sort -> selection (first copy of a
// position, local-minima rule, end filters, only_best_match) -> stream compaction -> traceback
// (or end position + cost only) -> host.  Returns the number of records written to `out`.
uint64_t Engine::post_process(const PostCtx& c, const SearchOpts& opts, uint64_t ncand, MatchSet& out) {
  const int m = c.m, k = c.k, W = c.W;
  unsigned long long* d_nsel = c.d_counts + 1;
  unsigned long long h_counts[4] = {0, 0, 0, 0};
  auto read_counts = [&]() {
    SB_CUDA(cudaMemcpyAsync(h_counts, c.d_counts, sizeof h_counts, cudaMemcpyDeviceToHost, stream_));
    SB_CUDA(cudaStreamSynchronize(stream_));
  };
  keys2_.ensure(ncand * sizeof(uint64_t));
  cost2_.ensure(ncand * sizeof(uint32_t));
  size_t tmp_bytes = 0;
  SB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_.as<uint64_t>(), keys2_.as<uint64_t>(),
                                          cost_.as<uint32_t>(), cost2_.as<uint32_t>(), ncand, 0, c.end_bit,
                                          stream_));
  cubtmp_.ensure(tmp_bytes);
  SB_CUDA(cub::DeviceRadixSort::SortPairs(cubtmp_.p, tmp_bytes, keys_.as<uint64_t>(), keys2_.as<uint64_t>(),
                                          cost_.as<uint32_t>(), cost2_.as<uint32_t>(), ncand, 0, c.end_bit,
                                          stream_));
  const uint64_t* skeys = keys2_.as<uint64_t>();
  const uint32_t* scost = cost2_.as<uint32_t>();
  const uint64_t* sel_keys = skeys;
  const uint32_t* sel_cost = scost;
  // overlapping re-scan windows of the prefilter report an end position more than once; the
  // selection kernel keeps the first copy only
  const bool end_filter = opts.pam_len > 0 || (opts.n_endpoint && opts.max_n_frac >= 0.f);
  const bool need_select = !opts.all_minima || c.dedup || end_filter || opts.only_best;
  if (need_select) {
    flags_.ensure(ncand);
    sel_.ensure(ncand * sizeof(uint64_t));
    EndFilter ef;
    memset(&ef, 0, sizeof ef);
    if (end_filter) {
      if (opts.pam_len > kMaxPam) throw CudaError("PAM longer than 16 characters is not supported");
      ef.text = c.text;
      ef.rev_flags = c.d_rev;
      ef.profile = profile_;
      ef.m = m, ef.k = k;
      ef.pam_len = opts.pam_len;
      for (int i = 0; i < opts.pam_len; i++) {
        ef.pam[0][i] = opts.pam[i];
        ef.pam[1][i] = complement_byte(profile_, opts.pam[i]);
      }
      ef.n_endpoint = (opts.n_endpoint && opts.max_n_frac >= 0.f) ? 1 : 0;
      ef.max_n_frac = opts.max_n_frac;
    }
    SB_CUDA(launch_minima(skeys, scost, ncand, flags_.as<uint8_t>(), opts.all_minima, end_filter ? &ef : nullptr,
                          stream_));
    stats_.aux_launches++;
    if (opts.only_best) {
      best_.ensure((size_t)c.nslots * sizeof(unsigned long long));
      SB_CUDA(launch_best(skeys, scost, ncand, flags_.as<uint8_t>(), best_.as<unsigned long long>(), c.nslots,
                          stream_));
      stats_.aux_launches += 2;
    }
    size_t tmp2 = 0;
    SB_CUDA(cub::DeviceSelect::Flagged(nullptr, tmp2, skeys, flags_.as<uint8_t>(), sel_.as<uint64_t>(), d_nsel,
                                       ncand, stream_));
    cubtmp_.ensure(tmp2);
    SB_CUDA(cub::DeviceSelect::Flagged(cubtmp_.p, tmp2, skeys, flags_.as<uint8_t>(), sel_.as<uint64_t>(), d_nsel,
                                       ncand, stream_));
    sel_keys = sel_.as<uint64_t>();
    if (opts.without_trace) {
      sel_cost_.ensure(ncand * sizeof(uint32_t));
      SB_CUDA(cub::DeviceSelect::Flagged(nullptr, tmp2, scost, flags_.as<uint8_t>(), sel_cost_.as<uint32_t>(),
                                         d_nsel, ncand, stream_));
      cubtmp_.ensure(tmp2);
      SB_CUDA(cub::DeviceSelect::Flagged(cubtmp_.p, tmp2, scost, flags_.as<uint8_t>(), sel_cost_.as<uint32_t>(),
                                         d_nsel, ncand, stream_));
      sel_cost = sel_cost_.as<uint32_t>();
    }
  }
  // Small candidate lists (the normal case): trace every possible selection slot bounded by the
  // device-side count and fetch results + count with one synchronisation.  Large lists: read
  // the count first so that buffers and copies have the exact size.
  const bool fast = ncand <= 65536;
  uint64_t bound = ncand;
  if (need_select && !fast) {
    read_counts();
    bound = h_counts[1];
  }
  uint64_t nsel = 0;
  if (bound > 0) {
    const uint64_t words_per_match = opts.without_trace ? 0 : trace_words_per_match(m, k, W);
    const uint64_t max_scratch_words = ((W > kMaxScanWords ? 4096ull : 512ull) << 20) / 4;  // scratch per slice
    uint64_t slice = bound;
    const bool ovt = opts.alpha >= 0.f;
    while (slice > 1 && trace_slots(slice, W, opts.without_trace, ovt) * words_per_match > max_scratch_words)
      slice = (slice + 1) / 2;
    scratch_.ensure(trace_slots(slice, W, opts.without_trace, ovt) * words_per_match * sizeof(uint32_t));
    out_.ensure(bound * sizeof(GpuMatch));
    const uint32_t ops_words = opts.without_trace ? 0 : out.ops_words;
    ops_.ensure(bound * ops_words * sizeof(uint32_t));
    TraceArgs t;
    memset(&t, 0, sizeof t);
    t.text = c.text;
    t.profile = profile_;
    t.patterns = c.d_pat;
    t.rev_flags = c.d_rev;
    t.eq = c.d_eq;
    t.nrows = nrows_;
    t.sh0 = sh0_;
    t.msk0 = msk0_;
    t.m = m;
    t.k = k;
    t.W = W;
    t.keys = sel_keys;
    t.costs = opts.without_trace ? sel_cost : nullptr;
    t.max_n_frac = opts.without_trace ? -1.f : opts.max_n_frac;
    t.alpha = opts.alpha;
    t.max_overhang = opts.max_overhang;
    t.count_dev = (need_select && fast) ? d_nsel : nullptr;
    t.scratch = scratch_.as<uint32_t>();
    t.ops = ops_.as<uint32_t>();
    t.ops_words = ops_words;
    t.out = out_.as<GpuMatch>();
    for (uint64_t first = 0; first < bound; first += slice) {
      t.first = first;
      t.count = std::min(slice, bound - first);
      SB_CUDA(launch_trace(t, stream_));
      stats_.aux_launches++;
    }
    out.m.resize(bound);
    out.ops.resize(bound * out.ops_words);
    SB_CUDA(cudaMemcpyAsync(out.m.data(), out_.p, bound * sizeof(GpuMatch), cudaMemcpyDeviceToHost, stream_));
    if (ops_words)
      SB_CUDA(cudaMemcpyAsync(out.ops.data(), ops_.p, out.ops.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                              stream_));
  }
  if (need_select && fast) {
    read_counts();
    nsel = h_counts[1];
    out.m.resize(nsel);
    out.ops.resize(nsel * out.ops_words);
  } else {
    SB_CUDA(cudaStreamSynchronize(stream_));
    nsel = bound;
  }
  // traced N-fraction filter (src/search.rs:924-934, general.rs:399-402): the kernel flagged
  // the records, drop them here (keeps order)
  if (!opts.without_trace && opts.max_n_frac >= 0.f) {
    uint64_t w = 0;
    for (uint64_t i = 0; i < nsel; i++) {
      if (out.m[i].failed & 2u) continue;
      if (w != i) {
        out.m[w] = out.m[i];
        std::copy(out.ops.begin() + i * out.ops_words, out.ops.begin() + (i + 1) * out.ops_words,
                  out.ops.begin() + w * out.ops_words);
      }
      w++;
    }
    nsel = w;
    out.m.resize(nsel);
    out.ops.resize(nsel * out.ops_words);
  }
  return nsel;
}

void Engine::search(const DeviceText& text, const std::vector<Query>& queries, int m, int k, const SearchOpts& opts,
                    MatchSet& out) {
  const bool all_minima = opts.all_minima, include_pos0 = opts.include_pos0;
  const bool timing = HostTimers::on();
  const double t_entry = timing ? HostTimers::now_us() : 0.0;
  SB_CUDA(cudaSetDevice(device_));
  cudaGetLastError();  // a stale error left by another library in this thread is not ours to report
  stats_ = SearchStats();
  out.m.clear();
  out.ops.clear();
  const uint32_t nq = (uint32_t)queries.size();
  if (m <= 0) throw CudaError("empty pattern");
  const int W = round_words((m + 31) / 32);
  if (W < 0) throw CapacityError("pattern longer than 4096 characters is not supported");
  // more than 32 words: no row-tiled kernels (prefilter hits and, without a prefilter, windows that
  // cover the text are re-scanned one warp per window), no overhang
  const bool huge = W > kMaxScanWords;
  if (huge && opts.alpha >= 0.f)
    throw CapacityError("overhang is not supported for patterns longer than 1024 characters");
  if (nq == 0) {
    if (pg_) throw CudaError("a gathered search needs at least one query on every rank");
    return;
  }
  if (nq >= (1u << (64 - kPosBits))) throw CudaError("too many queries in one search");
  if (k < 0) k = 0;
  uint32_t nfwd = 0;
  while (nfwd < nq && !queries[nfwd].rev) nfwd++;
  for (uint32_t q = nfwd; q < nq; q++)
    if (!queries[q].rev) throw CudaError("forward queries must precede reversed ones");
  out.ops_words = (uint32_t)((m + k + 1 + 15) / 16);
  stats_.words = (uint32_t)W;

  const uint64_t n = text.n;
  SB_CUDA(cudaEventRecord(ev_[0], stream_));

  // ---- plan: exact piece prefilter or full scan ------------------------------------------------
  // Overhang: the pigeonhole argument of the prefilter does not hold for alignments that hang
  // over a text end, so the full scan runs; the edge kernel owns the end positions it changes.
  const bool ov = opts.alpha >= 0.f;
  if (ov && opts.pam_len > 0) throw CudaError("an end filter cannot be combined with overhang");
  FilterPlan fp;
  if (n > 0 && filter_mode_ != 0 && !ov) {
    std::vector<const uint8_t*> qptr(nq);
    for (uint32_t q = 0; q < nq; q++) qptr[q] = queries[q].bytes;
    fp = plan_filter(profile_, qptr.data(), nq, m, k, filter_mode_ == 2 ? 1e30 : 0.85);
  }
  // Dna: two characters per step through the class-pair table; others: byte-indexed table
  // (the pair table of a 4-word automaton has 512 distinct bytes per warp access: shared-memory
  //  bandwidth, not instructions, would bound it -- measured 2.4x slower than the byte table)
  // Both strands of a v1 search (forward queries followed by their reversed partners) share ONE
  // forward pass: the partner's pieces, matched back to front, occupy a second set of words.
  // Dna, one pattern (with or without its reversed partner): the q-gram bitmap finds the shares of
  // both strands in one forward pass at a cost per character that does not depend on m or k
  // (scan_core.cuh); it needs shares of >= 9 characters, shorter ones keep the piece automaton.
  QgramPlan qp;
  if (n > 0 && filter_mode_ != 0 && qgram_mode_ != 0 && !ov && profile_ == kDna && nfwd == 1 && nq <= 2)
    qp = plan_qgram(m, k, (int)nq, qgram_min_q_);
  const bool qgram = qp.enabled;
  if (qgram) fp.enabled = false;
  const bool fused = fp.enabled && fuse_strands_ && nfwd > 0 && nq == 2 * nfwd && fp.WF <= 2;
  const int WT = fused ? 2 * fp.WF : fp.WF;
  const bool pair = profile_ == kDna && WT <= pair_max_words_;
  const size_t tab_words = pair ? (size_t)kPairTableWords * WT : (size_t)256 * WT;
  upload_params(queries, m, W, fp, pair, fused, qgram ? &qp : nullptr);
  const double t_uploaded = timing ? HostTimers::now_us() : 0.0;
  if (timing) HostTimers::add(HostTimers::kPre, t_uploaded - t_entry);
  double t_last_sync = t_uploaded;
  uint8_t* dst = d_stage_.as<uint8_t>();
  unsigned long long* d_counts = reinterpret_cast<unsigned long long*>(dst + off_counts_);
  unsigned long long* d_cand_count = d_counts;      // [0] candidates
  unsigned long long* d_nsel = d_counts + 1;         // [1] selected candidates
  unsigned long long* d_hit_count = d_counts + 2;    // [2] prefilter hits
  const uint32_t* d_eq = reinterpret_cast<const uint32_t*>(dst + off_eq_);
  const uint8_t* d_pat = dst + off_pat_;
  const uint8_t* d_rev = dst + off_rev_;
  const uint32_t* d_feq = reinterpret_cast<const uint32_t*>(dst + off_feq_);

  const int occ = scan_blocks_per_sm(std::min(W, kMaxScanWords), false, variant_, nrows_);
  stats_.blocks_per_sm = (uint32_t)occ;
  // blocks per tile of one scan launch: the queries of a direction (pairs of them in scan2_kernel)
  auto launch_cols = [&](uint32_t cnt) { return (W == 1 && cnt >= 2 && scan2_) ? (cnt + 1) / 2 : cnt; };
  const uint32_t cols = std::max<uint32_t>(1, std::max(launch_cols(nfwd), launch_cols(nq - nfwd)));
  const ScanGeom g = choose_geom(n, m, k, cols, occ * sm_count_);
  if ((uint64_t)g.rows * g.ltot > text.alloc) throw CudaError("internal: text padding too small for tiling");
  stats_.ltot = g.ltot;
  stats_.rows = g.rows;

  if (cand_cap_ == 0) cand_cap_ = 1ull << 20;

  ScanArgs a;
  memset(&a, 0, sizeof a);
  a.text = text.d;
  a.n = n;
  a.g = g;
  a.sh0 = sh0_;
  a.msk0 = msk0_;
  a.nrows = nrows_;
  a.rowbytes = (uint32_t)W * 4u;
  a.m = m;
  a.k = k;
  a.cand_count = d_cand_count;
  OverhangArgs oa;
  if (ov) {
    a.emit_min = std::min<uint64_t>(n, (uint64_t)m + (uint64_t)k);
    overhang_args(opts, m, k, W, oa);
    oa.text = TextRef{text.d, n, nullptr, nullptr, nq};
    oa.rev_flags = d_rev;
    oa.nslots = nq;
  }

  // end position 0 (empty text prefix) has cost m: a candidate iff m <= k.
  // (reference src/search.rs:1320-1322; never reported for an empty text, :1314-1316)
  std::vector<uint64_t> k0;
  std::vector<uint32_t> c0;
  if (include_pos0 && m <= k && n > 0 && !ov)
    for (uint32_t q = 0; q < nq; q++) {
      k0.push_back(cand_key(q, 0));
      c0.push_back((uint32_t)m);
    }
  bool counts_fresh = true;  // the staged upload zeroed the counters
  auto reset_candidates = [&]() {
    keys_.ensure(cand_cap_ * sizeof(uint64_t));
    cost_.ensure(cand_cap_ * sizeof(uint32_t));
    if (!k0.empty()) {
      SB_CUDA(cudaMemcpyAsync(keys_.p, k0.data(), k0.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, stream_));
      SB_CUDA(cudaMemcpyAsync(cost_.p, c0.data(), c0.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, stream_));
    }
    if (!counts_fresh || !k0.empty()) {
      const unsigned long long init = k0.size();
      SB_CUDA(cudaMemcpyAsync(d_cand_count, &init, sizeof init, cudaMemcpyHostToDevice, stream_));
    }
    counts_fresh = false;
    a.cand_keys = keys_.as<uint64_t>();
    a.cand_cost = cost_.as<uint32_t>();
    a.cand_cap = cand_cap_;
  };
  unsigned long long h_counts[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // [4] = q-gram hits that passed the confirmation
  auto read_counts = [&]() {
    SB_CUDA(cudaMemcpyAsync(h_counts, d_counts, sizeof h_counts, cudaMemcpyDeviceToHost, stream_));
    SB_CUDA(cudaStreamSynchronize(stream_));
    if (timing) t_last_sync = HostTimers::now_us();
  };
  auto elapsed = [&](cudaEvent_t e0, cudaEvent_t e1) {
    float ms = 0;
    SB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    return ms;
  };

  // Fast tail for small candidate lists (the normal case): one block sorts + selects, the
  // traceback runs on the device-side selection count and writes straight into pinned host
  // memory.  Queued right behind the candidate producers, so a search needs ONE host
  // synchronisation; longer lists set the `big` flag and take the general path below.
  int end_bit = kPosBits;
  while ((1ull << (end_bit - kPosBits)) < nq) end_bit++;
  const uint64_t words_per_match = trace_words_per_match(m, k, W);
  const bool small_path = words_per_match <= 4096 && !opts.special();
  unsigned long long* d_big = d_counts + 3;
  if (small_path) {
    const size_t need_out = (size_t)kSmallCandidates * sizeof(GpuMatch);
    const size_t need_ops = (size_t)kSmallCandidates * out.ops_words * sizeof(uint32_t);
    if (need_out + need_ops > h_small_cap_) {
      if (h_small_) cudaFreeHost(h_small_);
      h_small_ = nullptr;
      h_small_cap_ = 0;
      SB_CUDA(cudaHostAlloc((void**)&h_small_, 2 * (need_out + need_ops), cudaHostAllocDefault));
      h_small_cap_ = 2 * (need_out + need_ops);
    }
    sel_small_.ensure((size_t)kSmallCandidates * sizeof(uint64_t));
    scratch_.ensure(trace_threads(kSmallCandidates) * words_per_match * sizeof(uint32_t));
  }
  GpuMatch* h_small_out = reinterpret_cast<GpuMatch*>(h_small_);
  uint32_t* h_small_ops = reinterpret_cast<uint32_t*>(h_small_ + (size_t)kSmallCandidates * sizeof(GpuMatch));
  // Multi-GPU gather over peer memory (peer_gather.cu): the traceback leaves the records in this
  // rank's slot of the receive buffer, push + collect follow in the same stream.  Exactly one
  // exchange per search keeps the ranks in lock step; a search whose result is not complete in
  // the slot (long candidate list, overflow, re-scan) marks its slot `overflow` and every rank
  // falls back to the caller's collective.
  PeerGather* pg = pg_;
  gather_ok_ = false;
  if (pg && pg->broken()) throw CudaError("this peer gather timed out earlier and cannot be used again");
  const bool pg_slot = pg && small_path && out.ops_words <= pg->max_ops_words() && pg->cap() >= (size_t)kSmallCandidates;
  bool pg_pushed = false, small_in_slot = false;
  unsigned long long pg_hit_limit = ~0ull;  // prefilter routes: more hits than this -> the slot is marked incomplete
  auto pg_exchange = [&](bool force_overflow) {
    if (!pg || pg_pushed) return;
    pg_pushed = true;
    SB_CUDA(pg->exchange(d_counts, a.cand_cap, pg_hit_limit, out.ops_words, force_overflow, n,
                         pg_user_, stream_));
    stats_.aux_launches += 2;
  };
  bool tail_event = false;
  auto queue_small_tail = [&]() {
    if (!small_path) {
      pg_exchange(true);
      return;
    }
    SB_CUDA(launch_post_small(a.cand_keys, a.cand_cost, d_cand_count, a.cand_cap, sel_small_.as<uint64_t>(), d_nsel,
                              d_big, all_minima, end_bit, stream_));
    TraceArgs t;
    memset(&t, 0, sizeof t);
    t.text = TextRef{text.d, n, nullptr, nullptr, nq};
    t.max_n_frac = -1.f;
    t.alpha = -1.f;
    t.profile = profile_;
    t.patterns = d_pat;
    t.rev_flags = d_rev;
    t.eq = d_eq;
    t.nrows = nrows_;
    t.sh0 = sh0_;
    t.msk0 = msk0_;
    t.m = m;
    t.k = k;
    t.W = W;
    t.keys = sel_small_.as<uint64_t>();
    t.first = 0;
    t.count = kSmallCandidates;
    t.count_dev = d_nsel;
    t.scratch = scratch_.as<uint32_t>();
    t.ops_words = out.ops_words;
    small_in_slot = pg_slot && !pg_pushed;
    if (small_in_slot) {
      size_t ops_off = 0;
      uint8_t* rec = pg->local_records(&ops_off);  // device memory: this rank's slot
      t.out = reinterpret_cast<GpuMatch*>(rec);
      t.ops = reinterpret_cast<uint32_t*>(rec + ops_off);
    } else {
      t.ops = h_small_ops;   // pinned host memory, written by the kernel over PCIe
      t.out = h_small_out;
    }
    SB_CUDA(launch_trace(t, stream_));
    stats_.aux_launches += 2;
    pg_exchange(!small_in_slot);
    // end of the search when this tail completes it: the caller synchronises right behind it
    SB_CUDA(cudaEventRecord(ev_[3], stream_));
    tail_event = true;
  };
  // after the synchronisation that follows a tail: did every rank deliver a complete result?
  auto pg_check = [&]() {
    if (pg && pg_pushed && pg->pipelined()) {
      // pipelined exchange: the host mirror holds the PREVIOUS step (Searcher reads it); whether
      // this step's own records are complete in the slot follows from the counters
      if (pg->has_result() && pg->timed_out()) {
        pg->mark_broken();
        throw CudaError("peer gather timed out (SASSY_B200_GATHER_TIMEOUT_S): a rank did not reach the previous search");
      }
      gather_ok_ = small_in_slot && h_counts[3] == 0 && h_counts[0] <= a.cand_cap &&
                   h_counts[2] <= pg_hit_limit && h_counts[1] <= pg->cap();
      return gather_ok_;
    }
    if (!pg || !pg_pushed || gather_ok_) return gather_ok_;
    if (pg->timed_out()) {
      pg->mark_broken();
      throw CudaError("peer gather timed out (SASSY_B200_GATHER_TIMEOUT_S): a rank did not reach this search; "
                      "the gather object cannot be used again");
    }
    gather_ok_ = pg->ok();
    return gather_ok_;
  };
  bool small_done = false;

  uint64_t ncand = 0;
  bool filtered = false;

  // ---- candidates, route 1: exact piece prefilter + re-scan of the hit neighbourhoods -----
  if (fp.enabled || qgram) {
    {  // room for 2x the expected number of hits (uniform text), within 8 M .. 128 M entries
      const double expect = 2.0 * (qgram ? qp.rate : fp.rate) * (double)n * nq;
      uint64_t want = (uint64_t)std::min(std::max(expect, 8.0 * 1048576.0), 128.0 * 1048576.0);
      if (want > hit_cap_) hit_cap_ = want;
    }
    hits_.ensure(hit_cap_ * sizeof(uint64_t));
    const int focc = qgram ? qgram_blocks_per_sm(qp.q, qp.s, variant_) : filter_blocks_per_sm(WT, variant_, pair);
    ScanGeom gf = choose_geom(n, m, k, qgram ? 1 : nq, focc * sm_count_);
    gf.nwarm = 1;  // a piece plus its delay line is at most 32 characters; a q-gram window 16
    // Short rows keep the rows of a warp close together in memory (a warp's TMA box gathers 64 bytes
    // of each of its 32 rows): 4 KB rows read 4-5 % faster than the 13-16 KB rows that fill whole
    // waves, at 1.6 % warm-up overhead (profiles/r02_tile_experiments.txt).  SASSY_B200_FILTER_ROW_BYTES
    // overrides (tiling experiments).
    const int row_bytes = filter_row_bytes_ >= 128 ? filter_row_bytes_ : (gf.ltot > 4096 ? 4096 : 0);
    if (row_bytes >= 128 && n > 0) {
      gf.ltot = std::min<uint32_t>(kMaxRowBytes, (uint32_t)row_bytes / kRowAlign * kRowAlign);
      gf.rows = (uint32_t)((n + gf.ltot - 1) / gf.ltot);
      gf.nstage = gf.ltot / kStageBytes;
    }
    const bool qseq = qgram && variant_ == kVariantTma && qgram_seq_ && n < (1ull << 36) && kStageBytes == 64;
    CUtensorMap ftmap;
    memset(&ftmap, 0, sizeof ftmap);
    if (variant_ == kVariantTma && !qseq) make_tensor_map(&ftmap, text, gf);
    reset_candidates();
    ScanArgs f = a;
    f.g = gf;
    for (int w = 0; w < kMaxFilterWords; w++) f.finit[w] = fp.finit[w], f.fdelay[w] = fp.fdelay[w];
    if (fused)
      for (int w = 0; w < fp.WF; w++) f.finit[fp.WF + w] = fp.finit[w], f.fdelay[fp.WF + w] = fp.fdelay[w];
    f.fused = fused ? 1 : 0;
    f.hit_keys = hits_.as<uint64_t>();
    f.hit_count = d_hit_count;
    f.hit_cap = hit_cap_;
    SB_CUDA(cudaEventRecord(ev_[1], stream_));
    if (qgram) {
      f.nq = 1;
      f.qs_base = 0;
      f.feq = d_feq;
      f.fused = nq == 2 ? 1 : 0;  // every hit is re-scanned for the reversed partner (slot 1) too
      if (qseq) {  // contiguous 2 KB tiles: the text as [ceil(n / 64)][64] bytes
        ScanGeom gs;
        gs.ltot = kStageBytes;
        gs.rows = (uint32_t)((n + 2047) / 2048 * 32);
        gs.nstage = 1, gs.nwarm = 0;
        make_tensor_map(&ftmap, text, gs);
        SB_CUDA(launch_qgram_seq(qp.q, qp.s, &ftmap, f, sm_count_, stream_));
        gf.ltot = 2048;
        gf.rows = gs.rows / 32;
      } else {
        SB_CUDA(launch_qgram(qp.q, qp.s, variant_, &ftmap, f, stream_));
      }
      stats_.scan_launches++;
    } else if (nfwd) {
      f.nq = nfwd;
      f.qs_base = 0;
      f.feq = d_feq;
      SB_CUDA(launch_filter(WT, false, variant_, pair, &ftmap, f, stream_));
      stats_.scan_launches++;
    }
    if (nq > nfwd && !fused && !qgram) {
      f.nq = nq - nfwd;
      f.qs_base = nfwd;
      f.feq = d_feq + (size_t)nfwd * tab_words;
      SB_CUDA(launch_filter(fp.WF, true, variant_, pair, &ftmap, f, stream_));
      stats_.scan_launches++;
    }
    SB_CUDA(cudaEventRecord(ev_[2], stream_));
    // the re-scan reads the hit count on the device: no host round trip between the two kernels
    ScanArgs v = a;
    v.nq = nq;
    v.qs_base = 0;
    v.eq = d_eq;
    v.hit_keys = hits_.as<uint64_t>();
    v.hit_count = d_hit_count;
    v.hit_cap = hit_cap_;
    // Regional fallback.  A tile = the kScanThreads rows of one block of the scan geometry.  When the
    // prefilter fires so often that re-scanning the hit neighbourhoods would cost a sizeable part of
    // a full scan (repeats, low-complexity sequence), a second pass marks the tiles whose hits are
    // not worth re-scanning as dense, drops their hits and runs the bit-parallel scan over exactly
    // those tiles: a satellite or a poly-A stretch then costs the scan of its own tiles, not of the
    // whole text.  The first pass carries a device-side guard (refine / verify return at once above
    // `heavy_hits`), so a pathological text does not pay for a useless re-scan first.
    const uint32_t ntiles = (g.rows + kScanThreads - 1) / kScanThreads;
    const uint64_t tile_bytes = (uint64_t)kScanThreads * g.ltot;
    // per-hit cost in scanned characters: an unrefined hit re-scans 2(m+k)+16 characters; a q-gram
    // hit costs its refinement plus, in a repeat, a re-scan of about that size
    const double hit_cost = qgram ? 32.0 + 0.25 * (2.0 * (m + k) + kHitChars) : 2.0 * (m + k) + kHitChars;
    const unsigned long long heavy_hits = (unsigned long long)std::max(1024.0, 0.25 * (double)n * nq / hit_cost);
    v.guard_count = d_hit_count;
    v.guard_limit = heavy_hits;
    pg_hit_limit = std::min<unsigned long long>(hit_cap_, heavy_hits);
    auto enable_regional = [&]() {
      const double max_hits = 0.5 * (double)tile_bytes * nq / hit_cost;  // re-scan <= half a tile scan
      tiles_.ensure((size_t)ntiles * 9 + 64);
      uint32_t* d_tile_counts = tiles_.as<uint32_t>();
      uint32_t* d_tile_list = d_tile_counts + ntiles;
      uint8_t* d_dense = reinterpret_cast<uint8_t*>(d_tile_list + ntiles);
      uint32_t* d_dense_count = reinterpret_cast<uint32_t*>(d_counts + 5);
      SB_CUDA(cudaMemsetAsync(d_tile_counts, 0, (size_t)ntiles * 9, stream_));
      SB_CUDA(cudaMemsetAsync(d_dense_count, 0, sizeof(unsigned long long), stream_));
      ScanArgs h = v;
      h.hit_keys = hits_.as<uint64_t>();
      h.hit_count = d_hit_count;
      h.tile_bytes = tile_bytes;
      SB_CUDA(launch_tile_marks(h, ntiles, d_tile_counts, /*min_hits=*/0, (uint32_t)std::min(max_hits, 4.0e9), d_dense,
                                d_tile_list, d_dense_count, stream_));
      stats_.aux_launches += 2;
      v.tile_bytes = tile_bytes;
      v.dense = d_dense;
      v.guard_limit = 0;
      a.tile_list = d_tile_list;
      a.tile_count = d_dense_count;
    };
    auto scan_dense_tiles = [&]() {  // the listed tiles, with the exact recurrences (both directions)
      if (!a.tile_list) return;
      CUtensorMap tmap;
      memset(&tmap, 0, sizeof tmap);
      if (variant_ == kVariantTma) make_tensor_map(&tmap, text, g);
      ScanArgs sd = a;
      if (nfwd) {
        sd.reset_idx = 0, sd.nq = nfwd, sd.qs_base = 0, sd.eq = d_eq;
        SB_CUDA(launch_scan(W, false, variant_, &tmap, sd, stream_));
        stats_.scan_launches++;
      }
      if (nq > nfwd) {
        sd.reset_idx = n - 1, sd.nq = nq - nfwd, sd.qs_base = nfwd, sd.eq = d_eq + (size_t)nfwd * nrows_ * W;
        SB_CUDA(launch_scan(W, true, variant_, &tmap, sd, stream_));
        stats_.scan_launches++;
      }
    };
    if (fused)  // a reversed query's hit marks the START of its piece in scan direction
      for (int p = 0; p < fp.npieces; p++) v.rev_lead = std::max<uint32_t>(v.rev_lead, (uint32_t)fp.piece[p].len);
    // (piece-automaton hits ARE share occurrences: refining them costs a pass over ~10^6 hits and
    //  buys shorter but unaligned windows -- measured slower on c2; q-gram hits are mostly false)
    const bool refine = conf_pieces_ > 0 && (refine_mode_ == 2 || (refine_mode_ == 1 && qgram));
    // refine (Dna q-gram hits) -> verify -> scan of dense tiles (regional pass only) -> tail -> counters
    auto rescan_pass = [&]() {
      ScanArgs vv = v;
      if (refine) {
        // every hit is refined exactly by one thread -- which share of the pattern occurs behind it,
        // and where -- and the (few) survivors are written to a second list as nominal end
        // positions: the re-scan covers 2k + 1 end positions per entry instead of 16 + m + k, on
        // dense warps (refine_hit in scan_core.cuh)
        hits2_.ensure(hit_cap_ * (sizeof(uint64_t) + sizeof(uint32_t)));
        uint32_t* spans = reinterpret_cast<uint32_t*>(hits2_.as<uint64_t>() + hit_cap_);
        ScanArgs cf = v;
        cf.qconf = reinterpret_cast<const uint32_t*>(dst + off_qconf_);
        cf.qnp = (uint32_t)conf_pieces_;
        SB_CUDA(cudaMemsetAsync(d_counts + 4, 0, sizeof(unsigned long long), stream_));
        SB_CUDA(launch_refine(cf, d_rev, hits2_.as<uint64_t>(), spans, d_counts + 4, stream_));
        stats_.aux_launches++;
        vv.hit_keys = hits2_.as<uint64_t>();
        vv.hit_span = spans;
        vv.hit_count = d_counts + 4;
        vv.hit_exact = 1;
        vv.rev_lead = 0;
        vv.dense = nullptr;  // dropped during the refinement already
      }
      vv.cand_keys = a.cand_keys, vv.cand_cost = a.cand_cost, vv.cand_cap = a.cand_cap;
      SB_CUDA(launch_verify(W, vv, d_rev, stream_));
      stats_.aux_launches++;
      scan_dense_tiles();
      SB_CUDA(cudaEventRecord(ev_[4], stream_));
      queue_small_tail();
      read_counts();
    };
    rescan_pass();
    if (h_counts[2] > heavy_hits && h_counts[2] <= hit_cap_ && !huge) {
      // many hits: second pass with the dense tiles scanned whole (the hit list is still valid)
      stats_.retries++;
      reset_candidates();
      enable_regional();
      rescan_pass();
    }
    stats_.dense_tiles = a.tile_list ? (uint32_t)(h_counts[5] & 0xFFFFFFFFu) : 0u;
    stats_.filter_ms = elapsed(ev_[1], ev_[2]);
    stats_.verify_ms = elapsed(ev_[2], ev_[4]);
    unsigned long long nhits = h_counts[2];
    stats_.hits = nhits;
    stats_.confirmed = refine ? h_counts[4] : nhits;
    stats_.filter_words = qgram ? 1u : (uint32_t)WT;
    stats_.filter_len = qgram ? (uint32_t)qp.q : (uint32_t)fp.L;
    stats_.filter_kind = qgram ? 2u : 1u;
    // too many hits (repetitive text, unlucky pieces): the re-scan costs more than the scan
    // (q-gram hits are confirmed first: ~64 character-steps each unless a whole share is there)
    const double rescan = refine ? (double)nhits * 24.0 + (double)h_counts[4] * (2.0 * m + 3.0 * k)
                                 : (double)nhits * (2.0 * (m + k) + kHitChars);
    if (refine && h_counts[4] > hit_cap_) nhits = hit_cap_ + 1;  // refined list overflowed: as a hit overflow
    if (pg_check()) {  // every rank's result is complete and already gathered
      stats_.ltot = gf.ltot;
      stats_.rows = gf.rows;
      stats_.blocks_per_sm = (uint32_t)focc;
      ncand = h_counts[0];
      filtered = true;
      small_done = true;
      stats_.scan_ms = stats_.filter_ms;
    } else if (nhits > hit_cap_ || (huge && nhits > heavy_hits)) {
      // the hit list overflowed: hits are lost, only the full scan is exact (more than 32 words: there
      // is no regional pass, the guarded first pass did not run)
      (void)rescan;
      stats_.filter_fallback = 1;
    } else {
      stats_.ltot = gf.ltot;
      stats_.rows = gf.rows;
      stats_.blocks_per_sm = (uint32_t)focc;
      for (int attempt = 0;; attempt++) {
        const unsigned long long cnt = h_counts[0];
        if (cnt <= cand_cap_) {
          ncand = cnt;
          break;
        }
        if (attempt >= 3) throw CudaError("candidate buffer overflow after retries");
        cand_cap_ = (size_t)(cnt + cnt / 8 + 1024);
        stats_.retries++;
        reset_candidates();
        SB_CUDA(cudaEventRecord(ev_[2], stream_));
        rescan_pass();
        stats_.verify_ms += elapsed(ev_[2], ev_[4]);
      }
      filtered = true;
      small_done = small_path && h_counts[3] == 0;
      stats_.scan_ms = stats_.filter_ms;  // the dominant kernel of this route
    }
  }

  // ---- candidates, route 2: full scan with the bit-parallel recurrences ---------------------
  a.tile_list = nullptr, a.tile_count = nullptr;
  if (!filtered) {
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof tmap);
    if (variant_ == kVariantTma && n > 0) make_tensor_map(&tmap, text, g);
    for (int attempt = 0;; attempt++) {
      reset_candidates();
      SB_CUDA(cudaEventRecord(ev_[1], stream_));
      if (n > 0 && huge) {
        // more than 32 words: windows of kCoverStride end positions (plus m + k characters of
        // warm-up each) cover the text, one warp per window and query (verify_wide_kernel)
        const uint64_t per = (n + kCoverStride - 1) / kCoverStride;
        const uint64_t entries = per * nq;
        hits_.ensure(entries * sizeof(uint64_t));
        hits2_.ensure(entries * sizeof(uint32_t));
        SB_CUDA(launch_cover(hits_.as<uint64_t>(), hits2_.as<uint32_t>(), d_counts + 4, nq, n, kCoverStride,
                             (uint32_t)k, stream_));
        ScanArgs vv = a;
        vv.nq = nq, vv.qs_base = 0, vv.eq = d_eq;
        vv.hit_keys = hits_.as<uint64_t>();
        vv.hit_span = hits2_.as<uint32_t>();
        vv.max_span = (uint32_t)kCoverStride;
        vv.hit_count = d_counts + 4;
        vv.hit_cap = entries;
        vv.hit_exact = 1;
        SB_CUDA(launch_verify(W, vv, d_rev, stream_));
        stats_.aux_launches++;
        stats_.scan_launches++;
      } else if (n > 0) {
        // batches of one-word patterns: two patterns per thread (scan2_kernel)
        if (nfwd) {
          a.reset_idx = 0;
          a.nq = nfwd;
          a.qs_base = 0;
          a.eq = d_eq;
          if (W == 1 && nfwd >= 2 && scan2_)
            SB_CUDA(launch_scan2(false, variant_, &tmap, a, stream_));
          else
            SB_CUDA(launch_scan(W, false, variant_, &tmap, a, stream_));
          stats_.scan_launches++;
        }
        if (nq > nfwd) {
          a.reset_idx = n - 1;
          a.nq = nq - nfwd;
          a.qs_base = nfwd;
          a.eq = d_eq + (size_t)nfwd * nrows_ * W;
          if (W == 1 && nq - nfwd >= 2 && scan2_)
            SB_CUDA(launch_scan2(true, variant_, &tmap, a, stream_));
          else
            SB_CUDA(launch_scan(W, true, variant_, &tmap, a, stream_));
          stats_.scan_launches++;
        }
        stats_.swar_lanes = (W == 1 && scan2_ && (nfwd >= 2 || nq - nfwd >= 2)) ? 2u : 1u;
      }
      if (ov) {
        a.nq = nq;
        a.qs_base = 0;
        a.eq = d_eq;
        SB_CUDA(launch_overhang_edges(W, a, oa, stream_));
        stats_.aux_launches++;
      }
      SB_CUDA(cudaEventRecord(ev_[2], stream_));
      queue_small_tail();
      read_counts();
      stats_.scan_ms += elapsed(ev_[1], ev_[2]);
      const unsigned long long cnt = h_counts[0];
      if (pg_check()) {
        ncand = cnt;
        small_done = true;
        break;
      }
      if (cnt <= cand_cap_) {
        ncand = cnt;
        small_done = small_path && h_counts[3] == 0;
        break;
      }
      if (attempt >= 3) throw CudaError("candidate buffer overflow after retries");
      cand_cap_ = (size_t)(cnt + cnt / 8 + 1024);  // dense text: re-run with an exact-size buffer
      stats_.retries++;
    }
  }
  stats_.candidates = ncand;

  // ---- sort, select (de-duplicate + local minima + end filters), trace ------------------------
  uint64_t nsel = 0;
  if (small_done) {
    nsel = h_counts[1];
    const GpuMatch* src_m = h_small_out;
    const uint32_t* src_ops = h_small_ops;
    if (small_in_slot && pg->pipelined()) {
      nsel = 0;  // this step's records stay in the device slot; they reach the host with the next collect
    } else if (small_in_slot) {  // the collect kernel mirrored this rank's slot into pinned host memory
      const PeerGather::Slot sl = pg->slot(pg->rank());
      src_m = sl.records;
      src_ops = sl.ops;
    }
    out.m.assign(src_m, src_m + nsel);
    out.ops.assign(src_ops, src_ops + nsel * out.ops_words);
  } else if (ncand > 0) {
    PostCtx c;
    c.text = TextRef{text.d, n, nullptr, nullptr, nq};
    c.nslots = nq;
    c.m = m, c.k = k, c.W = W;
    c.d_pat = d_pat, c.d_rev = d_rev, c.d_eq = d_eq;
    c.d_counts = d_counts;
    c.dedup = filtered;
    c.end_bit = end_bit;
    nsel = post_process(c, opts, ncand, out);
  }
  if (!(small_done && tail_event)) {  // (the fast tail recorded the end and was synchronised already)
    SB_CUDA(cudaEventRecord(ev_[3], stream_));
    SB_CUDA(cudaStreamSynchronize(stream_));
  }
  float total = 0;
  SB_CUDA(cudaEventElapsedTime(&total, ev_[0], ev_[3]));
  stats_.total_ms = total;
  stats_.matches = nsel;
  if (timing) {
    const double t_end = HostTimers::now_us();
    HostTimers::add(HostTimers::kGpuWait, t_last_sync - t_uploaded);
    HostTimers::add(HostTimers::kPost, t_end - t_last_sync);
  }
  if (transfer_pending_) {  // the text of this search came from the host just before it
    SB_CUDA(cudaEventElapsedTime(&transfer_ms_, ev_[5], ev_[6]));
    transfer_pending_ = false;
    stats_.transfer_ms = transfer_ms_;
    stats_.transfer_packed = transfer_packed_ ? 1 : 0;
    stats_.transfer_bytes = transfer_bytes_;
  }
}

void Engine::flush_gather(PeerGather& pg) {
  SB_CUDA(cudaSetDevice(device_));
  SB_CUDA(pg.flush(stream_));
  SB_CUDA(cudaStreamSynchronize(stream_));
  if (pg.has_result() && pg.timed_out()) {
    pg.mark_broken();
    throw CudaError("peer gather timed out (SASSY_B200_GATHER_TIMEOUT_S) while flushing the pipeline");
  }
}

// Every query against every text.  The texts are packed back to back (16-byte aligned starts)
// into one device buffer per call; candidates carry slot = text * nq + query, and the shared
// post-processing resolves a slot back to its text through TextRef.
void Engine::search_texts(const uint8_t* const* texts, const uint64_t* lens, size_t ntexts,
                          const std::vector<Query>& queries, int m, int k, const SearchOpts& opts, MatchSet& out) {
  SB_CUDA(cudaSetDevice(device_));
  cudaGetLastError();
  stats_ = SearchStats();
  out.m.clear();
  out.ops.clear();
  const uint32_t nq = (uint32_t)queries.size();
  if (m <= 0) throw CudaError("empty pattern");
  const int W = round_words((m + 31) / 32);
  if (W < 0 || W > kMaxScanWords)
    throw CapacityError("pattern longer than 1024 characters is not supported in a search over many texts");
  if (k < 0) k = 0;
  out.ops_words = (uint32_t)((m + k + 1 + 15) / 16);
  stats_.words = (uint32_t)W;
  if (nq == 0 || ntexts == 0) return;
  const uint64_t nslots = (uint64_t)ntexts * nq;
  if (nslots >= (1ull << (64 - kPosBits))) throw CudaError("too many (text, query) pairs in one search");

  // ---- pack and upload the texts -------------------------------------------------------------
  const double t_stage0 = HostTimers::on() ? HostTimers::now_us() : 0.0;
  std::vector<uint64_t> meta(2 * ntexts);
  uint64_t total = 0;
  for (size_t i = 0; i < ntexts; i++) {
    if (lens[i] >= (1ull << kPosBits)) throw CapacityError("text longer than 2^40 bytes is not supported");
    meta[i] = total;
    meta[ntexts + i] = lens[i];
    total += (lens[i] + 15) & ~15ull;
  }
  // Enough work for the row-tiled kernels: the texts are scanned as ONE concatenated text, all
  // queries at once (2 T lane-steps/s instead of the 0.25 T of one thread per (text, query) pair
  // walking its text alone); see the candidate section below.  SASSY_B200_TEXTS_CONCAT: 0 never,
  // 2 always (tests).
  int concat_mode = 1;
  if (const char* e = getenv("SASSY_B200_TEXTS_CONCAT")) concat_mode = atoi(e);
  const bool concat = concat_mode != 0 && !(opts.alpha >= 0.f) && m > k && total > 0 &&
                      (concat_mode == 2 || total >= (1ull << 20));
  // the concatenation is tiled like any text: room for one extra row of any tiling behind it
  const size_t text_bytes = concat ? padded_alloc(total) : (size_t)total + 64;
  const size_t meta_off = (text_bytes + 255) & ~(size_t)255;
  const size_t packed_bytes = meta_off + meta.size() * sizeof(uint64_t);
  // staged in PINNED memory (kept between calls) by a few threads: hundreds of megabytes of reads
  // would otherwise cross PCIe from pageable memory at a fraction of the link rate
  if (packed_bytes > h_texts_cap_) {
    if (h_texts_) cudaFreeHost(h_texts_);
    h_texts_ = nullptr;
    h_texts_cap_ = 0;
    SB_CUDA(cudaHostAlloc((void**)&h_texts_, packed_bytes + packed_bytes / 4, cudaHostAllocDefault));
    h_texts_cap_ = packed_bytes + packed_bytes / 4;
  }
  uint8_t* packed = h_texts_;
  d_texts_.ensure(packed_bytes);
  SB_CUDA(cudaEventRecord(ev_[0], stream_));
  {
    // The texts are separate host buffers: a few threads copy them range by range (consecutive
    // texts, ~8 MB each) into the pinned staging buffer, and every finished range is handed to the
    // copy engine at once, so the upload runs behind the host copies instead of after them.
    const size_t nthreads =
        total > (8u << 20) ? std::min<size_t>(16, std::max(1u, std::thread::hardware_concurrency())) : 1;
    auto copy_range = [&](size_t t0, size_t t1) {
      for (size_t i = t0; i < t1; i++) {
        if (lens[i]) memcpy(packed + meta[i], texts[i], lens[i]);
        const uint64_t end = meta[i] + lens[i];
        const uint64_t next = i + 1 < ntexts ? meta[i + 1] : text_bytes;
        if (next > end) memset(packed + end, 0, next - end);  // alignment padding reads as zero bytes
      }
    };
    if (nthreads == 1) {
      copy_range(0, ntexts);
      if (meta_off > text_bytes) memset(packed + text_bytes, 0, meta_off - text_bytes);
      memcpy(packed + meta_off, meta.data(), meta.size() * sizeof(uint64_t));
      SB_CUDA(cudaMemcpyAsync(d_texts_.p, packed, packed_bytes, cudaMemcpyHostToDevice, stream_));
    } else {
      // ranges of consecutive texts with about the same number of bytes
      const size_t nranges = std::min<size_t>(ntexts, std::max<size_t>(nthreads * 2, (size_t)(total >> 23)));
      std::vector<size_t> first(nranges + 1, ntexts);
      first[0] = 0;
      {
        size_t r = 1;
        for (size_t i = 0; i < ntexts && r < nranges; i++)
          if (meta[i] >= total / nranges * r) first[r++] = i;
      }
      std::unique_ptr<std::atomic<uint8_t>[]> done(new std::atomic<uint8_t>[nranges]);
      for (size_t r = 0; r < nranges; r++) done[r].store(0);
      std::atomic<size_t> next_range{0};
      std::vector<std::thread> th;
      for (size_t t = 0; t < nthreads; t++)
        th.emplace_back([&] {
          for (;;) {
            const size_t r = next_range.fetch_add(1);
            if (r >= nranges) return;
            copy_range(first[r], first[r + 1]);
            done[r].store(1, std::memory_order_release);
          }
        });
      cudaError_t err = cudaSuccess;
      for (size_t r = 0; r < nranges; r++) {
        while (done[r].load(std::memory_order_acquire) == 0) std::this_thread::sleep_for(std::chrono::microseconds(20));
        const uint64_t b0 = first[r] < ntexts ? meta[first[r]] : text_bytes;
        const uint64_t b1 = first[r + 1] < ntexts ? meta[first[r + 1]] : text_bytes;
        if (b1 > b0 && err == cudaSuccess)
          err = cudaMemcpyAsync(d_texts_.as<uint8_t>() + b0, packed + b0, b1 - b0, cudaMemcpyHostToDevice, stream_);
      }
      for (auto& x : th) x.join();
      SB_CUDA(err);
      if (meta_off > text_bytes) memset(packed + text_bytes, 0, meta_off - text_bytes);
      memcpy(packed + meta_off, meta.data(), meta.size() * sizeof(uint64_t));
      SB_CUDA(cudaMemcpyAsync(d_texts_.as<uint8_t>() + text_bytes, packed + text_bytes, packed_bytes - text_bytes,
                              cudaMemcpyHostToDevice, stream_));
    }
  }
  const double t_stage1 = HostTimers::on() ? HostTimers::now_us() : 0.0;
  if (HostTimers::on()) HostTimers::add(HostTimers::kTextsStage, t_stage1 - t_stage0);
  const uint8_t* d_base = d_texts_.as<uint8_t>();
  const uint64_t* d_offs = reinterpret_cast<const uint64_t*>(d_base + meta_off);
  const uint64_t* d_lens = d_offs + ntexts;

  upload_params(queries, m, W, FilterPlan(), false, false);
  uint8_t* dst = d_stage_.as<uint8_t>();
  unsigned long long* d_counts = reinterpret_cast<unsigned long long*>(dst + off_counts_);
  const uint32_t* d_eq = reinterpret_cast<const uint32_t*>(dst + off_eq_);
  const uint8_t* d_pat = dst + off_pat_;
  const uint8_t* d_rev = dst + off_rev_;

  if (cand_cap_ == 0) cand_cap_ = 1ull << 20;
  ScanArgs a;
  memset(&a, 0, sizeof a);
  a.sh0 = sh0_;
  a.msk0 = msk0_;
  a.nrows = nrows_;
  a.rowbytes = (uint32_t)W * 4u;
  a.m = m;
  a.k = k;
  a.nq = nq;
  a.eq = d_eq;
  a.cand_count = d_counts;
  TextsArgs t;
  memset(&t, 0, sizeof t);
  t.base = d_base;
  t.offs = d_offs;
  t.lens = d_lens;
  t.rev_flags = d_rev;
  t.ntexts = (uint32_t)ntexts;
  t.nq = nq;
  t.include_pos0 = opts.include_pos0 ? 1 : 0;
  const bool ov = opts.alpha >= 0.f;
  if (ov && opts.pam_len > 0) throw CudaError("an end filter cannot be combined with overhang");
  t.overhang = ov ? 1 : 0;
  OverhangArgs oa;
  if (ov) {
    overhang_args(opts, m, k, W, oa);
    oa.text = TextRef{d_base, 0, d_offs, d_lens, nq};
    oa.rev_flags = d_rev;
    oa.nslots = (uint32_t)nslots;
  }

  unsigned long long h_counts[4] = {0, 0, 0, 0};
  uint64_t ncand = 0;
  // Concatenated scan: a cost <= k inside a text can only be LOWERED by what the scan carries over
  // from the previous text (the fresh state D[j][0] = j is the largest possible column), and only
  // within the first m + k end positions of the text (an alignment of cost <= k spans at most
  // m + k characters).  So: the row-tiled scan reports every candidate of the concatenation into a
  // raw list; concat_remap_kernel keeps those beyond the first m + k end positions of their text
  // (as (text, query) slots, text-relative positions) and texts_kernel computes the first m + k end
  // positions of every text from a fresh state.  Both directions: "first" is in scan direction.
  uint32_t nfwd = 0;
  while (nfwd < nq && !queries[nfwd].rev) nfwd++;
  DeviceText ctext;
  ScanGeom cg;
  CUtensorMap ctmap;
  memset(&ctmap, 0, sizeof ctmap);
  if (concat) {
    for (uint32_t q = nfwd; q < nq; q++)
      if (!queries[q].rev) throw CudaError("forward queries must precede reversed ones");
    ctext.d = d_texts_.as<uint8_t>();
    ctext.n = total;
    ctext.alloc = text_bytes;
    ctext.owned = false;
    const int occ = scan_blocks_per_sm(W, false, variant_, nrows_);
    auto launch_cols = [&](uint32_t cnt) { return (W == 1 && cnt >= 2 && scan2_) ? (cnt + 1) / 2 : cnt; };
    cg = choose_geom(total, m, k, std::max<uint32_t>(1, std::max(launch_cols(nfwd), launch_cols(nq - nfwd))),
                     occ * sm_count_);
    if ((uint64_t)cg.rows * cg.ltot > ctext.alloc) throw CudaError("internal: text padding too small for tiling");
    if (variant_ == kVariantTma) make_tensor_map(&ctmap, ctext, cg);
    stats_.ltot = cg.ltot;
    stats_.blocks_per_sm = (uint32_t)occ;
  }
  for (int attempt = 0;; attempt++) {
    keys_.ensure(cand_cap_ * sizeof(uint64_t));
    cost_.ensure(cand_cap_ * sizeof(uint32_t));
    a.cand_keys = keys_.as<uint64_t>();
    a.cand_cost = cost_.as<uint32_t>();
    a.cand_cap = cand_cap_;
    if (attempt > 0) SB_CUDA(cudaMemsetAsync(d_counts, 0, 4 * sizeof(unsigned long long), stream_));
    SB_CUDA(cudaEventRecord(ev_[1], stream_));
    if (concat) {
      keys2_.ensure(cand_cap_ * sizeof(uint64_t));
      cost2_.ensure(cand_cap_ * sizeof(uint32_t));
      ScanArgs c = a;  // the concatenation as one text; raw candidates, counted in d_counts[2]
      c.text = ctext.d;
      c.n = total;
      c.g = cg;
      c.cand_keys = keys2_.as<uint64_t>();
      c.cand_cost = cost2_.as<uint32_t>();
      c.cand_count = d_counts + 2;
      if (nfwd) {
        c.reset_idx = 0, c.nq = nfwd, c.qs_base = 0, c.eq = d_eq;
        if (W == 1 && nfwd >= 2 && scan2_)
          SB_CUDA(launch_scan2(false, variant_, &ctmap, c, stream_));
        else
          SB_CUDA(launch_scan(W, false, variant_, &ctmap, c, stream_));
        stats_.scan_launches++;
      }
      if (nq > nfwd) {
        c.reset_idx = total - 1, c.nq = nq - nfwd, c.qs_base = nfwd, c.eq = d_eq + (size_t)nfwd * nrows_ * W;
        if (W == 1 && nq - nfwd >= 2 && scan2_)
          SB_CUDA(launch_scan2(true, variant_, &ctmap, c, stream_));
        else
          SB_CUDA(launch_scan(W, true, variant_, &ctmap, c, stream_));
        stats_.scan_launches++;
      }
      stats_.swar_lanes = (W == 1 && scan2_ && (nfwd >= 2 || nq - nfwd >= 2)) ? 2u : 1u;
      const uint64_t skip = (uint64_t)m + (uint64_t)k;
      SB_CUDA(launch_concat_remap(keys2_.as<uint64_t>(), cost2_.as<uint32_t>(), d_counts + 2, cand_cap_, a, t, total,
                                  skip, stream_));
      stats_.aux_launches++;
      t.prefix = (uint32_t)skip;
    }
    SB_CUDA(launch_texts(W, a, t, stream_));
    stats_.scan_launches++;
    if (ov) {
      SB_CUDA(launch_overhang_edges(W, a, oa, stream_));
      stats_.aux_launches++;
    }
    SB_CUDA(cudaEventRecord(ev_[2], stream_));
    SB_CUDA(cudaMemcpyAsync(h_counts, d_counts, sizeof h_counts, cudaMemcpyDeviceToHost, stream_));
    SB_CUDA(cudaStreamSynchronize(stream_));
    float ms = 0;
    SB_CUDA(cudaEventElapsedTime(&ms, ev_[1], ev_[2]));
    stats_.scan_ms += ms;
    const unsigned long long most = std::max(h_counts[0], concat ? h_counts[2] : 0ull);
    if (most <= cand_cap_) {
      ncand = h_counts[0];
      break;
    }
    if (attempt >= 3) throw CudaError("candidate buffer overflow after retries");
    cand_cap_ = (size_t)(most + most / 8 + 1024);
    stats_.retries++;
  }
  stats_.candidates = ncand;
  uint64_t nsel = 0;
  if (ncand > 0) {
    PostCtx c;
    c.text = TextRef{d_base, 0, d_offs, d_lens, nq};
    c.nslots = (uint32_t)nslots;
    c.m = m, c.k = k, c.W = W;
    c.d_pat = d_pat, c.d_rev = d_rev, c.d_eq = d_eq;
    c.d_counts = d_counts;
    c.dedup = false;
    c.end_bit = kPosBits;
    while ((1ull << (c.end_bit - kPosBits)) < nslots) c.end_bit++;
    nsel = post_process(c, opts, ncand, out);
  }
  SB_CUDA(cudaEventRecord(ev_[3], stream_));
  SB_CUDA(cudaStreamSynchronize(stream_));
  float total_ms = 0;
  SB_CUDA(cudaEventElapsedTime(&total_ms, ev_[0], ev_[3]));
  stats_.total_ms = total_ms;
  stats_.matches = nsel;
  stats_.rows = (uint32_t)ntexts;
  if (HostTimers::on()) HostTimers::add(HostTimers::kTextsGpu, HostTimers::now_us() - t_stage1);
}

}  // namespace sb
