// Host side of the GPU search path: owns the device buffers of one searcher
// and runs scan -> sort -> local-minima -> traceback on one CUDA stream.
// Mirrors the role of the reference's `Searcher<P>` internals
// (src/search.rs:227-256 scratch buffers, :884-937 search_one_strand,
//  :1372-1517 process_matches); the public surface is in searcher.h / c_api.cu.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "peer_gather.h"
#include "transport.h"

namespace sb {

struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
// A size limit of this implementation (not an error of the reference): the C ABI's search()
// reports it without aborting the process (include/sassy.h).
// Host-side phase timers of the search entry points (SASSY_B200_HOST_TIMING=1 prints the averages
// to stderr at exit; profiling aid, not part of the product path's behaviour).
struct HostTimers {
  enum { kPre, kGpuWait, kPost, kMerge, kSearchCall, kTextsStage, kTextsGpu, kTextsConvert, kCount };
  static bool on();
  static void add(int which, double us);
  static double now_us();
};

struct CapacityError : CudaError {
  using CudaError::CudaError;
};

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  void ensure(size_t bytes);
  void release();
  template <class T> T* as() const { return static_cast<T*>(p); }
};

// A text resident in HBM, padded with zero bytes so that any row tiling up to
// kMaxRowBytes can be laid over it (the device analogue of the reference's
// CachedRev, src/search.rs:144-166: upload once, search many times, both strands).
struct DeviceText {
  uint8_t* d = nullptr;
  uint64_t n = 0;
  size_t alloc = 0;
  bool owned = true;
};

// One searched sequence: the bytes the scan compares against and the direction.
struct Query {
  const uint8_t* bytes;  // m bytes
  bool rev;              // scan the reversed text (v1 reverse-complement strand)
};

struct SearchStats {
  float scan_ms = 0;      // scan kernel(s) only (CUDA events on the engine stream)
  float total_ms = 0;     // scan + sort + minima + traceback + result copies
  uint32_t scan_launches = 0;
  uint32_t aux_launches = 0;  // minima + trace launches of ours
  uint64_t candidates = 0;
  uint64_t matches = 0;
  uint32_t ltot = 0, rows = 0, words = 0, blocks_per_sm = 0;
  uint32_t retries = 0;
  // exact piece prefilter (0 words = full scan)
  float filter_ms = 0;   // prefilter kernel(s)
  float verify_ms = 0;   // re-scan of the hit neighbourhoods
  uint64_t hits = 0;
  uint32_t filter_words = 0, filter_len = 0;
  uint32_t filter_fallback = 0;  // prefilter ran but produced too many hits: full scan used
  float transfer_ms = 0;         // host->device transfer of the text, when the search got a host text
  uint32_t transfer_packed = 0;  // 1: sent at 2 bits per character (Dna transport encoding)
  uint64_t transfer_bytes = 0;   // bytes that crossed PCIe for the text
  uint32_t filter_kind = 0;      // 0 none, 1 piece automaton (Shift-And), 2 q-gram bitmap, 3 SWAR suffix scan
  uint32_t swar_lanes = 0;       // patterns per THREAD of the scan that produced the candidates (0/1 = one, 2 = scan2_kernel)
  uint64_t confirmed = 0;        // prefilter hits that were re-scanned (q-gram: after the exact confirmation)
  uint32_t dense_tiles = 0;      // tiles of the scan geometry that were scanned whole (regional fallback)
};

struct MatchSet {
  std::vector<GpuMatch> m;
  std::vector<uint32_t> ops;  // m.size() * ops_words, 2-bit op codes
  uint32_t ops_words = 0;
};

// Per-search options (reference Searcher fields src/search.rs:227-256 and the arguments of
// search / search_all / search_with_fn).
struct SearchOpts {
  bool all_minima = false;     // search_all: every end position with cost <= k
  bool include_pos0 = false;   // v1: end position 0 is a candidate when m <= k (src/search.rs:1320-1322)
  bool without_trace = false;  // Searcher::without_trace (src/search.rs:446-449)
  bool only_best = false;      // Searcher::only_best_match (src/search.rs:441-444)
  float max_n_frac = -1.f;     // Searcher::set_max_n_frac (src/search.rs:452-458); < 0 = off
  bool n_endpoint = false;     // also apply the pre-trace N end-point filter (v1, src/search.rs:907-919)
  const uint8_t* pam = nullptr;  // end filter of search_with_fn as used by bin/crispr.rs:198-205
  int pam_len = 0;
  float alpha = -1.f;          // overhang cost per pattern character (src/search.rs:231-233); < 0 = off
  int max_overhang = -1;       // Searcher::with_max_overhang (src/search.rs:436-439); < 0 = unlimited
  bool special() const {
    return without_trace || only_best || max_n_frac >= 0.f || pam_len > 0 || alpha >= 0.f;
  }
};

class Engine {
 public:
  Engine(int profile, int device);
  ~Engine();
  Engine(const Engine&) = delete;
  Engine& operator=(const Engine&) = delete;

  int profile() const { return profile_; }
  int device() const { return device_; }
  cudaStream_t stream() const { return stream_; }

  DeviceText* upload_text(const uint8_t* host, uint64_t n);        // host -> HBM (new buffer)
  DeviceText* adopt_device_text(const void* dptr, uint64_t n);      // device -> padded device copy
  void free_text(DeviceText* t);
  // Re-usable staging text for the plain search(pattern, text) ABI.
  DeviceText* stage_text(const uint8_t* host, uint64_t n);
  // 0 = always send bytes; 1 = send large Dna texts at 2 bits per character (default)
  void set_transport(int mode) { transport_mode_ = mode; }
  float last_transfer_ms() const { return transfer_ms_; }
  bool last_transfer_packed() const { return transfer_packed_; }

  // Queries must all have length m; forward queries must precede reversed ones.
  // include_pos0: also consider end position 0 (cost m) when m <= k (v1 only,
  // reference src/search.rs:1320-1322).
  void search(const DeviceText& text, const std::vector<Query>& queries, int m, int k, const SearchOpts& opts,
              MatchSet& out);
  // Every query against every text: one thread per (text, query) pair; GpuMatch::qs =
  // text index * queries.size() + query index.  For many short texts (reference
  // Searcher::search_texts / search_many, src/search.rs:531-640).
  void search_texts(const uint8_t* const* texts, const uint64_t* lens, size_t ntexts,
                    const std::vector<Query>& queries, int m, int k, const SearchOpts& opts, MatchSet& out);

  // Multi-GPU: while a PeerGather is attached every search() ends with one exchange of the
  // match records over peer memory (all ranks must search in lock step).  gather_ok() tells
  // whether the last search left every rank's complete result in pg->slot(r); if not, `out`
  // holds the local result and the caller falls back to its own collective.
  void set_gather(PeerGather* pg, uint64_t user) { pg_ = pg, pg_user_ = user; }
  bool gather_ok() const { return gather_ok_; }
  // Pipelined PeerGather: collect the last pushed step into the host mirror (one synchronisation).
  void flush_gather(PeerGather& pg);

  const SearchStats& stats() const { return stats_; }
  void set_variant(int v) { variant_ = v; }
  int variant() const { return variant_; }
  // 0 = never prefilter, 1 = prefilter when the expected re-scan work is small (default),
  // 2 = prefilter whenever a piece layout exists (tests)
  void set_filter_mode(int m) { filter_mode_ = m; }

 private:
  struct PostCtx {
    TextRef text;
    uint32_t nslots = 0;  // query slots (queries, or texts x queries)
    int m = 0, k = 0, W = 0;
    const uint8_t* d_pat = nullptr;
    const uint8_t* d_rev = nullptr;
    const uint32_t* d_eq = nullptr;
    unsigned long long* d_counts = nullptr;
    bool dedup = false;   // candidate list may hold a position more than once
    int end_bit = 64;
  };
  uint64_t post_process(const PostCtx& c, const SearchOpts& opts, uint64_t ncand, MatchSet& out);
  // Overhang: fills the arguments of the edge kernel (left-column deltas, wildcard steps).
  void overhang_args(const SearchOpts& opts, int m, int k, int W, OverhangArgs& o) const;
  void upload_params(const std::vector<Query>& queries, int m, int W, const FilterPlan& fp, bool pair, bool fused,
                     const QgramPlan* qp = nullptr);
  void make_tensor_map(CUtensorMap* map, const DeviceText& text, const ScanGeom& g) const;
  // host -> dst (device, padded): packed transport for large Dna texts, else plain copies
  void send_text(uint8_t* dst, size_t dst_alloc, const uint8_t* host, uint64_t n);

  int profile_;
  int device_;
  int variant_;
  PeerGather* pg_ = nullptr;  // not owned
  uint64_t pg_user_ = 0;
  bool gather_ok_ = false;
  int filter_mode_ = 1;
  bool fuse_strands_ = true;
  int pair_max_words_ = 4;  // Dna: two characters per automaton step up to this many words
  int qgram_mode_ = 1;      // 0: never use the q-gram bitmap prefilter (SASSY_B200_QGRAM=0), 1: when planned
  int qgram_min_q_ = 6;     // SASSY_B200_QGRAM_MIN_Q
  bool qgram_seq_ = true;   // contiguous-tile q-gram kernel (SASSY_B200_QGRAM_SEQ=0: row-tiled kernel)
  size_t off_qconf_ = 0;
  int conf_pieces_ = 0;       // pieces per query in the refinement records of the last upload (0: none)
  bool scan2_ = true;         // SASSY_B200_SCAN2=0: one pattern per thread for batches of one-word patterns too
  int refine_mode_ = 1;       // SASSY_B200_REFINE: 0 never refine hits, 1 q-gram hits (default), 2 piece-automaton hits too
  int filter_row_bytes_ = 0;  // SASSY_B200_FILTER_ROW_BYTES (experiments): bytes per thread row of the prefilter
  int transport_mode_ = 1;
  float transfer_ms_ = 0;
  bool transfer_pending_ = false;
  uint64_t transfer_bytes_ = 0;
  bool transfer_packed_ = false;
  PackPool* pool_ = nullptr;
  uint8_t* h_pack_ = nullptr;  // pinned staging of the packed text
  size_t h_pack_cap_ = 0;
  DevBuf d_pack_;
  cudaStream_t unpack_stream_ = nullptr;  // expansion of the packed text, group by group behind the copies
  std::vector<cudaEvent_t> copy_ev_;      // one per host->device copy in flight (+ 2 for the expansion hand-over)
  int sm_count_ = 148;
  cudaStream_t stream_ = nullptr;
  cudaEvent_t ev_[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  uint32_t nrows_, sh0_, msk0_;

  DevBuf keys_, cost_, keys2_, cost2_, flags_, sel_, cubtmp_;
  DevBuf scratch_, ops_, out_;
  DevBuf hits_, hits2_, tiles_, d_stage_, sel_small_, best_, sel_cost_, d_texts_;
  uint8_t* h_small_ = nullptr;  // pinned: results of the small-list fast path, written by the GPU
  size_t h_small_cap_ = 0;
  uint8_t* h_texts_ = nullptr;  // pinned staging of search_texts (texts back to back + offsets)
  size_t h_texts_cap_ = 0;
  uint8_t* h_stage_ = nullptr;  // pinned staging for the per-search parameter block
  size_t stage_cap_ = 0;
  size_t off_counts_ = 0, off_eq_ = 0, off_pat_ = 0, off_rev_ = 0, off_feq_ = 0;
  uint64_t hit_cap_ = 0;
  uint64_t cand_cap_ = 0;
  DeviceText staged_;
  SearchStats stats_;
  void* encode_tiled_ = nullptr;
};

}  // namespace sb
