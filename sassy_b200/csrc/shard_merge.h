// One text cut into slabs over several GPUs (SURVEY 8e "Partitioning"): host-side merge.
//
// Every rank searches its slab plus an (m + k) halo on both sides with search_all semantics
// (every end position with cost <= k, traced).  Inside the slab those values are exact: an
// alignment of cost <= k spans at most m + k characters, the argument of the reference's own
// lane overlap (src/search.rs:1018-1049).  The records of all ranks are gathered, the ones a
// slab does not own are dropped (src/search.rs:1202-1240 prunes lane overlaps the same way),
// coordinates become global, and the local-minima rule (src/search.rs:1344-1368 ==
// src/pattern_tiling/minima.rs:9-52, select_candidate in scan_core.cuh) runs on the merged list,
// so a plateau or a run of minima that crosses a slab border is selected exactly as in an
// unsharded search.  Pure host code: no device is touched.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/sassy_gpu.h"
#include "scan_core.cuh"

namespace sb {

struct SlabInfo {
  uint64_t window_off;  // global position of the slab window's first character
  uint64_t own_lo;      // the slab owns the global text range [own_lo, own_hi)
  uint64_t own_hi;
};

// recs[i].text_idx = slab index, coordinates relative to that slab's window, strand 0/1 with the
// v1 meaning (reverse-complement matches were found on the reversed window and mapped back).
// Returns the indices of the kept records in output order (forward matches by ascending end,
// then reverse-complement matches by ascending end on the reversed text, per pattern) and
// rewrites their coordinates to the global text.
template <class Rec>
inline std::vector<size_t> merge_slab_matches(std::vector<Rec>& recs, const SlabInfo* slabs, size_t n_slabs,
                                              uint64_t n_global, bool all_minima) {
  struct Item {
    uint64_t key;  // (pattern, strand) slot << kPosBits | end position in scan direction
    uint32_t cost;
    size_t idx;
  };
  std::vector<Item> items;
  items.reserve(recs.size());
  for (size_t i = 0; i < recs.size(); i++) {
    Rec& r = recs[i];
    if (r.text_idx >= n_slabs) continue;
    const SlabInfo& s = slabs[r.text_idx];
    r.text_start += s.window_off;
    r.text_end += s.window_off;
    r.text_idx = 0;
    uint64_t pos;
    bool own;
    if ((int)r.strand == 0) {
      // forward: end position e in (own_lo, own_hi]; e = 0 (empty prefix, m <= k) belongs to the first slab
      pos = r.text_end;
      own = (pos > s.own_lo && pos <= s.own_hi) || (pos == 0 && s.own_lo == 0);
    } else {
      // reversed text: end position e' = n - text_start; owned iff text_start in [own_lo, own_hi),
      // e' = 0 (text_start = n) belongs to the last slab
      pos = n_global - r.text_start;
      own = (r.text_start >= s.own_lo && r.text_start < s.own_hi) || (r.text_start == n_global && s.own_hi == n_global);
    }
    if (!own) continue;
    const uint64_t slot = r.pattern_idx * 2 + (uint64_t)(int)r.strand;
    items.push_back(Item{cand_key((uint32_t)slot, pos), (uint32_t)r.cost, i});
  }
  std::sort(items.begin(), items.end(), [](const Item& a, const Item& b) { return a.key < b.key; });
  std::vector<uint64_t> keys(items.size());
  std::vector<uint32_t> cost(items.size());
  for (size_t i = 0; i < items.size(); i++) keys[i] = items[i].key, cost[i] = items[i].cost;
  std::vector<size_t> keep;
  for (size_t i = 0; i < items.size(); i++)
    if (select_candidate(keys.data(), cost.data(), i, items.size(), all_minima)) keep.push_back(items[i].idx);
  // output order: per pattern forward before reverse (slot order), ascending scan position: the
  // order of a single-GPU search of one pattern; stable for several patterns
  return keep;
}

}  // namespace sb
