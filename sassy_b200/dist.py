"""Multi-GPU plumbing for the search path: one process per GPU (torchrun), independent
(pattern shard | text shard) searches per rank, and ONE collective at the end: an
all-gather of fixed-size match records (NCCL over NVLink when the tensors are on the GPU,
gloo on CPU for the host-logic tests).  The reference has no distributed layer; its
closest analogue is the rayon fan-out of Searcher::search_many (src/search.rs:531-603,
1520-1550), whose result is likewise the concatenation of per-task match lists."""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist

from .searcher import Match

_OPS = "=XID"
_REC = 8  # int64 fields per record header


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Round-robin shard: item i goes to rank i % world."""
    return list(range(rank, n_items, world))


def pack_matches(matches: Sequence[Match], ops_words: int) -> torch.Tensor:
    """[n, 8 + ops_words] int64: header + 2-bit packed CIGAR ops (32 ops per word)."""
    t = torch.zeros((len(matches), _REC + ops_words), dtype=torch.int64)
    for i, m in enumerate(matches):
        ops = m._ops
        if len(ops) > 32 * ops_words:
            raise ValueError("ops_words too small for CIGAR")
        row = [m.pattern_idx, m.text_idx, m.text_start, m.text_end, m.pattern_start, m.pattern_end,
               m.cost, (len(ops) << 1) | (1 if m.strand == "-" else 0)]
        words = [0] * ops_words
        for a, ch in enumerate(ops):
            words[a >> 5] |= _OPS.index(ch) << ((a & 31) * 2)
        # keep within signed int64
        words = [w - (1 << 64) if w >= (1 << 63) else w for w in words]
        t[i] = torch.tensor(row + words, dtype=torch.int64)
    return t


def unpack_matches(t: torch.Tensor) -> List[Match]:
    out = []
    rows = t.tolist()
    for r in rows:
        nops = r[7] >> 1
        words = [w + (1 << 64) if w < 0 else w for w in r[_REC:]]
        ops = "".join(_OPS[(words[a >> 5] >> ((a & 31) * 2)) & 3] for a in range(nops))
        out.append(Match(r[0], r[1], r[2], r[3], r[4], r[5], r[6], "-" if r[7] & 1 else "+", ops))
    return out


def gather_matches(matches: Sequence[Match], max_ops: int, device=None, group=None) -> List[Match]:
    """All ranks receive the concatenation (rank order) of every rank's matches.

    Two collectives of which only the second carries data: an all-gather of the counts
    (8 bytes per rank) and one all-gather of the records padded to the largest count
    (NCCL has no gatherv)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return list(matches)
    world = dist.get_world_size(group)
    ops_words = (max_ops + 31) // 32
    local = pack_matches(matches, ops_words)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else "cpu"
    cnt = torch.tensor([local.shape[0]], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(counts, cnt, group=group)
    counts = [int(c.item()) for c in counts]
    mx = max(counts)
    if mx == 0:
        return []
    buf = torch.zeros((mx, local.shape[1]), dtype=torch.int64, device=device)
    buf[: local.shape[0]] = local.to(device)
    allbuf = torch.empty((world * mx, local.shape[1]), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allbuf, buf, group=group)
    allbuf = allbuf.cpu().view(world, mx, local.shape[1])
    out: List[Match] = []
    for r in range(world):
        out.extend(unpack_matches(allbuf[r, : counts[r]]))
    return out
