"""Multi-GPU plumbing for the search path: one process per GPU (torchrun), independent
(pattern shard | text shard) searches per rank, and ONE collective at the end: an
all-gather of fixed-size match records (NCCL over NVLink when the tensors are on the GPU,
gloo on CPU for the host-logic tests).  The reference has no distributed layer; its
closest analogue is the rayon fan-out of Searcher::search_many (src/search.rs:531-603,
1520-1550), whose result is likewise the concatenation of per-task match lists."""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
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


def _local_block(matches, max_ops: int):
    """(records uint8 [n, 72], ops uint8 [n, max_ops]) of a MatchList or a list of Match."""
    from .searcher import MatchList, _REC_DTYPE
    n = len(matches)
    if isinstance(matches, MatchList):
        recs = matches.records
        raw = np.frombuffer(matches._ops, dtype=np.uint8)
    else:
        recs = np.zeros(n, dtype=_REC_DTYPE)
        chunks = []
        off = 0
        for i, m in enumerate(matches):
            recs[i] = (m.pattern_idx, m.text_idx, m.text_start, m.text_end, m.pattern_start, m.pattern_end, m.cost,
                       1 if m.strand == "-" else 0, (0, 0, 0), len(m._ops), 0, off)
            chunks.append(m._ops.encode())
            off += len(m._ops)
        raw = np.frombuffer(b"".join(chunks), dtype=np.uint8)
    ops = np.zeros((n, max_ops), dtype=np.uint8)
    if n:
        lens = recs["ops_len"].astype(np.int64)
        if lens.max(initial=0) > max_ops:
            raise ValueError("max_ops too small for CIGAR")
        offs = recs["ops_off"].astype(np.int64)
        idx = offs[:, None] + np.arange(max_ops)[None, :]
        mask = np.arange(max_ops)[None, :] < lens[:, None]
        ops[mask] = raw[np.minimum(idx, max(len(raw) - 1, 0))][mask]
    return np.ascontiguousarray(recs).view(np.uint8).reshape(n, _REC_DTYPE.itemsize), ops


class _GatherBuffers:
    """Re-used staging for gather_matches: pinned host blocks and device blocks of one capacity."""

    def __init__(self, world: int, cap: int, width: int, device):
        self.cap = cap
        self.block = 8 + cap * width
        cuda = torch.device(device).type == "cuda"
        self.send_host = torch.zeros(self.block, dtype=torch.uint8, pin_memory=cuda)
        self.recv_host = torch.zeros(world * self.block, dtype=torch.uint8, pin_memory=cuda)
        if cuda:
            self.send_dev = torch.zeros(self.block, dtype=torch.uint8, device=device)
            self.recv_dev = torch.zeros(world * self.block, dtype=torch.uint8, device=device)
        else:
            self.send_dev, self.recv_dev = self.send_host, self.recv_host
        self.cuda = cuda


_GATHER = {}


def gather_matches(matches, max_ops: int, device=None, group=None):
    """All ranks receive the concatenation (rank order) of every rank's matches (a MatchList).

    ONE collective in the common case: every rank contributes a fixed-capacity block
    [count | records | ops]; only if some rank holds more matches than the capacity is the
    exchange repeated with a larger block (NCCL has no gatherv)."""
    from .searcher import MatchList, _REC_DTYPE
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return matches
    world = dist.get_world_size(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else "cpu"
    recs, ops = _local_block(matches, max_ops)
    n = recs.shape[0]
    rsz = _REC_DTYPE.itemsize
    width = rsz + max_ops
    key = (id(group), max_ops, str(device))
    while True:
        gb = _GATHER.get(key)
        if gb is None:
            gb = _GATHER[key] = _GatherBuffers(world, 256, width, device)
        cap = gb.cap
        block = gb.send_host.numpy()
        block[:8] = np.frombuffer(np.int64(n).tobytes(), dtype=np.uint8)
        take = min(n, cap)
        body = block[8:].reshape(cap, width)
        body[:take, :rsz] = recs[:take]
        body[:take, rsz:] = ops[:take]
        if gb.cuda:
            gb.send_dev.copy_(gb.send_host, non_blocking=True)
            dist.all_gather_into_tensor(gb.recv_dev, gb.send_dev, group=group)
            gb.recv_host.copy_(gb.recv_dev, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        else:
            dist.all_gather_into_tensor(gb.recv_host, gb.send_host, group=group)
        allb = gb.recv_host.numpy().reshape(world, gb.block)
        counts = allb[:, :8].copy().view(np.int64).reshape(world).tolist()
        if max(counts) <= cap:
            break
        _GATHER[key] = _GatherBuffers(world, 2 * max(counts), width, device)  # same decision on every rank
    out_recs = []
    out_ops = []
    off = 0
    for r in range(world):
        if counts[r] == 0:
            continue
        body = allb[r, 8:].reshape(cap, width)[:counts[r]]
        rr = np.ascontiguousarray(body[:, :rsz]).view(_REC_DTYPE).reshape(-1).copy()
        lens = rr["ops_len"].astype(np.int64)
        mask = np.arange(max_ops)[None, :] < lens[:, None]
        flat = body[:, rsz:][mask]
        rr["ops_off"] = off + np.concatenate(([0], np.cumsum(lens)[:-1]))
        off += int(lens.sum())
        out_recs.append(rr)
        out_ops.append(flat.tobytes())
    return MatchList(np.concatenate(out_recs) if out_recs else np.zeros(0, dtype=_REC_DTYPE), b"".join(out_ops))
