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


def _local_block(matches):
    """(records as uint8 [n*72], raw CIGAR op bytes as uint8) of a MatchList or a list of Match;
    ops_off of the records is relative to the returned op bytes."""
    from .searcher import MatchList, _REC_DTYPE
    if isinstance(matches, MatchList):
        recs = matches.records
        raw = np.frombuffer(matches._ops, dtype=np.uint8)
    else:
        n = len(matches)
        recs = np.zeros(n, dtype=_REC_DTYPE)
        chunks = []
        off = 0
        for i, m in enumerate(matches):
            recs[i] = (m.pattern_idx, m.text_idx, m.text_start, m.text_end, m.pattern_start, m.pattern_end, m.cost,
                       1 if m.strand == "-" else 0, (0, 0, 0), len(m._ops), 0, off)
            chunks.append(m._ops.encode())
            off += len(m._ops)
        raw = np.frombuffer(b"".join(chunks), dtype=np.uint8)
    return np.ascontiguousarray(recs).view(np.uint8).reshape(-1), raw


class _GatherBuffers:
    """Re-used staging for gather_matches: pinned host blocks and device blocks of one capacity.
    Block layout: [n_records:int64][n_op_bytes:int64][records: cap*72][op bytes: cap*max_ops]."""

    def __init__(self, world: int, cap: int, max_ops: int, rsz: int, device):
        self.cap = cap
        self.rec_bytes = cap * rsz
        self.ops_bytes = cap * max_ops
        self.block = 16 + self.rec_bytes + self.ops_bytes
        cuda = torch.device(device).type == "cuda"
        self.send_host = torch.zeros(self.block, dtype=torch.uint8, pin_memory=cuda)
        self.recv_host = torch.zeros(world * self.block, dtype=torch.uint8, pin_memory=cuda)
        if cuda:
            self.send_dev = torch.zeros(self.block, dtype=torch.uint8, device=device)
            self.recv_dev = torch.zeros(world * self.block, dtype=torch.uint8, device=device)
        else:
            self.send_dev, self.recv_dev = self.send_host, self.recv_host
        self.cuda = cuda
        self.send_np = self.send_host.numpy()
        self.recv_np = self.recv_host.numpy().reshape(world, self.block)
        self.head = self.send_np[:16].view(np.int64)


_GATHER = {}


def tag_rank(matches, rank: int):
    """Copy of a MatchList with text_idx = rank (what the fused PeerGather reports as source)."""
    from .searcher import MatchList
    recs = matches.records.copy()
    recs["text_idx"] = rank
    return MatchList(recs, matches._ops)


def gather_matches(matches, max_ops: int, device=None, group=None):
    """All ranks receive the concatenation (rank order) of every rank's matches (a MatchList).

    ONE collective in the common case: every rank contributes a fixed-capacity block
    [counts | records | CIGAR op bytes]; only if some rank holds more than the capacity is
    the exchange repeated with a larger block (NCCL has no gatherv)."""
    from .searcher import MatchList, _REC_DTYPE
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return matches
    world = dist.get_world_size(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else "cpu"
    rsz = _REC_DTYPE.itemsize
    recs, raw = _local_block(matches)
    n = recs.size // rsz
    key = (id(group), max_ops, str(device))
    while True:
        gb = _GATHER.get(key)
        if gb is None:
            gb = _GATHER[key] = _GatherBuffers(world, 256, max_ops, rsz, device)
        fits = n <= gb.cap and raw.size <= gb.ops_bytes
        gb.head[0] = n
        gb.head[1] = raw.size
        if fits:
            gb.send_np[16:16 + recs.size] = recs
            gb.send_np[16 + gb.rec_bytes:16 + gb.rec_bytes + raw.size] = raw
        if gb.cuda:
            gb.send_dev.copy_(gb.send_host, non_blocking=True)
            dist.all_gather_into_tensor(gb.recv_dev, gb.send_dev, group=group)
            gb.recv_host.copy_(gb.recv_dev, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        else:
            dist.all_gather_into_tensor(gb.recv_host, gb.send_host, group=group)
        heads = gb.recv_np[:, :16].copy().view(np.int64).reshape(world, 2)
        need = int(max(heads[:, 0].max(), -(-int(heads[:, 1].max()) // max(max_ops, 1))))
        if need <= gb.cap:
            break
        _GATHER[key] = _GatherBuffers(world, 2 * need, max_ops, rsz, device)  # same decision on every rank
    out_recs = []
    out_ops = []
    off = 0
    for r in range(world):
        cnt, nops = int(heads[r, 0]), int(heads[r, 1])
        if cnt == 0:
            continue
        rr = gb.recv_np[r, 16:16 + cnt * rsz].copy().view(_REC_DTYPE)
        rr["ops_off"] += off
        off += nops
        out_recs.append(rr)
        out_ops.append(gb.recv_np[r, 16 + gb.rec_bytes:16 + gb.rec_bytes + nops].tobytes())
    return MatchList(np.concatenate(out_recs) if out_recs else np.zeros(0, dtype=_REC_DTYPE), b"".join(out_ops))


class PeerGather:
    """Search + gather of every rank's matches through ONE fused exchange over peer memory
    (csrc/peer_gather.cu): the traceback kernel leaves the records in this rank's slot of a
    receive buffer, a kernel stores them into every peer's buffer over NVLink and releases a
    step flag, a second kernel acquires all flags and moves the records to pinned host memory.
    torch.distributed is used once, to exchange the 64-byte CUDA IPC handles, and for the
    fall-back all-gather when some rank's result does not fit the exchange.

    All ranks must call search()/search_encoded() in lock step.  Results: all ranks' matches in
    rank order, text_idx = source rank, pattern_idx local to the source rank's pattern set.

    pipelined=True (search_sharded / flush_sharded only): a call pushes its records and collects
    those of the PREVIOUS call, which the peers pushed a whole search earlier, so no rank waits for
    the slowest one inside a step; calls return the previous call's result (None at first),
    flush_sharded() the last one.  Results must fit the slots (cap records per rank)."""

    def __init__(self, searcher, max_ops: int, cap: int = 2048, group=None, pipelined: bool = False):
        import ctypes
        self.pipelined = bool(pipelined)
        from . import _native
        self._lib = _native.load()
        self._searcher = searcher
        self._group = group
        self.max_ops = max_ops
        if dist.is_initialized():
            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        else:
            self.world, self.rank = 1, 0
        self.fallbacks = 0
        self._h = None
        # Set-up is collective: no rank raises between two collectives, the outcome is agreed on
        # by an all-reduce (which doubles as the barrier before the first exchange).
        err = None
        buf = (ctypes.c_uint8 * 64)()
        self._h = self._lib.sassy_gpu_gather_create(searcher._h, self.world, self.rank, cap, max_ops)
        if not self._h:
            err = _native.last_error()
        elif self.world > 1 and self._lib.sassy_gpu_gather_handle(self._h, buf) != 0:
            err = _native.last_error()
        if self.world > 1:
            dev = torch.device("cuda", searcher.device) if dist.get_backend(group) == "nccl" else "cpu"
            mine = torch.tensor(list(buf), dtype=torch.uint8, device=dev)
            allh = torch.empty(self.world * 64, dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(allh, mine, group=group)
            if err is None:
                blob = bytes(allh.cpu().numpy().tobytes())
                if self._lib.sassy_gpu_gather_connect(self._h, blob) != 0:
                    err = _native.last_error()
            ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0 and err is None:
                err = "peer memory could not be mapped on another rank"
        if err is not None:
            self._free_local()
            raise RuntimeError(err)
        if self.pipelined:
            self._lib.sassy_gpu_gather_set_pipelined(self._h, 1)

    @classmethod
    def create_or_none(cls, searcher, max_ops: int, device=None, group=None, pipelined: bool = False):
        """Collective constructor: every rank gets a PeerGather, or -- if peer memory cannot be
        mapped on some rank (no P2P path between two GPUs) -- every rank gets None and the caller
        uses gather_matches (NCCL) instead."""
        try:
            return cls(searcher, max_ops, group=group, pipelined=pipelined)  # raises on every rank or on none
        except RuntimeError as e:
            rank = dist.get_rank(group) if dist.is_initialized() else 0
            import sys
            print(f"[sassy_b200] peer-memory gather unavailable (rank {rank}): {e}", file=sys.stderr, flush=True)
            return None

    def _free_local(self):
        if self._h:
            self._lib.sassy_gpu_gather_free(self._h)
            self._h = None

    def close(self):
        if self._h:
            if self.world > 1 and dist.is_initialized():
                dist.barrier(group=self._group)  # no peer may still be writing into our buffer
            self._lib.sassy_gpu_gather_free(self._h)
            self._h = None

    def __del__(self):
        try:
            if self._h:
                self._lib.sassy_gpu_gather_free(self._h)
                self._h = None
        except Exception:
            pass

    def _finish(self, res, complete: int):
        from .searcher import MatchList
        ms = self._searcher._collect(res)
        if complete or self.world == 1:
            return ms
        # some rank's result did not fit the fused exchange: plain all-gather of the local lists
        self.fallbacks += 1
        return gather_matches(tag_rank(ms, self.rank), self.max_ops, group=self._group)

    def search(self, pattern: bytes, text, k: int, all_minima: bool = False):
        import ctypes
        from .searcher import _as_buffer
        paddr, plen, keep = _as_buffer(pattern)
        ok = ctypes.c_int(0)
        res = self._lib.sassy_gpu_search_text_gathered(self._searcher._h, self._h, paddr, plen, text._h, k,
                                                       int(all_minima), ctypes.byref(ok))
        return self._finish(res, ok.value)

    def search_sharded(self, pattern: bytes, window_text, k: int, layout, n_global: int, all_minima: bool = False):
        """One text cut into slabs (slab_layout): this rank's window -> the matches of the unsharded
        search (sassy_gpu_search_text_sharded); NCCL all-gather + merge_slabs when the records of
        some rank do not fit the fused exchange."""
        import ctypes
        from .searcher import _as_buffer
        paddr, plen, keep = _as_buffer(pattern)
        slabs = self._slabs(layout)
        ok = ctypes.c_int(0)
        res = self._lib.sassy_gpu_search_text_sharded(self._searcher._h, self._h, paddr, plen, window_text._h, k,
                                                      int(all_minima), slabs.ctypes.data, len(layout), n_global,
                                                      ctypes.byref(ok))
        ms = self._searcher._collect(res)
        if ok.value == 2:  # pipelined mode, first call: nothing collected yet
            return None
        if ok.value:
            return ms
        self.fallbacks += 1
        allm = gather_matches(tag_rank(ms, self.rank), self.max_ops, group=self._group)
        return merge_slabs(allm, layout, n_global, all_minima)

    def _slabs(self, layout):
        """sassy_gpu_Slab array of a slab_layout() (cached: the layout of a sharded text does not change)."""
        key = id(layout)
        cached = getattr(self, "_slab_cache", None)
        if cached is None or cached[0] != key or cached[1] != len(layout):
            slabs = np.zeros((len(layout), 3), dtype=np.uint64)
            for r, (wlo, _whi, lo, hi) in enumerate(layout):
                slabs[r] = (wlo, lo, hi)
            self._slab_cache = cached = (key, len(layout), slabs, layout)
        return cached[2]

    def flush_sharded(self, pattern_len: int, layout, n_global: int, all_minima: bool = False):
        """Pipelined mode: the result of the last search_sharded call (None if nothing is pending)."""
        import ctypes
        slabs = self._slabs(layout)
        state = ctypes.c_int(0)
        res = self._lib.sassy_gpu_text_sharded_flush(self._searcher._h, self._h, pattern_len, int(all_minima),
                                                     slabs.ctypes.data, len(layout), n_global, ctypes.byref(state))
        ms = self._searcher._collect(res)
        return None if state.value == 2 else ms

    def search_encoded(self, enc, text, k: int, all_minima: bool = False):
        import ctypes
        ok = ctypes.c_int(0)
        res = self._lib.sassy_gpu_search_encoded_gathered(self._searcher._h, self._h, enc._h, text._h, k,
                                                          int(all_minima), ctypes.byref(ok))
        return self._finish(res, ok.value)


# ---------------------------------------------------------------------------------------------
# One text cut into slabs (SURVEY 8e "Partitioning"; the reference's lane chunking with overlap,
# src/search.rs:1018-1049,1202-1240, at GPU granularity).

def slab_layout(n_global: int, world: int, m: int, k: int):
    """[(window_lo, window_hi, own_lo, own_hi)] per rank: slab r owns [own_lo, own_hi), its search
    window adds an (m + k) halo on both sides (left: warm-up of the forward scan and the traceback
    window; right: the same for the reverse-complement strand, which scans right to left)."""
    halo = m + k
    per = -(-n_global // world)
    out = []
    for r in range(world):
        lo, hi = min(n_global, r * per), min(n_global, (r + 1) * per)
        out.append((max(0, lo - halo), min(n_global, hi + halo), lo, hi))
    return out


def merge_slabs(matches, layout, n_global: int, all_minima: bool = False):
    """Gathered per-slab search_all records (text_idx = source rank, window coordinates) -> the
    matches of the unsharded search: ownership filter, global coordinates, local-minima rule on the
    merged list (sassy_gpu_merge_slabs, host code in csrc/shard_merge.h)."""
    import ctypes
    from . import _native
    from .searcher import MatchList, _REC_DTYPE
    lib = _native.load()
    recs, raw = _local_block(matches)
    n = recs.size // _REC_DTYPE.itemsize
    slabs = np.zeros((len(layout), 3), dtype=np.uint64)
    for r, (wlo, _whi, lo, hi) in enumerate(layout):
        slabs[r] = (wlo, lo, hi)
    recs = np.ascontiguousarray(recs)
    raw = np.ascontiguousarray(raw)
    res = lib.sassy_gpu_merge_slabs(recs.ctypes.data if n else None, n, raw.ctypes.data if raw.size else None,
                                    slabs.ctypes.data, len(layout), n_global, int(all_minima))
    if not res:
        raise RuntimeError(_native.last_error())
    try:
        cnt = lib.sassy_gpu_result_len(res)
        if cnt == 0:
            return MatchList(np.zeros(0, dtype=_REC_DTYPE), b"")
        addr = ctypes.cast(lib.sassy_gpu_result_matches(res), ctypes.c_void_p).value
        out = np.frombuffer(bytes((ctypes.c_char * (cnt * _REC_DTYPE.itemsize)).from_address(addr)), dtype=_REC_DTYPE)
        total = int(out[-1]["ops_off"]) + int(out[-1]["ops_len"])
        ops_ptr = lib.sassy_gpu_result_ops(res)
        ops = bytes((ctypes.c_char * total).from_address(ops_ptr)) if ops_ptr and total else b""
        return MatchList(out, ops)
    finally:
        lib.sassy_gpu_result_free(res)


def search_text_sharded(searcher, pattern: bytes, window_text, k: int, n_global: int, all_minima: bool = False,
                        peer_gather: "PeerGather" = None, group=None, layout=None):
    """Searcher::search / search_all of ONE text that is cut into `world` slabs, one per rank.

    `window_text` holds this rank's window slab_layout(n_global, world, m, k)[rank] (a DeviceText,
    bytes or a (address, length) tuple).  Every rank searches its window with search_all, the
    records are exchanged (fused peer-memory gather when `peer_gather` is given, else one
    all-gather) and merged on every rank; the result equals the unsharded search of the whole
    text, including runs of minima that cross a slab border.  Collective: call on all ranks."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    m = len(pattern)
    if layout is None:
        layout = slab_layout(n_global, world, m, k)
    if peer_gather is not None:
        # search_all + fused gather + merge inside the library (one call, one host synchronisation)
        return peer_gather.search_sharded(pattern, window_text, k, layout, n_global, all_minima)
    else:
        ms = searcher.search_all(pattern, window_text, k)
        if world > 1:
            ms = gather_matches(tag_rank(ms, rank), max_ops=m + k + 1, group=group)
        else:
            ms = tag_rank(ms, 0)
    return merge_slabs(ms, layout, n_global, all_minima)
