"""Python surface of the GPU search path, mirroring the reference's pyo3 module
(`sassy.Searcher`, `sassy.Match`; reference src/python.rs:26-220) plus the Rust-only
encoded-pattern API (src/search.rs:404-433).  All work is done by libsassy_b200.so."""
from __future__ import annotations

import collections.abc
import ctypes
import math
from typing import Iterable, List, Optional, Sequence, Union

import numpy as np

from . import _native

_REC_DTYPE = np.dtype([("pattern_idx", "<u8"), ("text_idx", "<u8"), ("text_start", "<u8"), ("text_end", "<u8"),
                       ("pattern_start", "<u8"), ("pattern_end", "<u8"), ("cost", "<i4"), ("strand", "u1"),
                       ("reserved", "u1", (3,)), ("ops_len", "<u4"), ("reserved2", "<u4"), ("ops_off", "<u8")])
assert _REC_DTYPE.itemsize == ctypes.sizeof(_native.GpuMatch) == 72


def _rle(ops: str) -> str:
    """pa_types::Cigar::to_string: `<cnt><op>` runs, count always printed (src/lib.rs:83,107)."""
    out = []
    i = 0
    n = len(ops)
    while i < n:
        j = i
        while j < n and ops[j] == ops[i]:
            j += 1
        out.append(f"{j - i}{ops[i]}")
        i = j
    return "".join(out)


class Match:
    """reference src/search.rs:35-62; getters as src/python.rs:155-220."""

    __slots__ = ("pattern_idx", "text_idx", "text_start", "text_end", "pattern_start", "pattern_end",
                 "cost", "strand", "_ops")

    def __init__(self, pattern_idx, text_idx, text_start, text_end, pattern_start, pattern_end, cost, strand, ops):
        self.pattern_idx = pattern_idx
        self.text_idx = text_idx
        self.text_start = text_start
        self.text_end = text_end
        self.pattern_start = pattern_start
        self.pattern_end = pattern_end
        self.cost = cost
        self.strand = strand  # "+" / "-"
        self._ops = ops

    @property
    def cigar(self) -> str:
        return _rle(self._ops)

    def _key(self):
        return (self.pattern_idx, self.text_idx, self.text_start, self.text_end, self.pattern_start,
                self.pattern_end, self.cost, self.strand, self._ops)

    def __eq__(self, other):
        return isinstance(other, Match) and self._key() == other._key()

    def __hash__(self):
        return hash(self._key())

    def __repr__(self):
        return (f"<Match pattern_start={self.pattern_start} text_start={self.text_start} "
                f"pattern_end={self.pattern_end} text_end={self.text_end} cost={self.cost} "
                f"strand='{self.strand}' cigar='{self.cigar}'>")


class MatchList(collections.abc.Sequence):
    """The Vec<Match> of one search: a lazy view over the native result records.

    Behaves like a list of Match (len, indexing, iteration, ==); Match objects are only
    materialised on access, so result sets with millions of matches cost one memcpy."""

    def __init__(self, recs: "np.ndarray", ops: bytes):
        self._recs = recs
        self._ops = ops
        self._list = None

    def __len__(self):
        return len(self._recs)

    def _make(self, r) -> Match:
        off = int(r["ops_off"])
        return Match(int(r["pattern_idx"]), int(r["text_idx"]), int(r["text_start"]), int(r["text_end"]),
                     int(r["pattern_start"]), int(r["pattern_end"]), int(r["cost"]), "-" if r["strand"] else "+",
                     self._ops[off:off + int(r["ops_len"])].decode())

    def _all(self) -> List[Match]:
        """Materialise every Match once (column-wise, much faster than record by record)."""
        if self._list is None:
            r = self._recs
            ops = self._ops.decode()
            cols = [r[f].tolist() for f in ("pattern_idx", "text_idx", "text_start", "text_end", "pattern_start",
                                            "pattern_end", "cost", "strand", "ops_off", "ops_len")]
            self._list = [Match(pi, ti, ts, te, ps, pe, c, "-" if st else "+", ops[oo:oo + ol])
                          for pi, ti, ts, te, ps, pe, c, st, oo, ol in zip(*cols)]
        return self._list

    def __getitem__(self, i):
        if isinstance(i, slice):
            return self._all()[i]
        if self._list is not None:
            return self._list[i]
        return self._make(self._recs[i])

    def __iter__(self):
        return iter(self._all())

    def __eq__(self, other):
        if isinstance(other, (list, MatchList)):
            return len(self) == len(other) and all(a == b for a, b in zip(self, other))
        return NotImplemented

    def __repr__(self):
        return f"MatchList({list(self)!r})"

    @property
    def records(self) -> "np.ndarray":
        """Structured array with the fields of sassy_gpu_Match (include/sassy_gpu.h)."""
        return self._recs


class DeviceText:
    """A text resident in HBM (device analogue of the reference's CachedRev, src/search.rs:144-166)."""

    def __init__(self, searcher: "Searcher", handle: int, n: int):
        self._searcher = searcher
        self._h = handle
        self.n = n

    def __len__(self):
        return self.n

    def free(self):
        if self._h:
            _native.load().sassy_gpu_text_free(self._searcher._h, self._h)
            self._h = None

    def __del__(self):
        try:
            if self._searcher._h:
                self.free()
        except Exception:
            pass


class EncodedPatterns:
    """reference src/pattern_tiling/general.rs:133-150."""

    def __init__(self, handle: int, n_patterns: int, m: int):
        self._h = handle
        self.n_patterns = n_patterns
        self.pattern_len = m

    def __del__(self):
        try:
            if self._h:
                _native.load().sassy_gpu_patterns_free(self._h)
                self._h = None
        except Exception:
            pass


def _bytes_data_offset() -> int:
    """Offset of the character data inside a CPython bytes object (0 if this is not CPython or the
    layout is not the expected one: then every buffer goes through ctypes)."""
    import platform
    if platform.python_implementation() != "CPython":
        return 0
    off = bytes.__basicsize__ - 1
    probe = b"sassy_b200 probe"
    addr = ctypes.cast(ctypes.c_char_p(probe), ctypes.c_void_p).value
    return off if addr == id(probe) + off else 0


_BYTES_DATA_OFFSET = _bytes_data_offset()


def _as_buffer(b):
    """Returns (address, length, keepalive) for bytes-like objects and (ptr, len) tuples."""
    if isinstance(b, tuple):  # (host address, length), e.g. pinned memory from host_alloc()
        return b[0], b[1], None
    if isinstance(b, bytes):
        return ctypes.cast(ctypes.c_char_p(b), ctypes.c_void_p).value or 0, len(b), b
    mv = memoryview(b).cast("B")
    if mv.readonly:
        data = bytes(mv)
        return ctypes.cast(ctypes.c_char_p(data), ctypes.c_void_p).value or 0, len(data), data
    arr = (ctypes.c_uint8 * len(mv)).from_buffer(mv)
    return ctypes.addressof(arr), len(mv), (arr, mv)


class Searcher:
    """`Searcher(alphabet, rc=True, alpha=None, max_n_frac=None)` as in src/python.rs:31-65.

    `device` selects the CUDA device (default: env SASSY_B200_DEVICE or 0).  `alpha` enables
    overhang (iupac only): pattern characters hanging over a text end cost alpha each."""

    def __init__(self, alphabet: str, rc: bool = True, alpha: Optional[float] = None,
                 max_n_frac: Optional[float] = None, device: Optional[int] = None):
        import os
        self._h = None
        self._lib = _native.load()
        a = alphabet.lower()
        if a not in ("ascii", "dna", "iupac"):
            raise ValueError(f"Unsupported alphabet: {alphabet}")  # src/python.rs:52-57
        if a == "ascii":
            rc = False  # src/python.rs:40-42
        if alpha is not None and a != "iupac":
            raise ValueError("Overhang is only supported for the iupac alphabet")  # src/search.rs:373-379
        if alpha is not None and not (0.0 <= alpha <= 1.0):
            raise ValueError("Alpha must be in range 0.0 <= alpha <= 1.0")
        if device is None:
            device = int(os.environ.get("SASSY_B200_DEVICE", "0"))
        self.alphabet = a
        self.rc = bool(rc)
        self.device = device
        h = self._lib.sassy_gpu_searcher(a.encode(), self.rc, math.nan if alpha is None else float(alpha), device)
        if not h:
            raise RuntimeError(_native.last_error())
        self._h = h
        if max_n_frac is not None:
            self.set_max_n_frac(max_n_frac)

    # -- Searcher options (reference src/search.rs:441-483) ----------------
    def set_max_n_frac(self, max_n_frac: float):
        """Searcher::set_max_n_frac: drop matches whose text holds more than this fraction of N
        (1.0 disables the filter)."""
        if self._lib.sassy_gpu_set_max_n_frac(self._h, float(max_n_frac)) != 0:
            raise ValueError(max_n_frac)
        return self

    def with_max_n_frac(self, max_n_frac: float):
        return self.set_max_n_frac(max_n_frac)

    def without_max_n_frac(self):
        return self.set_max_n_frac(1.0)

    def with_max_overhang(self, max_overhang: Optional[int]):
        """Searcher::with_max_overhang (src/search.rs:436-439)."""
        self._lib.sassy_gpu_set_max_overhang(self._h, -1 if max_overhang is None else int(max_overhang))
        return self

    def set_trace(self, trace: bool):
        """Searcher::set_trace / with_trace / without_trace: without trace a match carries the end
        position and cost only (the unknown fields hold USIZE_MAX, like the reference)."""
        self._lib.sassy_gpu_set_trace(self._h, int(bool(trace)))
        return self

    def without_trace(self):
        return self.set_trace(False)

    def with_trace(self):
        return self.set_trace(True)

    def only_best_match(self, on: bool = True):
        """Searcher::only_best_match: the rightmost match of minimal cost per (pattern, text, strand)."""
        self._lib.sassy_gpu_set_only_best_match(self._h, int(bool(on)))
        return self

    def close(self):
        if self._h:
            self._lib.sassy_searcher_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers ---------------------------------------------------------
    def _collect(self, res) -> "MatchList":
        if not res:
            msg = _native.last_error()
            if "IUPAC" in msg or "pattern" in msg.lower():
                raise ValueError(msg)
            raise RuntimeError(msg)
        lib = self._lib
        try:
            n = lib.sassy_gpu_result_len(res)
            if n == 0:
                return MatchList(np.zeros(0, dtype=_REC_DTYPE), b"")
            ms = lib.sassy_gpu_result_matches(res)
            # (c_char * size).from_address: sizes beyond 2 GiB are fine (ctypes.string_at takes a C int)
            addr = ctypes.cast(ms, ctypes.c_void_p).value
            recs = np.frombuffer(bytes((ctypes.c_char * (n * _REC_DTYPE.itemsize)).from_address(addr)), dtype=_REC_DTYPE)
            last = recs[-1]
            ops_ptr = lib.sassy_gpu_result_ops(res)
            total = int(last["ops_off"]) + int(last["ops_len"])
            ops = bytes((ctypes.c_char * total).from_address(ops_ptr)) if ops_ptr and total else b""
            return MatchList(recs, ops)
        finally:
            lib.sassy_gpu_result_free(res)

    def set_variant(self, variant: str):
        """'tma' (default) or 'ldg' -- scan kernel data path (A/B)."""
        if self._lib.sassy_gpu_set_variant(self._h, {"tma": 0, "ldg": 1}[variant]) != 0:
            raise ValueError(variant)

    def set_filter(self, mode: str):
        """'off', 'auto' (default) or 'force' -- the exact piece prefilter in front of the scan."""
        if self._lib.sassy_gpu_set_filter(self._h, {"off": 0, "auto": 1, "force": 2}[mode]) != 0:
            raise ValueError(mode)

    def set_transport(self, mode: str):
        """'packed' (default: large Dna host texts cross PCIe at 2 bits per character) or 'bytes'."""
        if self._lib.sassy_gpu_set_transport(self._h, {"bytes": 0, "packed": 1}[mode]) != 0:
            raise ValueError(mode)

    def stats(self) -> dict:
        st = _native.GpuStats()
        self._lib.sassy_gpu_stats(self._h, ctypes.byref(st))
        return {f: getattr(st, f) for f, _ in st._fields_ if not f.startswith("reserved")}

    def upload_text(self, text) -> DeviceText:
        addr, n, keep = _as_buffer(text)
        h = self._lib.sassy_gpu_text_upload(self._h, addr, n)
        if not h:
            raise RuntimeError(_native.last_error())
        return DeviceText(self, h, n)

    def text_from_device(self, device_ptr: int, n: int) -> DeviceText:
        h = self._lib.sassy_gpu_text_from_device(self._h, device_ptr, n)
        if not h:
            raise RuntimeError(_native.last_error())
        return DeviceText(self, h, n)

    # -- reference API ---------------------------------------------------
    def _search(self, pattern, text, k: int, all_minima: bool) -> List[Match]:
        paddr, plen, pkeep = _as_buffer(pattern)
        if isinstance(text, DeviceText):
            res = self._lib.sassy_gpu_search_text(self._h, paddr, plen, text._h, k, int(all_minima))
        else:
            taddr, tlen, tkeep = _as_buffer(text)
            res = self._lib.sassy_gpu_search(self._h, paddr, plen, taddr, tlen, k, int(all_minima))
        return self._collect(res)

    def search(self, pattern, text, k: int) -> List[Match]:
        """Searcher::search (src/search.rs:510-525; src/python.rs:67-81)."""
        return self._search(pattern, text, k, False)

    def search_all(self, pattern, text, k: int) -> List[Match]:
        """Searcher::search_all (src/search.rs:685-700; src/python.rs:139-152)."""
        return self._search(pattern, text, k, True)

    def search_with_pam(self, pattern, text, k: int, pam, all_minima: bool = True) -> List[Match]:
        """Searcher::search_with_fn (src/search.rs:767-784) with the end filter of the reference's
        CRISPR mode (bin/crispr.rs:198-221): an end position is kept only if the len(pam) text
        characters before it match `pam` exactly (complemented PAM on the reverse strand)."""
        paddr, plen, pkeep = _as_buffer(pattern)
        maddr, mlen, mkeep = _as_buffer(pam)
        if isinstance(text, DeviceText):
            res = self._lib.sassy_gpu_search_pam_text(self._h, paddr, plen, text._h, k, int(all_minima), maddr, mlen)
        else:
            taddr, tlen, tkeep = _as_buffer(text)
            res = self._lib.sassy_gpu_search_pam(self._h, paddr, plen, taddr, tlen, k, int(all_minima), maddr, mlen)
        return self._collect(res)

    @staticmethod
    def _ptr_array(items):
        """(void* array, size_t array, keepalives) for a sequence of bytes-like objects."""
        n = len(items)
        if n > 64 and _BYTES_DATA_OFFSET and all(type(x) is bytes for x in items):
            # many texts (reads): the buffer of a CPython bytes object lies at a fixed offset behind
            # id(obj), so the pointer array is one vectorised add instead of n ctypes casts (the
            # library copies the texts into its pinned staging buffer with several threads)
            lens_np = np.fromiter(map(len, items), dtype=np.uint64, count=n)
            ptrs_np = np.fromiter(map(id, items), dtype=np.uint64, count=n) + np.uint64(_BYTES_DATA_OFFSET)
            return (ctypes.c_void_p(ptrs_np.ctypes.data), ctypes.c_void_p(lens_np.ctypes.data),
                    (items, ptrs_np, lens_np))
        bufs = [_as_buffer(x) for x in items]
        ptrs = (ctypes.c_void_p * max(1, len(bufs)))(*[b[0] for b in bufs])
        lens = (ctypes.c_size_t * max(1, len(bufs)))(*[b[1] for b in bufs])
        return ptrs, lens, bufs

    def search_patterns(self, patterns: Sequence, text, k: int) -> List[Match]:
        """Searcher::search_patterns (src/search.rs:648-683): equal-length patterns, one text."""
        if not patterns:
            return MatchList(np.zeros(0, dtype=_REC_DTYPE), b"")
        m = len(patterns[0])
        if any(len(p) != m for p in patterns):
            raise ValueError("All patterns passed to search_patterns must have the same length")
        ptrs, lens, keep = self._ptr_array(patterns)
        taddr, tlen, tkeep = _as_buffer(text)
        return self._collect(self._lib.sassy_gpu_search_patterns(self._h, ptrs, len(patterns), m, taddr, tlen, k))

    def search_texts(self, pattern, texts: Sequence, k: int) -> List[Match]:
        """Searcher::search_texts (src/search.rs:615-640): one pattern, many (short) texts."""
        paddr, plen, pkeep = _as_buffer(pattern)
        ptrs, lens, keep = self._ptr_array(texts)
        return self._collect(self._lib.sassy_gpu_search_texts(self._h, paddr, plen, ptrs, lens, len(texts), k))

    def search_many(self, patterns: Sequence, texts: Sequence, k: int, threads: int = 0,
                    mode: str = "single") -> List[Match]:
        """Searcher::search_many (src/search.rs:531-603; src/python.rs:83-116): every pattern
        against every text, pattern_idx / text_idx filled in.  The reference's three modes return
        the same set of matches; here short texts are searched by one kernel launch per pattern
        length (one thread per (text, pattern, strand)) and long texts by the row-tiled scan.
        `threads` is accepted for signature compatibility."""
        if mode not in ("single", "batch_patterns", "batch_texts"):
            raise ValueError("Unsupported search mode. Must be one of 'single', 'batch_patterns', or 'batch_texts'")
        pp, pl, pk = self._ptr_array(patterns)
        tp, tl, tk = self._ptr_array(texts)
        return self._collect(self._lib.sassy_gpu_search_many(self._h, pp, pl, len(patterns), tp, tl, len(texts), k))

    def encode_patterns(self, patterns: Sequence[bytes]) -> EncodedPatterns:
        """Searcher::encode_patterns (src/search.rs:404-413): equal-length patterns."""
        if not patterns:
            raise ValueError("no patterns")
        m = len(patterns[0])
        if any(len(p) != m for p in patterns):
            raise ValueError("all patterns must have the same length")
        blob = b"".join(bytes(p) for p in patterns)
        h = self._lib.sassy_gpu_encode_patterns(self._h, ctypes.cast(ctypes.c_char_p(blob), ctypes.c_void_p),
                                                len(patterns), m)
        if not h:
            raise ValueError(_native.last_error())
        return EncodedPatterns(h, len(patterns), m)

    def _search_encoded(self, enc: EncodedPatterns, text, k: int, all_minima: bool) -> List[Match]:
        if isinstance(text, DeviceText):
            res = self._lib.sassy_gpu_search_encoded(self._h, enc._h, text._h, k, int(all_minima))
        else:
            taddr, tlen, tkeep = _as_buffer(text)
            res = self._lib.sassy_gpu_search_encoded_host(self._h, enc._h, taddr, tlen, k, int(all_minima))
        return self._collect(res)

    def search_encoded_patterns(self, enc: EncodedPatterns, text, k: int) -> List[Match]:
        """Searcher::search_encoded_patterns (src/search.rs:415-423)."""
        return self._search_encoded(enc, text, k, False)

    def search_all_encoded_patterns(self, enc: EncodedPatterns, text, k: int) -> List[Match]:
        """Searcher::search_all_encoded_patterns (src/search.rs:426-433)."""
        return self._search_encoded(enc, text, k, True)


def host_alloc(nbytes: int) -> int:
    """Pinned host memory (address) for texts passed as (address, length) tuples."""
    p = _native.load().sassy_gpu_host_alloc(nbytes)
    if not p:
        raise MemoryError(_native.last_error())
    return p


def host_free(addr: int):
    _native.load().sassy_gpu_host_free(addr)


def device_count() -> int:
    return _native.load().sassy_gpu_device_count()
