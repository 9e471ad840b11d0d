"""Test-only backend: drives libsassy_b200_emu.so, the CPU emulation of the per-thread CUDA
logic (sassy_b200/csrc/emu.cpp), and applies the same strand/coordinate mapping as
sassy_b200/csrc/searcher.cu.  Lets the CPU-only suite check tiling / warm-up / minima /
traceback logic against the oracle without a GPU."""
from __future__ import annotations

import ctypes
from typing import List, Sequence

from oracle import Match, rle, complement, reverse_complement, OracleError
import oracle as _oracle
from sassy_b200 import build as _build

PROFILE = {"dna": 0, "iupac": 1, "ascii": 2}
OPS = "=XID"


class _GpuMatch(ctypes.Structure):
    _fields_ = [
        ("text_start", ctypes.c_uint64),
        ("text_end", ctypes.c_uint64),
        ("qs", ctypes.c_uint32),
        ("cost", ctypes.c_int32),
        ("nops", ctypes.c_uint32),
        ("failed", ctypes.c_uint32),
    ]


class _EmuOpts(ctypes.Structure):
    _fields_ = [("without_trace", ctypes.c_int), ("only_best", ctypes.c_int), ("n_endpoint", ctypes.c_int),
                ("max_n_frac", ctypes.c_float), ("pam", ctypes.c_char_p), ("pam_len", ctypes.c_int),
                ("alpha", ctypes.c_float), ("max_overhang", ctypes.c_int)]


USIZE_MAX = 2**64 - 1
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(_build.build_emu())
        lib.emu_search.restype = ctypes.c_void_p
        lib.emu_search.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_uint32, ctypes.c_int,
                                   ctypes.c_char_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_uint32, ctypes.c_int, ctypes.c_int]
        lib.emu_search_opts.restype = ctypes.c_void_p
        lib.emu_search_opts.argtypes = lib.emu_search.argtypes + [ctypes.POINTER(_EmuOpts)]
        lib.emu_len.restype = ctypes.c_size_t
        lib.emu_len.argtypes = [ctypes.c_void_p]
        lib.emu_matches.restype = ctypes.POINTER(_GpuMatch)
        lib.emu_matches.argtypes = [ctypes.c_void_p]
        lib.emu_ops.restype = ctypes.POINTER(ctypes.c_uint32)
        lib.emu_ops.argtypes = [ctypes.c_void_p]
        lib.emu_ops_words.restype = ctypes.c_uint32
        lib.emu_ops_words.argtypes = [ctypes.c_void_p]
        lib.emu_candidates.restype = ctypes.c_uint64
        lib.emu_candidates.argtypes = [ctypes.c_void_p]
        lib.emu_ltot.restype = ctypes.c_uint32
        lib.emu_ltot.argtypes = [ctypes.c_void_p]
        lib.emu_rows.restype = ctypes.c_uint32
        lib.emu_rows.argtypes = [ctypes.c_void_p]
        lib.emu_hits.restype = ctypes.c_uint64
        lib.emu_hits.argtypes = [ctypes.c_void_p]
        lib.emu_filter_words.restype = ctypes.c_int
        lib.emu_filter_words.argtypes = [ctypes.c_void_p]
        lib.emu_filter_len.restype = ctypes.c_int
        lib.emu_filter_len.argtypes = [ctypes.c_void_p]
        lib.emu_dense_tiles.restype = ctypes.c_uint32
        lib.emu_dense_tiles.argtypes = [ctypes.c_void_p]
        lib.emu_free.argtypes = [ctypes.c_void_p]
        _LIB = lib
    return _LIB


class EmuBackend:
    def __init__(self, ltot: int = 0, bpw: int = 444, use_filter: int = 0):
        self.ltot = ltot
        self.bpw = bpw
        self.use_filter = use_filter  # 0 full scan, 1 prefilter whenever possible, -1 engine's rule
        self.last_geom = None
        self.last_filter = None

    def _run(self, alphabet, queries: Sequence[bytes], rev: Sequence[int], text: bytes, k: int, all_minima: bool,
             pos0: bool, opts=None, raw=False):
        lib = _lib()
        m = len(queries[0])
        if opts is None:
            r = lib.emu_search(PROFILE[alphabet.lower()], b"".join(queries), bytes(rev), len(queries), m, text,
                               len(text), k, int(all_minima), int(pos0), self.ltot, self.bpw, self.use_filter)
        else:
            r = lib.emu_search_opts(PROFILE[alphabet.lower()], b"".join(queries), bytes(rev), len(queries), m, text,
                                    len(text), k, int(all_minima), int(pos0), self.ltot, self.bpw, self.use_filter,
                                    ctypes.byref(opts))
        assert r, "emu_search failed"
        try:
            n = lib.emu_len(r)
            ms = lib.emu_matches(r)
            ops = lib.emu_ops(r)
            ow = lib.emu_ops_words(r)
            self.last_geom = (lib.emu_ltot(r), lib.emu_rows(r))
            self.last_filter = (lib.emu_filter_words(r), lib.emu_filter_len(r), lib.emu_hits(r))
            self.last_dense_tiles = lib.emu_dense_tiles(r)
            out = []
            for i in range(n):
                g = ms[i]
                if g.failed & 1:
                    raise OracleError("trace failed")
                s = "".join(OPS[(ops[i * ow + (a >> 4)] >> ((a & 15) * 2)) & 3] for a in range(g.nops))
                if raw:
                    out.append((g.qs, g.text_start, g.text_end, g.cost, s, (g.failed >> 8) & 0xFFF, g.failed >> 20))
                else:
                    out.append((g.qs, g.text_start, g.text_end, g.cost, s))
            return out
        finally:
            lib.emu_free(r)

    def search(self, alphabet, pattern: bytes, text: bytes, k: int, rc: bool = False, all_minima: bool = False):
        if alphabet.lower() == "iupac" and not _oracle._lib().oracle_iupac_valid(pattern, len(pattern)):
            raise OracleError("Pattern is not valid IUPAC")
        queries = [pattern]
        rev = [0]
        if rc:
            queries.append(complement(alphabet, pattern))
            rev.append(1)
        n = len(text)
        res = []
        for qs, ts, te, cost, ops in self._run(alphabet, queries, rev, text, k, all_minima, True):
            if qs == 0:
                res.append(Match(0, ts, te, 0, len(pattern), cost, "+", rle(ops)))
            else:
                res.append(Match(0, n - te, n - ts, 0, len(pattern), cost, "-", rle(ops)))
        return res

    def search_opts(self, alphabet, pattern: bytes, text: bytes, k: int, rc: bool = False, all_minima: bool = False,
                    without_trace: bool = False, only_best: bool = False, max_n_frac=None, pam=None, alpha=None,
                    max_overhang=None):
        """v1 search under the Searcher options, with the mapping of Searcher::convert_v1 (searcher.cu)."""
        queries, rev = [pattern], [0]
        if rc:
            queries.append(complement(alphabet, pattern))
            rev.append(1)
        o = _EmuOpts(int(without_trace), int(only_best), 1, -1.0 if max_n_frac is None else float(max_n_frac),
                     pam, len(pam) if pam else 0, -1.0 if alpha is None else float(alpha),
                     -1 if max_overhang is None else int(max_overhang))
        n, m = len(text), len(pattern)
        res = []
        for qs, ts, te, cost, ops, ps, over in self._run(alphabet, queries, rev, text, k, all_minima, True, o, raw=True):
            pstart = USIZE_MAX if without_trace else ps
            if qs == 0:
                res.append(Match(0, ts, te, pstart, m - over, cost, "+", rle(ops)))
            else:
                res.append(Match(0, n - te, USIZE_MAX if without_trace else n - ts, pstart, m - over, cost, "-", rle(ops)))
        return res

    def search_encoded(self, alphabet, patterns: Sequence[bytes], text: bytes, k: int, rc: bool = False,
                       all_minima: bool = False, alpha=None, max_overhang=None, max_n_frac=None):
        P = len(patterns)
        queries = list(patterns)
        if rc:
            queries += [reverse_complement("iupac", p) for p in patterns]
        res = []
        if alpha is not None or max_n_frac is not None:  # Searcher::search_encoded_raw + convert_v2
            o = _EmuOpts(0, 0, 1 if alpha is not None else 0, -1.0 if max_n_frac is None else float(max_n_frac), None, 0,
                         -1.0 if alpha is None else float(alpha), -1 if max_overhang is None else int(max_overhang))
            m = len(patterns[0])
            for qs, ts, te, cost, ops, ps, over in self._run(alphabet, queries, [0] * len(queries), text, k,
                                                             all_minima, False, o, raw=True):
                res.append(Match(qs % P, ts, te, ps, m - over, cost, "-" if qs >= P else "+", rle(ops)))
            return res
        for qs, ts, te, cost, ops in self._run(alphabet, queries, [0] * len(queries), text, k, all_minima, False):
            res.append(Match(qs % P, ts, te, 0, len(patterns[0]), cost, "-" if qs >= P else "+", rle(ops)))
        return res
