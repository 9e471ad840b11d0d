"""CPU oracle for the sassy hot path -- TEST INFRASTRUCTURE ONLY.

ctypes wrapper over ``oracle/sassy_oracle.c`` (a scalar restatement of the
reference's Searcher::search / search_all / search_encoded_patterns, see the
header of that file for the reference file:line citations and how parity is
pinned).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package; the
product package ``sassy_b200`` never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass
from typing import List, Sequence

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

PROFILE = {"dna": 0, "iupac": 1, "ascii": 2}


class _OracleMatch(ctypes.Structure):
    _fields_ = [
        ("text_start", ctypes.c_uint64),
        ("text_end", ctypes.c_uint64),
        ("pattern_idx", ctypes.c_uint32),
        ("pattern_start", ctypes.c_uint32),
        ("pattern_end", ctypes.c_uint32),
        ("cost", ctypes.c_int32),
        ("strand", ctypes.c_uint32),
        ("ops_len", ctypes.c_uint32),
        ("ops_off", ctypes.c_uint64),
        ("text_idx", ctypes.c_uint64),
    ]


@dataclass(frozen=True, order=True)
class Match:
    """Same fields as the reference's ``Match`` (src/search.rs:35-62)."""

    pattern_idx: int
    text_start: int
    text_end: int
    pattern_start: int
    pattern_end: int
    cost: int
    strand: str  # "+" / "-" as in src/python.rs:201-206
    cigar: str
    text_idx: int = 0


USIZE_MAX = 2**64 - 1  # what the reference stores in the untraced fields (src/search.rs:1466-1469)


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (a few hundred ms). Returns the .so path."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "sassy_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(
            ["gcc", "-O2", "-fPIC", "-Wall", "-Wextra", "-std=c11", "-shared", "-o", so, src]
        )
    return so


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build())
        lib.oracle_out_new.restype = ctypes.c_void_p
        lib.oracle_out_free.argtypes = [ctypes.c_void_p]
        lib.oracle_out_len.argtypes = [ctypes.c_void_p]
        lib.oracle_out_len.restype = ctypes.c_size_t
        lib.oracle_out_matches.argtypes = [ctypes.c_void_p]
        lib.oracle_out_matches.restype = ctypes.POINTER(_OracleMatch)
        lib.oracle_out_ops.argtypes = [ctypes.c_void_p]
        lib.oracle_out_ops.restype = ctypes.c_void_p
        lib.oracle_search.argtypes = [
            ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t,
            ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
        ]
        lib.oracle_search.restype = ctypes.c_int
        lib.oracle_search_encoded.argtypes = [
            ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_char_p,
            ctypes.c_size_t, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
        ]
        lib.oracle_search_encoded.restype = ctypes.c_int
        lib.oracle_search_opts.argtypes = [
            ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t,
            ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
            ctypes.c_char_p, ctypes.c_size_t, ctypes.c_float, ctypes.c_int64, ctypes.c_void_p,
        ]
        lib.oracle_search_opts.restype = ctypes.c_int
        lib.oracle_search_many.argtypes = [
            ctypes.c_int, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_void_p,
            ctypes.c_size_t, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
            ctypes.c_float, ctypes.c_int64, ctypes.c_void_p,
        ]
        lib.oracle_search_many.restype = ctypes.c_int
        lib.oracle_search_encoded_nfrac.argtypes = [
            ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_char_p,
            ctypes.c_size_t, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_void_p,
        ]
        lib.oracle_search_encoded_nfrac.restype = ctypes.c_int
        lib.oracle_search_encoded_opts.argtypes = [
            ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_char_p,
            ctypes.c_size_t, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float,
            ctypes.c_int64, ctypes.c_void_p,
        ]
        lib.oracle_search_encoded_opts.restype = ctypes.c_int
        lib.oracle_bottom_row.argtypes = [
            ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t,
            ctypes.c_int, ctypes.c_void_p,
        ]
        lib.oracle_complement.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p]
        lib.oracle_reverse_complement.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p]
        lib.oracle_iupac_valid.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
        lib.oracle_iupac_valid.restype = ctypes.c_int
        _LIB = lib
    return _LIB


def rle(ops: str) -> str:
    """pa_types::Cigar::to_string: run-length ``<cnt><op>`` with cnt always
    printed (pinned by the reference doctests, src/lib.rs:83,107)."""
    out = []
    i = 0
    while i < len(ops):
        j = i
        while j < len(ops) and ops[j] == ops[i]:
            j += 1
        out.append(f"{j - i}{ops[i]}")
        i = j
    return "".join(out)


def _collect(lib, out) -> List[Match]:
    n = lib.oracle_out_len(out)
    ms = lib.oracle_out_matches(out)
    ops_ptr = lib.oracle_out_ops(out)
    res = []
    for i in range(n):
        m = ms[i]
        ops = ctypes.string_at(ops_ptr + m.ops_off, m.ops_len).decode() if m.ops_len else ""
        res.append(
            Match(
                pattern_idx=m.pattern_idx,
                text_start=m.text_start,
                text_end=m.text_end,
                pattern_start=USIZE_MAX if m.pattern_start == 0xFFFFFFFF else m.pattern_start,
                pattern_end=m.pattern_end,
                cost=m.cost,
                strand="-" if m.strand else "+",
                cigar=rle(ops),
                text_idx=m.text_idx,
            )
        )
    return res


class OracleError(RuntimeError):
    pass


def search(alphabet: str, pattern: bytes, text: bytes, k: int, rc: bool = False,
           all_minima: bool = False, without_trace: bool = False, only_best: bool = False,
           max_n_frac: float | None = None, pam: bytes | None = None, alpha: float | None = None,
           max_overhang: int | None = None) -> List[Match]:
    """Searcher::<P>::new(rc, None).search / search_all, optionally under the Searcher options
    (without_trace, only_best_match, max_n_frac) and with the CRISPR end filter of
    search_with_fn (``pam``: the characters before the end position must match it)."""
    lib = _lib()
    out = lib.oracle_out_new()
    try:
        nf = -1.0 if max_n_frac is None or max_n_frac == 1.0 else float(max_n_frac)
        r = lib.oracle_search_opts(PROFILE[alphabet.lower()], pattern, len(pattern), text, len(text),
                                   k, int(rc), int(all_minima), int(without_trace), int(only_best), nf,
                                   pam, len(pam) if pam else 0, -1.0 if alpha is None else float(alpha),
                                   -1 if max_overhang is None else int(max_overhang), out)
        if r == -4:
            raise OracleError("Overhang is not supported for this profile")
        if r == -2:
            raise OracleError("Pattern is not valid IUPAC")
        if r != 0:
            raise OracleError(f"trace failed ({r})")
        return _collect(lib, out)
    finally:
        lib.oracle_out_free(out)


def search_many(alphabet: str, patterns: Sequence[bytes], texts: Sequence[bytes], k: int, rc: bool = False,
                without_trace: bool = False, only_best: bool = False,
                max_n_frac: float | None = None, alpha: float | None = None,
                max_overhang: int | None = None) -> List[Match]:
    """Searcher::search_many (SearchMode::Single order: pattern-major, then text)."""
    lib = _lib()
    out = lib.oracle_out_new()
    try:
        plens = (ctypes.c_uint64 * max(1, len(patterns)))(*[len(p) for p in patterns])
        tlens = (ctypes.c_uint64 * max(1, len(texts)))(*[len(t) for t in texts])
        nf = -1.0 if max_n_frac is None or max_n_frac == 1.0 else float(max_n_frac)
        r = lib.oracle_search_many(PROFILE[alphabet.lower()], b"".join(patterns), plens, len(patterns),
                                   b"".join(texts), tlens, len(texts), k, int(rc), int(without_trace),
                                   int(only_best), nf, -1.0 if alpha is None else float(alpha),
                                   -1 if max_overhang is None else int(max_overhang), out)
        if r == -2:
            raise OracleError("Pattern is not valid IUPAC")
        if r != 0:
            raise OracleError(f"trace failed ({r})")
        return _collect(lib, out)
    finally:
        lib.oracle_out_free(out)


def search_encoded(alphabet: str, patterns: Sequence[bytes], text: bytes, k: int, rc: bool = False,
                   all_minima: bool = False, max_n_frac: float | None = None, alpha: float | None = None,
                   max_overhang: int | None = None) -> List[Match]:
    """encode_patterns + search_encoded_patterns / search_all_encoded_patterns."""
    lib = _lib()
    m = len(patterns[0])
    assert all(len(p) == m for p in patterns)
    out = lib.oracle_out_new()
    try:
        nf = -1.0 if max_n_frac is None or max_n_frac == 1.0 else float(max_n_frac)
        r = lib.oracle_search_encoded_opts(PROFILE[alphabet.lower()], b"".join(patterns), len(patterns),
                                           m, text, len(text), k, int(rc), int(all_minima), nf,
                                           -1.0 if alpha is None else float(alpha),
                                           -1 if max_overhang is None else int(max_overhang), out)
        if r == -4:
            raise OracleError("Overhang is not supported for this profile")
        if r == -2:
            raise OracleError("Pattern is not valid IUPAC")
        if r != 0:
            raise OracleError(f"oracle_search_encoded failed ({r})")
        return _collect(lib, out)
    finally:
        lib.oracle_out_free(out)


def bottom_row(alphabet: str, pattern: bytes, text: bytes, rev: bool = False) -> List[int]:
    lib = _lib()
    buf = (ctypes.c_int32 * (len(text) + 1))()
    lib.oracle_bottom_row(PROFILE[alphabet.lower()], pattern, len(pattern), text, len(text), int(rev), buf)
    return list(buf)


def complement(alphabet: str, s: bytes) -> bytes:
    lib = _lib()
    buf = ctypes.create_string_buffer(len(s))
    lib.oracle_complement(PROFILE[alphabet.lower()], s, len(s), buf)
    return buf.raw


def reverse_complement(alphabet: str, s: bytes) -> bytes:
    lib = _lib()
    buf = ctypes.create_string_buffer(len(s))
    lib.oracle_reverse_complement(PROFILE[alphabet.lower()], s, len(s), buf)
    return buf.raw
