"""ctypes wrapper of oracle/sassy_cpu_port.c -- BASELINE / TEST INFRASTRUCTURE ONLY.

The multi-threaded SIMD restatement of the reference's v1 CPU engine, used by bench.py for
`cpu_baseline` and `--impl reference`.  Never imported by the product package."""
from __future__ import annotations

import ctypes
import os
import subprocess
import time
from typing import List, Sequence, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
PROFILE = {"dna": 0, "iupac": 1}


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libcpuport.so")
    src = os.path.join(_HERE, "sassy_cpu_port.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O3", "-march=native", "-fPIC", "-Wall", "-Wextra", "-std=gnu11", "-pthread",
                               "-shared", "-o", so, src])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        # -march=native: always (re)build on the machine that runs it
        lib = ctypes.CDLL(build(force=os.environ.get("SASSY_CPU_PORT_REBUILD", "1") == "1"))
        lib.cpu_port_search.restype = ctypes.c_size_t
        lib.cpu_port_search.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                        ctypes.POINTER(ctypes.c_double)]
        lib.cpu_port_lanes.restype = ctypes.c_int
        lib.cpu_port_v2_lanes.restype = ctypes.c_int
        lib.cpu_port_search_batch.restype = ctypes.c_size_t
        lib.cpu_port_search_batch.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_int, ctypes.c_void_p,
                                              ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_double)]
        _LIB = lib
    return _LIB


def lanes() -> int:
    return _lib().cpu_port_lanes()


def search_ends(alphabet: str, pattern: bytes, text, n: int, k: int, rc: bool, all_minima: bool, threads: int = 1,
                cap: int = 1 << 20) -> Tuple[List[Tuple[int, int, int]], float]:
    """[(end position in scan direction, cost, strand)], seconds.  `text` is bytes or an address."""
    lib = _lib()
    pos = (ctypes.c_uint64 * cap)()
    cost = (ctypes.c_int32 * cap)()
    strand = (ctypes.c_uint8 * cap)()
    sec = ctypes.c_double()
    addr = ctypes.cast(ctypes.c_char_p(text), ctypes.c_void_p) if isinstance(text, bytes) else ctypes.c_void_p(text)
    total = lib.cpu_port_search(PROFILE[alphabet.lower()], pattern, len(pattern), addr, n, k, int(rc),
                                int(all_minima), threads, pos, cost, strand, cap, ctypes.byref(sec))
    got = min(total, cap)
    return [(pos[i], cost[i], strand[i]) for i in range(got)], sec.value


def kind(threads: int) -> str:
    return f"C restatement of Sassy v1 (search.rs text-tiled u64x{lanes()} + early termination), {threads} threads"


def search_timed(alphabet: str, patterns: Sequence[bytes], k: int, rc: bool, text_addr: int, n: int,
                 threads: int):
    """One pass of every pattern over text[0:n] with `threads` threads.  (seconds, matches, kind)."""
    total = 0
    t0 = time.perf_counter()
    for p in patterns:
        ends, _ = search_ends(alphabet, p, text_addr, n, k, rc, False, threads)
        total += len(ends)
    sec = time.perf_counter() - t0
    kind = f"C restatement of Sassy v1 (search.rs text-tiled u64x{lanes()} + early termination), {threads} threads"
    return sec, total, kind


def calibrate(alphabet: str, patterns: Sequence[bytes], k: int, rc: bool) -> float:
    """Single-thread rate in text-bytes*patterns/s, measured on a small random text."""
    import random
    rng = random.Random(1)
    n = 1 << 22
    text = bytes(rng.choice(b"ACGT") for _ in range(1 << 16)) * (n >> 16)
    p = patterns[0]
    _, sec = search_ends(alphabet, p, text, n, k, rc, False, 1)
    _, sec = search_ends(alphabet, p, text, n, k, rc, False, 1)
    return n / max(sec, 1e-6)


def search_batch(patterns: Sequence[bytes], text, n: int, k: int, all_minima: bool = False, trace: bool = True,
                 prefilter: bool = True, threads: int = 1, cap: int = 1 << 22):
    """v2 engine (pattern tiling, forward strand, Iupac): [(pattern index, end, cost, text_start)], seconds."""
    import numpy as np
    lib = _lib()
    m = len(patterns[0])
    assert all(len(p) == m for p in patterns) and m <= 32
    pat = np.zeros(cap, dtype=np.uint32)
    pos = np.zeros(cap, dtype=np.uint64)
    cost = np.zeros(cap, dtype=np.int32)
    start = np.zeros(cap, dtype=np.uint64)
    sec = ctypes.c_double()
    addr = ctypes.cast(ctypes.c_char_p(text), ctypes.c_void_p) if isinstance(text, bytes) else ctypes.c_void_p(text)
    total = lib.cpu_port_search_batch(b"".join(patterns), len(patterns), m, addr, n, k, int(all_minima), int(trace),
                                      int(prefilter), threads, pat.ctypes.data, pos.ctypes.data, cost.ctypes.data,
                                      start.ctypes.data, cap, ctypes.byref(sec))
    if total > cap:
        return search_batch(patterns, text, n, k, all_minima, trace, prefilter, threads, cap=int(total) + 16)
    got = int(total)
    return list(zip(pat[:got].tolist(), pos[:got].tolist(), cost[:got].tolist(), start[:got].tolist())), sec.value


def kind_v2(threads: int, k: int, m: int) -> str:
    pf = ", 16-character suffix prefilter in u16 lanes" if 1 <= k <= 3 and m > 16 else ""
    return (f"C restatement of Sassy v2 (pattern_tiling/search.rs: u32x{_lib().cpu_port_v2_lanes()} pattern lanes{pf}) "
            f"+ traceback, {threads} threads")
