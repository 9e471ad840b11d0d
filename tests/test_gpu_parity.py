"""Parity of the CUDA path (through the C ABI) with the CPU oracle.  Needs a B200."""
import ctypes
import math
import random

import pytest

import oracle
from tests import kat_util
from tests.test_oracle_props import planted, rand_seq

pytestmark = pytest.mark.gpu


def key(m):
    return (m.pattern_idx, m.text_start, m.text_end, m.cost, m.strand, m.cigar)


# (scan data path, prefilter mode): full scan only, prefilter whenever a piece layout exists,
# and the production setting (prefilter when profitable) on the LDG data path.
@pytest.fixture(scope="module", params=[("tma", "off"), ("tma", "force"), ("ldg", "auto"), ("ldg", "force"),
                                        ("tma", "force", "refine-all")],
                ids=lambda p: "-".join(p))
def backend(request):
    import os
    from tests.gpu_backend import GpuBackend
    if len(request.param) > 2:  # piece-automaton hits refined too (read when a searcher is constructed)
        os.environ["SASSY_B200_REFINE"] = "2"
        os.environ["SASSY_B200_QGRAM"] = "0"
        b = GpuBackend(*request.param[:2])
        for alphabet in ("dna", "iupac"):
            for rc in (False, True):
                b._searcher(alphabet, rc)
        del os.environ["SASSY_B200_REFINE"], os.environ["SASSY_B200_QGRAM"]
        return b
    return GpuBackend(*request.param)


@pytest.fixture(scope="module")
def tma():
    from tests.gpu_backend import GpuBackend
    return GpuBackend("tma")


def test_native_library_is_loaded():
    """The product path is the CUDA extension; there is nothing to fall back to."""
    import sassy_b200
    from sassy_b200 import _native
    assert sassy_b200.device_count() >= 1
    with open("/proc/self/maps") as f:
        assert "libsassy_b200.so" in f.read()
    assert _native.load().sassy_gpu_searcher is not None


@pytest.mark.parametrize("case", kat_util.load_cases(), ids=lambda c: c["source"][:60])
def test_gpu_kat(backend, case):
    kat_util.check(backend, case)


@pytest.mark.parametrize("alphabet", ["dna", "iupac"])
def test_gpu_v1_fuzz(backend, alphabet):
    rng = random.Random(21)
    for it in range(120):
        m = rng.choice([1, 2, 5, 20, 23, 31, 32, 33, 40, 64, 65, 100, 130])
        n = rng.randrange(0, 20000)
        k = rng.randrange(0, max(1, m // 3) + 1)
        p, t = planted(rng, m, max(n, 1), k)
        t = t[:n]
        if alphabet == "iupac" and rng.random() < 0.3:
            t = bytes(c if rng.random() > 0.05 else ord(rng.choice("NRYSWKMBDHV")) for c in t)
        if rng.random() < 0.1:
            t = b"A" * n
            p = b"A" * m
        for allm in (False, True):
            want = oracle.search(alphabet, p, t, k, rc=True, all_minima=allm)
            got = backend.search(alphabet, p, t, k, rc=True, all_minima=allm)
            assert list(map(key, got)) == list(map(key, want)), (alphabet, p, t, k, allm)


def test_gpu_v2_fuzz(backend):
    rng = random.Random(22)
    for it in range(60):
        m = rng.choice([3, 8, 16, 23, 32, 33, 64])
        n = rng.randrange(1, 20000)
        k = rng.randrange(0, max(1, m // 4) + 1)
        P = rng.randrange(1, 40)
        pats = []
        t = bytearray(rand_seq(rng, n))
        for _ in range(P):
            p, tt = planted(rng, m, n, k)
            pats.append(p)
            pos = rng.randrange(0, max(1, n - m))
            t[pos:pos + m] = tt[pos:pos + m]
        t = bytes(t[:n])
        for allm in (False, True):
            want = oracle.search_encoded("iupac", pats, t, k, rc=True, all_minima=allm)
            got = backend.search_encoded("iupac", pats, t, k, rc=True, all_minima=allm)
            assert list(map(key, got)) == list(map(key, want)), (pats, t, k, allm)


def test_gpu_long_patterns(backend):
    rng = random.Random(23)
    for m in (150, 200, 260, 500, 1000):
        n = 30000
        k = 6
        p, t = planted(rng, m, n, k)
        want = oracle.search("dna", p, t, k, rc=True)
        got = backend.search("dna", p, t, k, rc=True)
        assert want and list(map(key, got)) == list(map(key, want)), m


def test_gpu_ascii(tma):
    import sassy_b200
    s = sassy_b200.Searcher("ascii", rc=False)
    text = b"the quick brown fox jumps over the lazy dog; the quack brown fax"
    for k in (0, 2, 4):
        for allm in (False, True):
            want = oracle.search("ascii", b"quick brown fox", text, k, rc=False, all_minima=allm)
            got = tma.search("ascii", b"quick brown fox", text, k, rc=False, all_minima=allm)
            assert list(map(key, got)) == list(map(key, want))
    ms = s.search(b"quick brown fox", text, 2)
    assert [(m.text_start, m.text_end, m.cost, m.cigar) for m in ms] == [(4, 19, 0, "15="), (49, 64, 2, "2=1X10=1X1=")]
    assert s.search(b"QUICK", text, 0) == []  # case-sensitive (reference Ascii<true>)


def test_gpu_multi_megabyte_rows(tma):
    """configs[0] shape (1 MB, m=20, k=1) and a 64 MB text: candidates at row boundaries."""
    import sassy_b200
    rng = random.Random(24)
    s = sassy_b200.Searcher("dna", rc=True)
    for n in (1 << 20, 1 << 24):
        p = rand_seq(rng, 20)
        table = bytes(b"ACGT"[c & 3] for c in range(256))
        t = bytearray(rng.randbytes(n).translate(table))
        # plant exact and 1-edit copies at every multiple of 128 +- a few, covering all row boundaries of any tiling
        planted_at = []
        for i in range(200):
            pos = rng.randrange(0, n - 40)
            if i % 2 == 0:
                pos = (pos // 128) * 128 - rng.randrange(0, 24)
                pos = max(pos, 0)
            t[pos:pos + 20] = p
            planted_at.append(pos)
        t = bytes(t)
        got = s.search(p, t, 1)
        ends = {m.text_end for m in got if m.strand == "+" and m.cost == 0}
        for pos in planted_at:
            assert pos + 20 in ends, (n, pos)
        if n == 1 << 20:
            want = oracle.search("dna", p, t, 1, rc=True)
            g = [(m.text_start, m.text_end, m.cost, m.strand, m.cigar) for m in got]
            w = [(m.text_start, m.text_end, m.cost, m.strand, m.cigar) for m in want]
            assert g == w


def test_gpu_candidate_overflow_retry(tma):
    """Dense candidates (every position matches) overflow the first buffer and trigger a re-scan."""
    import sassy_b200
    s = sassy_b200.Searcher("iupac", rc=False)
    n = 3_000_000
    t = b"N" * n
    ms = s.search(b"ACGTACGT", t, 0)  # one plateau -> a single local minimum at the text end
    assert [(m.text_start, m.text_end, m.cost) for m in ms] == [(n - 8, n, 0)]
    # every end position is a candidate (the regional pass scans the dense tiles whole; the hits of a
    # sparse last tile are re-scanned, which may list a few positions twice)
    st = s.stats()
    assert st["retries"] >= 1 and n - 7 <= st["candidates"] <= n - 7 + 20000, st
    assert st["dense_tiles"] >= 1


def test_prefilter_routes(tma):
    """The production rule: prefilter for selective pieces, full scan otherwise, fallback on repeats."""
    import sassy_b200
    rng = random.Random(26)
    s = sassy_b200.Searcher("dna", rc=True)
    n = 1 << 22
    table = bytes(b"ACGT"[c & 3] for c in range(256))
    t = bytearray(rng.randbytes(n).translate(table))
    p = rand_seq(rng, 20)
    for pos in (0, 5, 1000, n // 2 - 3, n - 20, n - 25):
        t[pos:pos + 20] = p
    t = bytes(t)
    want = oracle.search("dna", p, t, 2, rc=True)
    got = s.search(p, t, 2)
    st = s.stats()
    # rc=True: both strands share one pass, one automaton word each
    assert st["filter_words"] == 2 and st["filter_fallback"] == 0 and st["hits"] > 0
    assert [key(m) for m in got] == [key(m) for m in want] and len(got) >= 4
    s.set_filter("off")
    assert [key(m) for m in s.search(p, t, 2)] == [key(m) for m in got]
    assert s.stats()["filter_words"] == 0
    s.set_filter("auto")
    # k too large for selective pieces -> planned off
    s.search(p, t[:100000], 6)
    assert s.stats()["filter_words"] == 0
    # repetitive text: every word is a hit -> the regional pass scans the dense tiles whole, same answer
    rep = (p * (200000 // 20))
    a = s.search(p, rep, 1)
    st = s.stats()
    assert st["filter_fallback"] == 0 and st["dense_tiles"] >= 1 and st["retries"] >= 1
    s.set_filter("off")
    b = s.search(p, rep, 1)
    assert [key(m) for m in a] == [key(m) for m in b] and len(a) > 1000


def test_dna_packed_transport(tma):
    """Large Dna host texts cross PCIe at 2 bits per character; results equal the byte transport,
    mixed case included; a text with a byte outside ACGTacgt is sent as bytes."""
    import sassy_b200
    rng = random.Random(27)
    n = (9 << 20) + 37  # not a multiple of 64
    table = bytes(b"ACGTacgt"[c & 7] for c in range(256))
    t = bytearray(rng.randbytes(n).translate(table))
    p = rand_seq(rng, 24)
    for pos in (0, 3, 777, n // 2, n - 24, n - 61):
        t[pos:pos + 24] = p if pos % 2 else p.lower()
    t = bytes(t)
    s = sassy_b200.Searcher("dna", rc=True)
    s.set_transport("bytes")
    a = s.search(p, t, 2)
    assert s.stats()["transfer_packed"] == 0
    s.set_transport("packed")
    b = s.search(p, t, 2)
    assert s.stats()["transfer_packed"] == 1
    assert a == b and len(a) >= 5
    dt = s.upload_text(t)
    assert s.search(p, dt, 2) == a
    dt.free()
    want = oracle.search("dna", p, t[:200000], 2, rc=True)
    assert [key(m) for m in s.search(p, t[:200000], 2)] == [key(m) for m in want]  # small text: byte path
    t2 = bytearray(t)
    t2[n // 3] = ord("N")
    s.search(p, bytes(t2), 2)
    assert s.stats()["transfer_packed"] == 0
    # the Iupac profile never packs
    s2 = sassy_b200.Searcher("iupac", rc=False)
    s2.search(p, t, 1)
    assert s2.stats()["transfer_packed"] == 0


def test_dna_packed_transport_pinned_ring(tma):
    """A pinned host text longer than the staging ring (48 x 2 Mi characters): ring slots are
    re-used, chunks from the back of the text cross as plain bytes, matches straddle chunk borders
    and the packed / plain border; a foreign byte late in the text (after most of it has been
    expanded on the device) falls back to the byte transport with the same result."""
    import sassy_b200
    rng = random.Random(28)
    chunk = 2 << 20
    n = 120 * chunk + 4099
    table = bytes(b"ACGTacgt"[c & 7] for c in range(256))
    t = bytearray(rng.randbytes(n).translate(table))
    p = rand_seq(rng, 32)
    spots = [0, chunk - 16, 47 * chunk - 5, 48 * chunk - 31, 49 * chunk + 1, 90 * chunk - 16, 100 * chunk - 1,
             110 * chunk - 17, 119 * chunk - 3, n - 32]
    for pos in spots:
        t[pos:pos + 32] = p
    addr = sassy_b200.host_alloc(n)
    try:
        ctypes.memmove(addr, bytes(t), n)
        s = sassy_b200.Searcher("dna", rc=True)
        s.set_transport("bytes")
        want = s.search(p, (addr, n), 3)
        assert s.stats()["transfer_packed"] == 0
        assert sorted(m.text_start for m in want if m.cost == 0 and m.strand == "+") == sorted(spots)
        s.set_transport("packed")
        for _ in range(3):  # the split between packed and plain bytes differs from call to call
            got = s.search(p, (addr, n), 3)
            st = s.stats()
            assert st["transfer_packed"] == 1 and st["transfer_bytes"] < n
            assert got == want
        # a foreign byte near the end: its chunk usually crosses as plain bytes (no fall-back needed);
        # one in the first half is met by the packer, and the whole text is sent as bytes
        for at, must_fall_back in ((117 * chunk + 12345, False), (30 * chunk + 777, True)):
            ctypes.memset(addr + at, ord("N"), 1)
            s.set_transport("bytes")
            want2 = s.search(p, (addr, n), 3)
            s.set_transport("packed")
            got2 = s.search(p, (addr, n), 3)
            assert got2 == want2
            if must_fall_back:
                assert s.stats()["transfer_packed"] == 0
    finally:
        sassy_b200.host_free(addr)


def test_c_abi_search_symbol():
    """The reference's own entry points (include/sassy.h), called as c/example.c does."""
    from sassy_b200 import _native
    lib = _native.load()
    h = lib.sassy_searcher(b"dna", True, math.nan)
    assert h
    pattern = b"AAGGGGA"
    text = b"CCCCCCCCCAAGGGGACCCCCAAGGCGACCCCCCCCC"
    out = ctypes.POINTER(_native.CMatch)()
    n = lib.search(h, pattern, len(pattern), text, len(text), 1, ctypes.byref(out))
    got = [(out[i].text_start, out[i].text_end, out[i].pattern_start, out[i].pattern_end, out[i].cost,
            out[i].strand) for i in range(n)]
    want = [(m.text_start, m.text_end, m.pattern_start, m.pattern_end, m.cost, 1 if m.strand == "-" else 0)
            for m in oracle.search("dna", pattern, text, 1, rc=True)]
    assert got == want and n >= 2
    assert ctypes.sizeof(_native.CMatch) == 40
    lib.sassy_matches_free(out, n)
    # zero matches: still a non-null pointer (reference src/c.rs:112-127)
    n = lib.search(h, b"TTTTTTTT", 8, text, len(text), 0, ctypes.byref(out))
    assert n == 0 and bool(out)
    lib.sassy_matches_free(out, 0)
    lib.sassy_searcher_free(h)


def test_device_text_reuse(tma):
    import sassy_b200
    rng = random.Random(25)
    s = sassy_b200.Searcher("iupac", rc=True)
    p, t = planted(rng, 23, 50000, 3)
    dt = s.upload_text(t)
    a = s.search(p, dt, 3)
    b = s.search(p, t, 3)
    assert a == b and a
    enc = s.encode_patterns([p, rand_seq(rng, 23)])
    c = s.search_encoded_patterns(enc, dt, 3)
    d = s.search_encoded_patterns(enc, t, 3)
    assert c == d and c
    dt.free()


def test_gpu_long_patterns_many_pieces(tma):
    """m up to 1000 with k >= 8: the 8-word prefilter automaton, the warp-systolic re-scan and the
    warp-systolic traceback (patterns of >= 8 words)."""
    import sassy_b200
    rng = random.Random(29)
    s = sassy_b200.Searcher("dna", rc=True)
    for m, k, n in ((300, 8, 200_000), (1000, 8, 300_000), (1000, 3, 300_000), (520, 12, 100_000), (260, 9, 50_000)):
        p, t = planted(rng, m, n, k)
        t = bytearray(t)
        for j in range(5):  # a few more copies with edits
            a = rng.randrange(0, n - m - 20)
            q = bytearray(p)
            for _ in range(rng.randrange(0, k + 1)):
                q[rng.randrange(len(q))] = rng.choice(b"ACGT")
            t[a:a + len(q)] = q
        t = bytes(t[:n])
        for allm in (False, True):
            want = oracle.search("dna", p, t, k, rc=True, all_minima=allm)
            got = s.search_all(p, t, k) if allm else s.search(p, t, k)
            assert list(map(key, got)) == list(map(key, want)), (m, k, allm)
        assert len(want) >= 1
    assert s.stats()["filter_words"] in (1, 2, 4, 8)


def test_gpu_patterns_beyond_32_words():
    """Patterns of 1025..4096 characters: several words per lane in the warp-systolic re-scan and
    traceback; q-gram / piece-automaton hits, or (prefilter off, k too large) windows covering the text."""
    import sassy_b200
    from tests.test_oracle_props import mutate
    rng = random.Random(37)
    s = sassy_b200.Searcher("dna", rc=True)
    si = sassy_b200.Searcher("iupac", rc=True)
    routes = set()
    for m, k, n, mode in ((1025, 5, 90_000, "auto"), (1500, 12, 200_000, "auto"), (2048, 3, 70_000, "off"),
                          (2049, 20, 170_000, "auto"), (3000, 40, 50_000, "auto"), (4096, 9, 300_000, "auto"),
                          (4096, 63, 40_000, "off"), (1100, 300, 20_000, "auto")):
        p, t = planted(rng, m, n, k)
        t = bytearray(t)
        for at in (max(0, 8192 - m // 2), n // 2):
            q = bytearray(mutate(rng, p, rng.randrange(0, k + 1)))
            if at + len(q) < n:
                t[at:at + len(q)] = q
        t = bytes(t[:n])
        s.set_filter(mode)
        for allm in (False, True):
            want = oracle.search("dna", p, t, k, rc=True, all_minima=allm)
            got = s.search_all(p, t, k) if allm else s.search(p, t, k)
            assert list(map(key, got)) == list(map(key, want)), (m, k, mode, allm)
            for a, b in zip(got, want):
                assert a.cigar == b.cigar
        assert want, m
        routes.add(s.stats()["filter_kind"])
    assert routes >= {0, 2}, routes
    # Iupac pattern with ambiguity codes (piece automaton or cover), text with N
    m, k, n = 1300, 6, 60_000
    p, t = planted(rng, m, n, k)
    p, t = bytearray(p), bytearray(t)
    for _ in range(20):
        p[rng.randrange(m)] = rng.choice(b"NRYSW")
    p = bytes(p)
    concrete = bytes({"N": b"ACGT", "R": b"AG", "Y": b"CT", "S": b"CG", "W": b"AT"}.get(chr(c), bytes([c]))[0] for c in p)
    t[20_000:20_000 + m] = mutate(rng, concrete, 4)
    t[30_000] = ord("N")
    t = bytes(t[:n])
    for allm in (False, True):
        want = oracle.search("iupac", p, t, k, rc=True, all_minima=allm)
        got = si.search_all(p, t, k) if allm else si.search(p, t, k)
        assert want and list(map(key, got)) == list(map(key, want)), allm


def test_gpu_qgram_prefilter_fuzz():
    """The q-gram bitmap route (Dna, one pattern, shares >= 9 characters): every (Q, S) instantiation,
    both strands in one pass, against the oracle; also checks that the route is the one taken."""
    import sassy_b200
    from tests.test_oracle_props import mutate
    rng = random.Random(71)
    seen = set()
    searchers = {}
    for rc_ in (False, True):
        searchers[rc_] = sassy_b200.Searcher("dna", rc=rc_)
        searchers[rc_].set_filter("force")
    for it in range(90):
        m, k = rng.choice([(20, 1), (30, 2), (36, 3), (40, 3), (60, 3), (100, 8), (100, 3), (120, 5), (64, 2),
                           (200, 8), (18, 1), (27, 2), (150, 1), (500, 8), (1000, 8), (260, 6)])
        n = rng.randrange(0, 60000)
        p, t = planted(rng, m, max(n, 1), k)
        t = bytearray(t[:n])
        for _ in range(4):
            q = mutate(rng, p, rng.randrange(0, k + 1))
            if rng.random() < 0.5:
                q = oracle.reverse_complement("dna", q)
            pos = rng.randrange(0, max(1, n))
            if pos + len(q) <= n:
                t[pos:pos + len(q)] = q
        if rng.random() < 0.15:
            t = bytearray(c | 0x20 if rng.random() < 0.3 else c for c in t)
        t = bytes(t)
        rc = rng.random() < 0.7
        s = searchers[rc]
        for allm in (False, True):
            want = oracle.search("dna", p, t, k, rc=rc, all_minima=allm)
            got = (s.search_all if allm else s.search)(p, t, k)
            assert [(x.text_start, x.text_end, x.cost, x.strand, x.cigar) for x in got] == \
                   [(x.text_start, x.text_end, x.cost, x.strand, x.cigar) for x in want], (p, t, k, rc, allm)
            st = s.stats()
            if n:
                assert st["filter_kind"] == 2, st  # (small texts may fall back to the full scan afterwards)
                seen.add((st["filter_len"], m // (k + 1) >= 8 + 15, m // (k + 1) >= 8 + 7))
    assert len(seen) >= 5, seen  # Q in {6, 7, 8} and S in {4, 8, 16}


@pytest.mark.parametrize("seq", ["1", "0"], ids=["contiguous-tiles", "row-tiles"])
def test_gpu_qgram_large_text(seq, monkeypatch):
    """64 MB: thousands of false q-gram hits (rejected by the exact confirmation) next to planted
    copies on every kind of row / tile boundary; equals the full scan.  Both q-gram kernels: the
    contiguous-tile one (default) and the row-tiled one (SASSY_B200_QGRAM_SEQ=0)."""
    import sassy_b200
    monkeypatch.setenv("SASSY_B200_QGRAM_SEQ", seq)  # read when the searcher is constructed
    rng = random.Random(72)
    n = 1 << 26
    base = bytes(rng.choice(b"ACGT") for _ in range(1 << 16))
    for m, k in ((100, 8), (40, 3), (1000, 8)):
        p = rand_seq(rng, m)
        t = bytearray(base * (n >> 16))
        rcp = oracle.reverse_complement("dna", p)
        spots = [0, n - m, 13312 * 7 - 5, 9984 * 3 - m // 2, 16384 * 11 + 1, 64 * 1000 - 3, 5_000_011, 33_000_000,
                 2048 * 4001 - 9, 2048 * 9000 - m + 3, 64 * 77777 - 1]
        for i, pos in enumerate(spots):
            q = bytearray(p if i % 2 == 0 else rcp)
            for _ in range(i % (k + 1)):
                q[rng.randrange(m)] = rng.choice(b"ACGT")
            t[pos:pos + m] = q
        t = bytes(t)
        s = sassy_b200.Searcher("dna", rc=True)
        s.set_filter("force")
        got = s.search(p, t, k)
        st = s.stats()
        assert st["filter_kind"] == 2 and st["filter_fallback"] == 0 and st["hits"] > 0, st
        s.set_filter("off")
        want = s.search(p, t, k)
        assert list(map(key, got)) == list(map(key, want)) and len(got) >= len(spots)
        # the planted copies, checked by the oracle on their own neighbourhoods
        for pos in spots:
            lo, hi = max(0, pos - 2 * (m + k)), min(n, pos + m + 2 * (m + k))
            w = oracle.search("dna", p, t[lo:hi], k, rc=True, all_minima=True)
            assert any(x.text_start >= lo and x.text_end <= hi and
                       (x.text_start - lo, x.text_end - lo, x.cost, x.strand, x.cigar) in
                       {(y.text_start, y.text_end, y.cost, y.strand, y.cigar) for y in w} for x in got)


def test_gpu_ascii_fuzz(tma):
    """Ascii profile (case-sensitive byte equality, reference src/profiles/ascii.rs): random
    printable texts with planted, edited copies; forward strand only (the reference has no
    reverse complement for Ascii), search and search_all."""
    rng = random.Random(73)
    alpha = b"abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789 .,;:-_!?()[]{}"
    for it in range(80):
        m = rng.choice([1, 3, 7, 15, 31, 32, 33, 64, 100, 200])
        n = rng.randrange(0, 12000)
        k = rng.randrange(0, max(1, m // 3) + 1)
        small = rng.random() < 0.3
        al = alpha[:4] if small else alpha
        p = bytes(rng.choice(al) for _ in range(m))
        t = bytearray(rng.choice(al) for _ in range(n))
        for _ in range(3):
            q = bytearray(p)
            for _ in range(rng.randrange(0, k + 1)):
                op = rng.randrange(3)
                pos = rng.randrange(len(q))
                if op == 0:
                    q[pos] = rng.choice(al)
                elif op == 1:
                    q.insert(pos, rng.choice(al))
                elif len(q) > 1:
                    del q[pos]
            pos = rng.randrange(0, max(1, n))
            if pos + len(q) <= n:
                t[pos:pos + len(q)] = q
        t = bytes(t)
        for allm in (False, True):
            want = oracle.search("ascii", p, t, k, rc=False, all_minima=allm)
            got = tma.search("ascii", p, t, k, rc=False, all_minima=allm)
            assert list(map(key, got)) == list(map(key, want)), (p, t, k, allm)


def test_c_abi_search_beyond_capacity_does_not_abort():
    """A pattern of more than 4096 characters is a limit of this implementation, not one of the
    reference's panics: search() returns no matches, leaves the reason in sassy_gpu_last_error()
    and the process lives (include/sassy.h)."""
    from sassy_b200 import _native
    lib = _native.load()
    s = lib.sassy_searcher(b"dna", True, math.nan)
    assert s
    rng = random.Random(5)
    pat = bytes(rng.choice(b"ACGT") for _ in range(4097))
    text = pat + bytes(rng.choice(b"ACGT") for _ in range(5000))
    out = ctypes.POINTER(_native.CMatch)()
    n = lib.search(s, pat, len(pat), text, len(text), 3, ctypes.byref(out))
    assert n == 0 and bool(out)
    assert "4096" in _native.last_error()
    lib.sassy_matches_free(out, n)
    n = lib.search(s, pat[:4096], 4096, text, len(text), 3, ctypes.byref(out))  # the longest supported pattern
    assert n == 1 and out[0].text_start == 0 and out[0].text_end == 4096 and out[0].cost == 0
    assert _native.last_error() == ""
    lib.sassy_matches_free(out, n)
    lib.sassy_searcher_free(s)


@pytest.mark.parametrize("refine", ["1", "2"], ids=["default", "refine-all"])
def test_gpu_regional_fallback_low_complexity(refine, monkeypatch):
    """Repeat-rich text (poly-A, di-/tri-nucleotide repeats, a 7-mer satellite, up to 200 KB each)
    and patterns holding the repeat unit: the tiles the prefilter fires all over are scanned whole,
    the rest is prefiltered; results equal the full scan (and the oracle on a sub-range)."""
    import sassy_b200
    from tests.test_host_emulation import _low_complexity_case
    monkeypatch.setenv("SASSY_B200_REFINE", refine)
    rng = random.Random(88)
    dense_seen = 0
    for it in range(10):
        n = rng.randrange(2_000_000, 6_000_000)
        p, t, k = _low_complexity_case(rng, n)
        s = sassy_b200.Searcher("dna", rc=True)
        s.set_filter("force")
        full = sassy_b200.Searcher("dna", rc=True)
        full.set_filter("off")
        for allm in (False, True):
            got = (s.search_all if allm else s.search)(p, t, k)
            st = s.stats()
            want = (full.search_all if allm else full.search)(p, t, k)
            assert list(map(key, got)) == list(map(key, want)), (p, k, allm, st)
            assert st["filter_words"] > 0 and st["filter_fallback"] == 0, st
            dense_seen += st["dense_tiles"] > 0
        # the oracle on a window around the first repeat-borne match
        if want:
            x = want[len(want) // 2]
            lo, hi = max(0, x.text_start - 3000), min(n, x.text_end + 3000)
            w = {(y.text_start + lo, y.text_end + lo, y.cost, y.strand, y.cigar)
                 for y in oracle.search("dna", p, t[lo:hi], k, rc=True, all_minima=True)}
            assert (x.text_start, x.text_end, x.cost, x.strand, x.cigar) in w
    assert dense_seen >= 3, dense_seen
