"""CPU-only checks of the logic the CUDA kernels execute (scan_core.cuh / host_logic.h),
run through the host emulator and compared with the oracle: known-answer vectors, then a
seeded differential fuzz across row lengths so that matches straddle row boundaries."""
import os
import random

import pytest

import oracle
from tests import kat_util
from tests.emu_backend import EmuBackend
from tests.test_oracle_props import mutate, planted, rand_seq


def key(m):
    return (m.pattern_idx, m.text_start, m.text_end, m.cost, m.strand, m.cigar)


@pytest.mark.parametrize("case", kat_util.load_cases(), ids=lambda c: c["source"][:60])
def test_emu_kat(case):
    kat_util.check(EmuBackend(), case)


@pytest.mark.parametrize("case", kat_util.load_cases(), ids=lambda c: c["source"][:60])
def test_emu_kat_short_rows(case):
    kat_util.check(EmuBackend(ltot=128), case)


@pytest.mark.parametrize("alphabet", ["dna", "iupac"])
def test_emu_v1_fuzz(alphabet):
    rng = random.Random(11)
    for it in range(150):
        m = rng.choice([1, 2, 5, 20, 23, 31, 32, 33, 40, 64, 65, 100, 130])
        n = rng.randrange(0, 1500)
        k = rng.randrange(0, max(1, m // 3) + 1)
        p, t = planted(rng, m, max(n, 1), k)
        t = t[:n]
        if alphabet == "iupac" and rng.random() < 0.3:
            t = bytes(c if rng.random() > 0.05 else ord(rng.choice("NRYSWKMBDHV")) for c in t)
        if rng.random() < 0.15:  # homopolymer plateaus
            t = b"A" * n
            p = b"A" * m
        ltot = rng.choice([0, 128, 256, 384])
        for allm in (False, True):
            want = oracle.search(alphabet, p, t, k, rc=True, all_minima=allm)
            got = EmuBackend(ltot=ltot).search(alphabet, p, t, k, rc=True, all_minima=allm)
            assert list(map(key, got)) == list(map(key, want)), (alphabet, p, t, k, allm, ltot)


def test_emu_v2_fuzz():
    rng = random.Random(12)
    for it in range(80):
        m = rng.choice([3, 8, 16, 23, 32, 33, 64])
        n = rng.randrange(1, 1200)
        k = rng.randrange(0, max(1, m // 4) + 1)
        P = rng.randrange(1, 5)
        pats = []
        t = bytearray(rand_seq(rng, n))
        for _ in range(P):
            p, tt = planted(rng, m, n, k)
            pats.append(p)
            pos = rng.randrange(0, max(1, n - m))
            t[pos:pos + m] = tt[pos:pos + m]
        t = bytes(t[:n])
        ltot = rng.choice([0, 128, 256])
        for allm in (False, True):
            want = oracle.search_encoded("iupac", pats, t, k, rc=True, all_minima=allm)
            got = EmuBackend(ltot=ltot).search_encoded("iupac", pats, t, k, rc=True, all_minima=allm)
            assert sorted(map(key, got)) == sorted(map(key, want)), (pats, t, k, allm, ltot)
            assert list(map(key, got)) == list(map(key, want))  # same order: query slot, then end


@pytest.mark.parametrize("mode", [1, -1, 4])
def test_emu_prefilter_fuzz(mode):
    """Exact piece prefilter + re-scan of the hit neighbourhoods == full scan == oracle."""
    rng = random.Random(14)
    used = 0
    for it in range(150):
        m = rng.choice([2, 5, 12, 20, 23, 33, 64, 100])
        n = rng.randrange(0, 1500)
        k = rng.randrange(0, max(1, m // 4) + 1)
        p, t = planted(rng, m, max(n, 1), k)
        t = t[:n]
        if rng.random() < 0.1:
            t, p = b"A" * n, b"A" * m
        alphabet = rng.choice(["dna", "iupac"])
        for allm in (False, True):
            want = oracle.search(alphabet, p, t, k, rc=True, all_minima=allm)
            b = EmuBackend(ltot=rng.choice([0, 64, 128, 192]), use_filter=mode)
            got = b.search(alphabet, p, t, k, rc=True, all_minima=allm)
            assert list(map(key, got)) == list(map(key, want)), (alphabet, p, t, k, allm, b.last_filter)
            used += b.last_filter[0] > 0
    assert used > 50


def test_emu_qgram_prefilter_fuzz():
    """q-gram bitmap prefilter (sampled Q-grams of the k+1 shares, both strands in one table) + exact
    confirmation + re-scan == oracle, for every (Q, S) the kernel is instantiated for; rc on and off,
    planted copies with edits next to row borders, homopolymers, mixed case and non-ACGT text bytes
    (the Dna profile searches every byte as (c >> 1) & 3)."""
    rng = random.Random(21)
    seen = set()
    for it in range(260):
        m, k = rng.choice([(20, 1), (30, 2), (36, 3), (40, 3), (60, 3), (100, 8), (100, 3), (120, 5), (64, 2),
                           (200, 8), (99, 10), (18, 1), (27, 2), (150, 1)])
        n = rng.randrange(0, 2500)
        p, t = planted(rng, m, max(n, 1), k)
        t = bytearray(t[:n])
        for _ in range(3):  # copies (some reverse-complemented) straddling row borders of 64 / 128 / 192 bytes
            q = mutate(rng, p, rng.randrange(0, k + 1))
            if rng.random() < 0.5:
                q = oracle.reverse_complement("dna", q)
            pos = rng.choice([64, 128, 192, 256, 384]) * rng.randrange(1, 6) - rng.randrange(0, len(q) + 1)
            if 0 <= pos and pos + len(q) <= n:
                t[pos:pos + len(q)] = q
        if rng.random() < 0.1:
            t, p = bytearray(b"A" * n), b"A" * m
        if rng.random() < 0.2:
            t = bytearray(c | 0x20 if rng.random() < 0.3 else c for c in t)
        t = bytes(t)
        rc = rng.random() < 0.7
        for allm in (False, True):
            want = oracle.search("dna", p, t, k, rc=rc, all_minima=allm)
            # 2: contiguous-tile kernel (product default), 3: row-tiled kernel
            b = EmuBackend(ltot=rng.choice([0, 64, 128, 192]), use_filter=2 + (it + allm) % 2)
            got = b.search("dna", p, t, k, rc=rc, all_minima=allm)
            assert list(map(key, got)) == list(map(key, want)), (p, t, k, rc, allm, b.last_filter)
            if n and b.last_filter[0] == 1 and 6 <= b.last_filter[1] <= 8:  # else: no q-gram plan, piece automaton
                seen.add((m, k))
    assert len(seen) >= 11, seen


def test_qgram_plan_and_tables():
    """Every share is long enough for its sampled Q-gram (|share| >= Q + S - 1), the shares tile the
    pattern, and the table holds exactly the Q-grams at offsets 0..S-1 of every share."""
    import ctypes
    from tests.emu_backend import _lib
    lib = _lib()
    lib.emu_plan_qgram.restype = ctypes.c_int
    lib.emu_plan_qgram.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_int]
    out = (ctypes.c_int * 256)()
    for m in range(1, 260, 3):
        for k in range(0, 20):
            cnt = lib.emu_plan_qgram(m, k, 2, out, 256)
            enabled, q, s, npieces = out[0], out[1], out[2], out[3]
            if not enabled:
                assert m // (k + 1) < 9 or (k + 1) * 2 / 4 ** min(8, m // (k + 1) - 3) > 4e-3
                continue
            assert cnt == 4 + 2 * npieces and npieces == k + 1
            assert 6 <= q <= 8 and s in (4, 8, 16)
            pieces = [(out[4 + 2 * i], out[5 + 2 * i]) for i in range(npieces)]
            assert pieces[0][0] == 0 and sum(l for _, l in pieces) == m
            for (o, l), (o2, _) in zip(pieces, pieces[1:]):
                assert o + l == o2
            assert min(l for _, l in pieces) >= q + s - 1


def test_emu_prefilter_encoded():
    rng = random.Random(15)
    for it in range(40):
        m = rng.choice([16, 23, 32, 40])
        n = rng.randrange(1, 1200)
        k = rng.randrange(0, 3)
        pats = []
        t = bytearray(rand_seq(rng, n))
        for _ in range(rng.randrange(1, 4)):
            p, tt = planted(rng, m, n, k)
            pats.append(p)
            pos = rng.randrange(0, max(1, n - m))
            t[pos:pos + m] = tt[pos:pos + m]
        t = bytes(t[:n])
        want = oracle.search_encoded("iupac", pats, t, k, rc=True)
        b = EmuBackend(use_filter=1)
        got = b.search_encoded("iupac", pats, t, k, rc=True)
        assert list(map(key, got)) == list(map(key, want)) and b.last_filter[0] > 0


def test_emu_long_pattern_words():
    # every supported word count, incl. the padded ones (W = 6, 8, 16, 32)
    rng = random.Random(13)
    for m in (150, 200, 260, 500, 1000):
        n = 3000
        k = 6
        p, t = planted(rng, m, n, k)
        want = oracle.search("dna", p, t, k, rc=True)
        got = EmuBackend(ltot=rng.choice([0, 1152])).search("dna", p, t, k, rc=True)
        assert list(map(key, got)) == list(map(key, want)), m
        assert want, m


def test_emu_patterns_beyond_32_words():
    """Patterns of 1025..4096 characters (64 / 128 words): no row-tiled kernels; the prefilter routes
    re-scan their hits, the route without prefilter re-scans windows that cover the text."""
    rng = random.Random(19)
    for m, k, n, uf in ((1025, 5, 9000, 0), (1500, 12, 20000, 0), (2048, 3, 17000, -1), (2049, 20, 17000, -1),
                        (3000, 40, 9000, 1), (4096, 9, 30000, 0), (4096, 63, 10000, -1), (1100, 300, 5000, -1)):
        p, t = planted(rng, m, n, k)
        t = bytearray(t)
        q = bytearray(mutate(rng, p, rng.randrange(0, k + 1)))  # a second copy that straddles a window border
        at = max(0, 8192 - m // 2)
        if at + len(q) < n:
            t[at:at + len(q)] = q
        t = bytes(t[:n])
        for allm in (False, True):
            want = oracle.search("dna", p, t, k, rc=True, all_minima=allm)
            got = EmuBackend(use_filter=uf).search("dna", p, t, k, rc=True, all_minima=allm)
            assert list(map(key, got)) == list(map(key, want)), (m, k, uf, allm)
        assert want, m


def test_emu_k_ge_m_and_empty():
    b = EmuBackend()
    assert b.search("dna", b"ACG", b"", 1) == []
    for k in (3, 5):
        want = oracle.search("dna", b"ACG", b"TTACGTT", k, rc=True, all_minima=True)
        got = b.search("dna", b"ACG", b"TTACGTT", k, rc=True, all_minima=True)
        assert list(map(key, got)) == list(map(key, want))
        want = oracle.search("dna", b"ACG", b"TTACGTT", k, rc=True)
        got = b.search("dna", b"ACG", b"TTACGTT", k, rc=True)
        assert list(map(key, got)) == list(map(key, want))


def test_geometry_covers_text():
    b = EmuBackend(bpw=444)
    t = rand_seq(random.Random(5), 100000)
    b.search("dna", b"ACGTACGTACGTACGTACGT", t, 2)
    ltot, rows = b.last_geom
    assert ltot % 128 == 0 and rows * ltot >= len(t) and (rows - 1) * ltot < len(t)


def test_emu_prefilter_eight_words():
    """k + 1 >= 9 pieces of a long pattern need the 8-word automaton."""
    rng = random.Random(17)
    b = EmuBackend()
    b.use_filter = 1
    used = set()
    for it in range(6):
        m = rng.choice([200, 300, 500])
        k = rng.choice([8, 9, 12])
        p, t = planted(rng, m, 5000, k)
        got = b.search("dna", p, t, k, rc=True)
        want = oracle.search("dna", p, t, k, rc=True)
        assert [(x.text_start, x.text_end, x.cost, x.strand, x.cigar) for x in got] == \
               [(x.text_start, x.text_end, x.cost, x.strand, x.cigar) for x in want]
        used.add(b.last_filter[0])
    assert 8 in used


def test_filter_plan_invariants():
    """plan_filter (csrc/host_logic.h): k + 1 pairwise disjoint pieces inside the pattern, every piece
    with its 3 delay bits inside one 32-bit automaton word, no two pieces sharing a bit."""
    import ctypes
    from tests import emu_backend
    lib = emu_backend._lib() if hasattr(emu_backend, "_lib") else ctypes.CDLL(
        os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sassy_b200", "lib", "libsassy_b200_emu.so"))
    lib.emu_plan_filter.restype = ctypes.c_int
    lib.emu_plan_filter.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_uint32, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_double, ctypes.POINTER(ctypes.c_int), ctypes.c_int]
    rng = random.Random(19)
    seen_words = set()
    for it in range(400):
        m = rng.choice([4, 8, 20, 23, 50, 100, 300, 1000])
        k = rng.randrange(0, min(m - 1, 20) + 1)
        q = rand_seq(rng, m)
        buf = (ctypes.c_int * 512)()
        n = lib.emu_plan_filter(0, q, 1, m, k, 1e30, buf, 512)
        enabled, WF, npieces, L = buf[0], buf[1], buf[2], buf[3]
        if not enabled:
            continue
        assert n == 4 + 4 * npieces and npieces == k + 1 and WF in (1, 2, 4, 8)
        seen_words.add(WF)
        pieces = [tuple(buf[4 + 4 * i:8 + 4 * i]) for i in range(npieces)]
        covered = set()
        bits = {}
        for off, ln, word, bit in pieces:
            assert ln >= L >= 1 and 0 <= off and off + ln <= m and 0 <= word < WF
            span = set(range(off, off + ln))
            assert not (span & covered)
            covered |= span
            used = set(range(bit, bit + ln + 3))  # the piece and its delay line
            assert max(used) < 32
            assert not (used & bits.setdefault(word, set()))
            bits[word] |= used
    assert seen_words == {1, 2, 4, 8}


def okey(m):
    return (m.text_start, m.text_end, m.pattern_start, m.pattern_end, m.cost, m.strand, m.cigar)


@pytest.mark.parametrize("alphabet", ["dna", "iupac"])
def test_emu_options_fuzz(alphabet):
    """The post-processing functions the kernels share with the host (end_filter_pass, n_fraction_ok,
    the only_best rule, the untraced record) against the oracle, without a GPU."""
    rng = random.Random(41)
    b = EmuBackend(ltot=256)
    for it in range(150):
        m = rng.choice([4, 8, 20, 33])
        n = rng.randrange(0, 2500)
        k = rng.randrange(0, max(1, m // 4) + 1)
        p, t = planted(rng, m, max(n, 1), k)
        t = bytearray(t[:n])
        if alphabet == "iupac":
            for _ in range(rng.randrange(0, 3)):
                if len(t) > 2:
                    a = rng.randrange(len(t))
                    t[a:a + rng.randrange(1, 25)] = b"N" * min(rng.randrange(1, 25), len(t) - a)
        t = bytes(t[:n])
        opts = dict(without_trace=rng.random() < 0.3, only_best=rng.random() < 0.3,
                    max_n_frac=rng.choice([None, 0.0, 0.2]) if alphabet == "iupac" else None)
        pam = p[-3:] if rng.random() < 0.4 else None
        allm = rng.random() < 0.5
        want = oracle.search(alphabet, p, t, k, rc=True, all_minima=allm, pam=pam, **opts)
        got = b.search_opts(alphabet, p, t, k, rc=True, all_minima=allm, pam=pam, **opts)
        assert list(map(okey, got)) == list(map(okey, want)), (p, t, k, allm, opts, pam)


def test_emu_overhang_fuzz():
    """trace_one_ov, the overhang start state and the edge computation against the oracle."""
    rng = random.Random(42)
    b = EmuBackend(ltot=128)
    for it in range(200):
        m = rng.choice([3, 8, 20, 33, 70])
        n = rng.choice([0, 1, 2, 5, 17, 40, 200, 1500]) if it % 3 else rng.randrange(0, 300)
        k = rng.randrange(0, max(1, m // 3) + 1)
        p, t = planted(rng, m, max(n, 1), k)
        t = t[:n]
        if n >= 4 and rng.random() < 0.7:
            cut = rng.randrange(1, min(m, n))
            t = (p[cut:] + t[len(p) - cut:])[:n] if rng.random() < 0.5 else (t[:n - cut] + p[:cut])[:n]
        alpha = rng.choice([0.0, 0.3, 0.5, 1.0])
        mo = rng.choice([None, None, 0, 2, 5])
        opts = dict(without_trace=rng.random() < 0.2, only_best=rng.random() < 0.2)
        for allm in (False, True):
            want = oracle.search("iupac", p, t, k, rc=True, all_minima=allm, alpha=alpha, max_overhang=mo, **opts)
            got = b.search_opts("iupac", p, t, k, rc=True, all_minima=allm, alpha=alpha, max_overhang=mo, **opts)
            assert list(map(okey, got)) == list(map(okey, want)), (p, t, k, alpha, mo, allm, opts)


def test_emu_encoded_overhang_fuzz():
    """Encoded patterns with overhang (defined by the reference's v2 == v1 fuzz, see oracle)."""
    rng = random.Random(43)
    b = EmuBackend(ltot=128)
    for it in range(60):
        m = rng.choice([5, 12, 23, 40])
        n = rng.randrange(10, 200)
        k = rng.randrange(0, 4)
        pats = [rand_seq(rng, m) for _ in range(rng.randrange(1, 8))]
        t = bytearray(rand_seq(rng, n))
        p = pats[0]
        cut = rng.randrange(1, min(m, n))
        if rng.random() < 0.5:
            t[:m - cut] = p[cut:][:n]
        else:
            t[n - cut:] = p[:cut]
        t = bytes(t[:n])
        for allm in (False, True):
            want = oracle.search_encoded("iupac", pats, t, k, rc=True, all_minima=allm, alpha=0.5)
            got = b.search_encoded("iupac", pats, t, k, rc=True, all_minima=allm, alpha=0.5)
            kk = lambda x: (x.pattern_idx, x.text_start, x.text_end, x.pattern_start, x.pattern_end, x.cost, x.strand, x.cigar)
            assert sorted(map(kk, got)) == sorted(map(kk, want)), (pats, t, k, allm)


def _low_complexity_case(rng, n):
    """A pattern with a low-complexity stretch and a text that holds long repeats of exactly that
    stretch (poly-A, a dinucleotide repeat, a 7-mer satellite) between random sequence, plus planted
    copies of the pattern: the prefilter fires at (almost) every position of the repeats."""
    unit = rng.choice([b"A", b"AC", b"AAT", b"ACGTTGA"])
    m = rng.choice([20, 24, 40, 64, 100])
    k = rng.choice([1, 2, 3]) if m < 40 else rng.choice([2, 4, 8])
    rep = (unit * (m // len(unit) + 1))[:rng.randrange(m // 2, m - 2)]
    p = bytearray(rand_seq(rng, m))
    at = rng.randrange(0, m - len(rep) + 1)
    p[at:at + len(rep)] = rep
    p = bytes(p)
    t = bytearray(rand_seq(rng, n))
    for _ in range(rng.randrange(1, 4)):
        ln = rng.randrange(200, max(400, n // 3))
        pos = rng.randrange(0, max(1, n - ln))
        t[pos:pos + ln] = (unit * (ln // len(unit) + 1))[:ln]
    for _ in range(4):
        q = mutate(rng, p, rng.randrange(0, k + 1))
        pos = rng.randrange(0, max(1, n - len(q)))
        t[pos:pos + len(q)] = q
    return p, bytes(t[:n]), k


@pytest.mark.parametrize("mode", [1, 2, 4])
def test_emu_regional_fallback_low_complexity(mode, monkeypatch):
    """Repeat-rich text: tiles whose hits would cost more to re-scan than the tile itself are scanned
    whole, the rest goes through the prefilter; the union equals the oracle (both prefilter routes,
    both strands)."""
    monkeypatch.setenv("SASSY_EMU_FORCE_REGIONAL", "1")  # the engine takes this pass above `heavy_hits` only
    rng = random.Random(77 + mode)
    dense_seen = 0
    for it in range(24):
        n = rng.randrange(20_000, 70_000)
        p, t, k = _low_complexity_case(rng, n)
        for allm in (False, True):
            want = oracle.search("dna", p, t, k, rc=True, all_minima=allm)
            b = EmuBackend(ltot=rng.choice([64, 128]), use_filter=mode)
            got = b.search("dna", p, t, k, rc=True, all_minima=allm)
            assert list(map(key, got)) == list(map(key, want)), (p, k, allm, b.last_filter, b.last_dense_tiles)
            dense_seen += b.last_dense_tiles > 0
    assert dense_seen >= 8, dense_seen


def test_dna_transport_encoding():
    """Host->device transport of Dna texts (csrc/dna_pack.h): every host packer (scalar, AVX2,
    AVX-512 + GFNI where the CPU has them) writes the same bytes, the device-side decoder returns
    the canonical upper-case text, and a byte outside ACGTacgt is reported wherever it sits."""
    import ctypes
    from tests.emu_backend import _lib
    lib = _lib()
    lib.emu_dna_pack.restype = ctypes.c_int
    lib.emu_dna_pack.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p]
    lib.emu_dna_unpack.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p]
    best = lib.emu_dna_pack_best()
    rng = random.Random(77)
    for n in [0, 1, 7, 8, 9, 31, 32, 33, 63, 64, 255, 256, 257, 511, 1000, 4096 + 5, 70001]:
        text = bytes(rng.choice(b"ACGTacgt") for _ in range(n))
        size = (n + 7) // 8 * 2
        ref = None
        for level in range(best + 1):
            # an unaligned destination exercises the non-streaming store of the widest packer
            for shift in (0, 2):
                dst = ctypes.create_string_buffer(size + 64 + shift)
                buf = (ctypes.c_char * (size + 64)).from_buffer(dst, shift)
                assert lib.emu_dna_pack(level, text, n, ctypes.cast(buf, ctypes.c_char_p)) == 1
                got = bytes(buf[:size])
                if ref is None:
                    ref = got
                assert got == ref, (n, level, shift)
        out = ctypes.create_string_buffer(max(n, 1))
        lib.emu_dna_unpack(ref, n, out)
        assert out.raw[:n] == text.upper(), n
        if n:
            for bad in (b"N", b"\xc1", b"B", b"\x00", b"U"):
                pos = rng.randrange(n)
                t2 = text[:pos] + bad + text[pos + 1:]
                for level in range(best + 1):
                    dst = ctypes.create_string_buffer(size + 64)
                    assert lib.emu_dna_pack(level, t2, n, dst) == 0, (n, level, bad, pos)


def test_concat_locate():
    """Many texts scanned as one concatenated text (Engine::search_texts): a candidate's end position
    in the concatenation maps to (text, end position in the text) in scan direction, positions in the
    padding between texts map to nothing; empty texts never own a position."""
    import ctypes
    from tests.emu_backend import _lib
    lib = _lib()
    lib.emu_concat_locate.restype = ctypes.c_int
    lib.emu_concat_locate.argtypes = [ctypes.c_uint64, ctypes.c_int, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint64),
                                      ctypes.POINTER(ctypes.c_uint64), ctypes.c_uint32,
                                      ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint64)]
    rng = random.Random(91)
    for _ in range(30):
        nt = rng.randrange(1, 40)
        lens = [rng.choice([0, 0, 1, 5, 15, 16, 17, 100]) for _ in range(nt)]
        offs, total = [], 0
        for ln in lens:  # as Engine::search_texts: every text starts on a 16-byte boundary
            offs.append(total)
            total += (ln + 15) & ~15
        if total == 0:
            continue
        owner = [None] * total  # forward index -> (text, index in the text)
        for t, (o, ln) in enumerate(zip(offs, lens)):
            for i in range(ln):
                owner[o + i] = (t, i)
        c_offs = (ctypes.c_uint64 * nt)(*offs)
        c_lens = (ctypes.c_uint64 * nt)(*lens)
        ti, local = ctypes.c_uint32(0), ctypes.c_uint64(0)
        for pos in range(0, total + 2):
            for rev in (0, 1):
                ok = lib.emu_concat_locate(pos, rev, total, c_offs, c_lens, nt, ctypes.byref(ti), ctypes.byref(local))
                g = (total - pos) if rev else (pos - 1)  # forward index of the last character consumed
                want = owner[g] if 1 <= pos <= total else None
                if want is None:
                    assert ok == 0, (pos, rev)
                else:
                    t, i = want
                    assert ok == 1 and ti.value == t, (pos, rev, offs, lens)
                    assert local.value == ((lens[t] - i) if rev else (i + 1)), (pos, rev)
