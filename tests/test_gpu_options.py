"""CUDA path under the Searcher options and the multi-text / multi-pattern entry points,
against the oracle.  Needs a B200."""
import random

import pytest

import oracle
from tests.test_oracle_props import planted, rand_seq

pytestmark = pytest.mark.gpu


def key(m):
    return (m.pattern_idx, m.text_idx, m.text_start, m.text_end, m.pattern_start, m.pattern_end, m.cost, m.strand,
            m.cigar)


def mk(alphabet, rc, filt="auto", **opts):
    import sassy_b200
    s = sassy_b200.Searcher(alphabet, rc=rc, max_n_frac=opts.get("max_n_frac"))
    s.set_filter(filt)
    if opts.get("without_trace"):
        s.without_trace()
    if opts.get("only_best"):
        s.only_best_match()
    return s


def noisy(rng, t, alphabet):
    """Sprinkle N runs (and IUPAC codes) into a text."""
    t = bytearray(t)
    for _ in range(rng.randrange(0, 4)):
        if len(t) < 2:
            break
        a = rng.randrange(len(t))
        for i in range(a, min(len(t), a + rng.randrange(1, 30))):
            t[i] = ord("N")
    return bytes(t)


def test_reference_kats_for_options():
    s = mk("iupac", False, max_n_frac=0.5)
    p = b"ACGTACGTACGT"
    t = b"NNNNNNNNNNNNNAAAAAAAAAAAAAAAAAANNNNNNNGTACGT"
    assert [m.text_end for m in s.search_all(p, t, 1)] == [44]            # src/n_filter.rs:84-106
    s.without_max_n_frac()
    assert [m.text_end for m in s.search_all(p, t, 1)] == [11, 12, 13, 14, 43, 44]
    d = mk("dna", True)
    p = b"ATCGATCA"
    t = bytearray(b"G" * 100)
    t[10:10] = p
    t[50:50] = oracle.reverse_complement("dna", p)
    ms = d.search_with_pam(p, bytes(t), 0, p, all_minima=False)             # src/search.rs:2583-2607
    assert [(m.text_start, m.strand) for m in ms] == [(10, "+"), (50, "-")]


@pytest.mark.parametrize("filt", ["off", "force"])
@pytest.mark.parametrize("alphabet", ["dna", "iupac"])
def test_gpu_options_fuzz(alphabet, filt):
    rng = random.Random(31)
    searchers = {}
    for it in range(150):
        m = rng.choice([4, 8, 20, 23, 33, 70])
        n = rng.randrange(0, 6000)
        k = rng.randrange(0, max(1, m // 4) + 1)
        p, t = planted(rng, m, max(n, 1), k)
        t = t[:n]
        if alphabet == "iupac":
            t = noisy(rng, t, alphabet)
        opts = dict(without_trace=rng.random() < 0.3, only_best=rng.random() < 0.3,
                    max_n_frac=rng.choice([None, None, 0.0, 0.1, 0.5]) if alphabet == "iupac" else None)
        pam = p[-3:] if rng.random() < 0.4 else None
        allm = rng.random() < 0.5
        okey = (opts["without_trace"], opts["only_best"], opts["max_n_frac"])
        if okey not in searchers:
            searchers[okey] = mk(alphabet, True, filt, **opts)
        s = searchers[okey]
        want = oracle.search(alphabet, p, t, k, rc=True, all_minima=allm, pam=pam, **opts)
        if pam is not None:
            got = s.search_with_pam(p, t, k, pam, all_minima=allm)
        else:
            got = s.search_all(p, t, k) if allm else s.search(p, t, k)
        assert list(map(key, got)) == list(map(key, want)), (alphabet, p, t, k, allm, opts, pam)


def test_gpu_encoded_max_n_frac():
    rng = random.Random(32)
    s = mk("iupac", True, max_n_frac=0.2)
    for it in range(30):
        m = rng.choice([8, 23, 32])
        n = rng.randrange(1, 5000)
        k = rng.randrange(0, m // 4 + 1)
        pats = []
        t = bytearray(rand_seq(rng, n))
        for _ in range(rng.randrange(1, 12)):
            p, tt = planted(rng, m, n, k)
            pats.append(p)
            a = rng.randrange(0, max(1, n - m))
            t[a:a + m] = p[:max(0, min(m, n - a))]
        t = noisy(rng, bytes(t[:n]), "iupac")
        enc = s.encode_patterns(pats)
        for allm in (False, True):
            want = oracle.search_encoded("iupac", pats, t, k, rc=True, all_minima=allm, max_n_frac=0.2)
            got = s.search_all_encoded_patterns(enc, t, k) if allm else s.search_encoded_patterns(enc, t, k)
            kk = lambda x: (x.pattern_idx, x.text_start, x.text_end, x.cost, x.strand, x.cigar)
            assert sorted(map(kk, got)) == sorted(map(kk, want))


@pytest.mark.parametrize("concat", ["0", "2"], ids=["thread-per-pair", "concatenated"])
@pytest.mark.parametrize("alphabet", ["dna", "iupac"])
def test_gpu_search_many_fuzz(alphabet, concat, monkeypatch):
    """Mirrors the reference's search_many_fuzz (src/search.rs:3624-3730): random pattern and
    text sets, all three modes, against pattern-by-text single searches (the oracle).  Both
    many-text routes of Engine::search_texts: one thread per (text, query) pair, and the texts
    scanned as one concatenated text by the row-tiled kernels."""
    monkeypatch.setenv("SASSY_B200_TEXTS_CONCAT", concat)
    rng = random.Random(33)
    s = mk(alphabet, True)
    fwd = mk(alphabet, False)
    for it in range(40):
        np_, nt = rng.randrange(1, 30), rng.randrange(1, 40)
        m = rng.randrange(1, 100)
        pats = [rand_seq(rng, m) for _ in range(np_)]
        texts = [rand_seq(rng, rng.randrange(2, 1000)) for _ in range(nt)]
        for i in range(0, nt, 3):  # make sure there is something to find
            p = pats[rng.randrange(np_)]
            a = rng.randrange(0, max(1, len(texts[i])))
            texts[i] = (texts[i][:a] + p + texts[i][a:])[:1000]
        if it % 7 == 0:
            texts.append(b"")
        k = rng.randrange(0, m * 4 // 10 + 1)
        rc = rng.random() < 0.5
        srch = s if rc else fwd
        want = oracle.search_many(alphabet, pats, texts, k, rc=rc)
        for mode in ("single", "batch_patterns", "batch_texts"):
            got = srch.search_many(pats, texts, k, 0, mode)
            assert list(map(key, got)) == list(map(key, want)), (it, mode, m, k, rc)


@pytest.mark.parametrize("alphabet", ["dna", "iupac", "ascii"])
def test_gpu_search_many_reads(alphabet):
    """The nanopore-barcode shape at a size that takes the concatenated route by itself (> 1 MiB of
    reads): barcodes planted at the very start and the very end of reads (where the state carried
    over from the neighbouring read must not leak in), reads shorter than m + k, empty reads, a
    read that ends with the start of a barcode and is followed by one that continues it."""
    rng = random.Random(35)
    rc = alphabet != "ascii"
    s = mk(alphabet, rc)
    m, k = 24, 3
    sigma = b"ACGT" if alphabet != "ascii" else b"abcdefghijklmnop"
    seq = lambda n: bytes(rng.choice(sigma) for _ in range(n))
    pats = [seq(m) for _ in range(24)]
    reads = []
    total = 0
    while total < (1 << 20) + 50_000:
        ln = rng.choice([0, 5, 20, 27, 30, 64]) if rng.random() < 0.1 else rng.randrange(200, 3000)
        r = bytearray(seq(ln))
        if ln >= 100 and rng.random() < 0.6:
            bc = bytearray(rng.choice(pats))
            for _ in range(rng.randrange(0, k + 1)):
                bc[rng.randrange(m)] = rng.choice(sigma)
            where = rng.choice(["start", "end", "mid"])
            pos = 0 if where == "start" else (ln - m if where == "end" else rng.randrange(0, ln - m))
            pos = max(0, min(ln - m, pos + rng.randrange(-2, 3)))
            r[pos:pos + m] = bc
        reads.append(bytes(r))
        total += ln
    # a barcode split over two neighbouring reads must not be found
    p0 = pats[0]
    reads.append(seq(500) + p0[:12])
    reads.append(p0[12:] + seq(500))
    want = oracle.search_many(alphabet, pats, reads, k, rc=rc)
    got = s.search_many(pats, reads, k)
    assert list(map(key, got)) == list(map(key, want))
    assert len(want) > 100
    assert s.stats()["swar_lanes"] == 2  # the two-pattern row-tiled scan ran


def test_gpu_search_many_mixed_lengths_and_long_text():
    rng = random.Random(34)
    s = mk("dna", True)
    pats = [rand_seq(rng, m) for m in (12, 30, 12, 20, 30)]
    long_text = bytearray(rand_seq(rng, 400_000))
    for i, p in enumerate(pats):
        long_text[50_000 * (i + 1):50_000 * (i + 1) + len(p)] = p
    texts = [rand_seq(rng, 300), bytes(long_text), pats[1] + rand_seq(rng, 50) + pats[3]]
    want = oracle.search_many("dna", pats, texts, 2, rc=True)
    got = s.search_many(pats, texts, 2)
    assert list(map(key, got)) == list(map(key, want))
    assert len(got) >= 7


def test_gpu_search_texts_and_patterns():
    rng = random.Random(35)
    s = mk("dna", True)
    for opts in ({}, {"only_best": True}, {"without_trace": True}):
        so = mk("dna", True, **opts)
        p = rand_seq(rng, 24)
        texts = []
        for i in range(200):
            t = bytearray(rand_seq(rng, rng.randrange(0, 400)))
            if i % 2 and len(t) > 30:
                a = rng.randrange(0, len(t) - 24)
                q = bytearray(p)
                q[rng.randrange(24)] = ord("A")
                t[a:a + 24] = q
            texts.append(bytes(t))
        want = oracle.search_many("dna", [p], texts, 3, rc=True, **opts)
        got = so.search_texts(p, texts, 3)
        assert list(map(key, got)) == list(map(key, want)), opts
        assert len(got) >= 50
        pats = [rand_seq(rng, 16) for _ in range(37)]
        t = bytearray(rand_seq(rng, 30_000))
        for i, q in enumerate(pats):
            t[700 * i + 5:700 * i + 21] = q
        want = oracle.search_many("dna", pats, [bytes(t)], 2, rc=True, **opts)
        got = so.search_patterns(pats, bytes(t), 2)
        assert list(map(key, got)) == list(map(key, want)), opts


def okey(m):
    return (m.pattern_idx, m.text_idx, m.text_start, m.text_end, m.pattern_start, m.pattern_end, m.cost, m.strand,
            m.cigar)


def test_gpu_overhang_reference_kats():
    import sassy_b200
    s = sassy_b200.Searcher("iupac", rc=False, alpha=0.5)
    ms = s.search(b"GGGGAAAA", b"TTTTTTTTTTTTTTTTGGGG", 4)                      # src/search.rs:2430-2456
    assert len(ms) == 2 and any(m.text_end == 20 and m.pattern_end == 3 and m.cost == 2 for m in ms)
    m = s.search(b"ATCGATCG", b"ATCGGGGGGGGGG", 2)[0]                             # :2929-2942
    assert (m.pattern_start, m.pattern_end, m.text_start, m.text_end, m.cost, m.cigar) == (4, 8, 0, 4, 2, "4=")
    m = s.search(b"ATCGATCG", b"GGGGGGGATCG", 2)[0]                               # :2944-2958
    assert (m.pattern_start, m.pattern_end, m.text_start, m.text_end, m.cost, m.cigar) == (0, 4, 7, 11, 2, "4=")
    s.set_max_n_frac(0.0)
    assert len(s.search_all(b"AAAA", b"GGGGGG", 2)) == 4                          # src/n_filter.rs:66-82
    with pytest.raises(ValueError):
        sassy_b200.Searcher("dna", alpha=0.5)                                     # src/search.rs:2344-2348


def test_gpu_overhang_fuzz():
    import sassy_b200
    rng = random.Random(36)
    cache = {}
    for it in range(200):
        m = rng.choice([3, 8, 20, 33, 70])
        n = rng.choice([0, 1, 2, 5, 17, 40, 200, 3000, 70_000]) if it % 3 else rng.randrange(0, 400)
        k = rng.randrange(0, max(1, m // 3) + 1)
        p, t = planted(rng, m, max(n, 1), k)
        t = t[:n]
        # pattern prefixes / suffixes at the text ends: the overhang cases
        if n >= 4 and rng.random() < 0.7:
            cut = rng.randrange(1, min(m, n))
            t = (p[cut:] + t[len(p) - cut:])[:n] if rng.random() < 0.5 else (t[:n - cut] + p[:cut])[:n]
        alpha = rng.choice([0.0, 0.3, 0.5, 1.0])
        mo = rng.choice([None, None, 0, 2, 5])
        opts = dict(without_trace=rng.random() < 0.25, only_best=rng.random() < 0.25,
                    max_n_frac=rng.choice([None, None, 0.2]))
        rc = rng.random() < 0.7
        ck = (alpha, rc, opts["without_trace"], opts["only_best"], opts["max_n_frac"])
        if ck not in cache:
            s = sassy_b200.Searcher("iupac", rc=rc, alpha=alpha, max_n_frac=opts["max_n_frac"])
            if opts["without_trace"]:
                s.without_trace()
            if opts["only_best"]:
                s.only_best_match()
            cache[ck] = s
        s = cache[ck].with_max_overhang(mo)
        for allm in (False, True):
            want = oracle.search("iupac", p, t, k, rc=rc, all_minima=allm, alpha=alpha, max_overhang=mo, **opts)
            got = s.search_all(p, t, k) if allm else s.search(p, t, k)
            assert list(map(okey, got)) == list(map(okey, want)), (p, t, k, alpha, mo, rc, allm, opts)


def test_gpu_overhang_search_many():
    import sassy_b200
    rng = random.Random(37)
    s = sassy_b200.Searcher("iupac", rc=True, alpha=0.5)
    for it in range(15):
        m = rng.randrange(4, 40)
        pats = [rand_seq(rng, m) for _ in range(rng.randrange(1, 8))]
        texts = []
        for _ in range(rng.randrange(1, 30)):
            t = rand_seq(rng, rng.randrange(0, 300))
            p = pats[rng.randrange(len(pats))]
            cut = rng.randrange(1, m)
            texts.append((p[cut:] + t) if rng.random() < 0.5 else (t + p[:cut]))
        k = rng.randrange(0, m // 3 + 1)
        want = oracle.search_many("iupac", pats, texts, k, rc=True, alpha=0.5)
        got = s.search_many(pats, texts, k)
        assert list(map(okey, got)) == list(map(okey, want)), (it, m, k)


def test_reference_c_abi_with_alpha():
    """The reference's own four symbols (c/sassy.h:38-63) with overhang: sassy_searcher(alphabet, rc,
    alpha) + search, records with pattern_start / pattern_end inside the pattern."""
    import ctypes
    from sassy_b200 import _native
    lib = _native.load()
    h = lib.sassy_searcher(b"iupac", False, 0.5)
    assert h
    out = ctypes.POINTER(_native.CMatch)()
    p, t = b"ATCGATCG", b"ATCGGGGGGGGGG"
    n = lib.search(h, p, len(p), t, len(t), 2, ctypes.byref(out))
    want = oracle.search("iupac", p, t, 2, alpha=0.5)
    got = [(out[i].text_start, out[i].text_end, out[i].pattern_start, out[i].pattern_end, out[i].cost, out[i].strand)
           for i in range(n)]
    assert got == [(w.text_start, w.text_end, w.pattern_start, w.pattern_end, w.cost, 0) for w in want]
    assert got[0] == (0, 4, 4, 8, 2, 0)  # src/search.rs:2929-2942
    lib.sassy_matches_free(out, n)
    lib.sassy_searcher_free(h)


def test_c_program_against_the_library(tmp_path):
    """The doctest of the reference (src/lib.rs:62-107: ATCG in CCCATCACCC, k = 1) from a plain C program."""
    import subprocess
    from tests.test_c_abi import build_c_caller
    out = subprocess.check_output([build_c_caller(tmp_path), "dna", "ATCG", "CCCATCACCC", "1"]).decode().splitlines()
    assert out == ["3 7 0 4 1 0", "1 5 0 4 1 1", "cigar 3=1X", "cigar 2=1X1="]


def test_gpu_encoded_overhang():
    """Encoded patterns with overhang: equal to the forward v1 overhang search of every query, as the
    reference's fuzz_against_sassy_batch requires (src/pattern_tiling/search.rs:690-848,886-896)."""
    import sassy_b200
    rng = random.Random(38)
    s = sassy_b200.Searcher("iupac", rc=True, alpha=0.5)
    kk = lambda x: (x.pattern_idx, x.text_start, x.text_end, x.pattern_start, x.pattern_end, x.cost, x.strand, x.cigar)
    for it in range(40):
        m = rng.choice([5, 12, 23, 40, 59])
        n = rng.choice([10, 30, 59, 300, 5000])
        k = rng.randrange(0, 4)
        pats = [rand_seq(rng, m) for _ in range(rng.randrange(1, 26))]
        t = bytearray(rand_seq(rng, n))
        cut = rng.randrange(1, min(m, n))
        if rng.random() < 0.5:
            t[:m - cut] = pats[0][cut:][:n]
        else:
            t[n - cut:] = pats[0][:cut]
        t = bytes(t[:n])
        enc = s.encode_patterns(pats)
        for allm in (False, True):
            want = oracle.search_encoded("iupac", pats, t, k, rc=True, all_minima=allm, alpha=0.5)
            got = s.search_all_encoded_patterns(enc, t, k) if allm else s.search_encoded_patterns(enc, t, k)
            assert sorted(map(kk, got)) == sorted(map(kk, want)), (it, m, n, k, allm)


def test_two_searchers_on_two_host_threads():
    """The reference's documented usage is one searcher per worker thread (bin/grep.rs:488-498).
    Two threads, each with its own searchers (different pattern lengths -> different kernel
    instantiations and shared-memory limits), search concurrently; every result equals the
    oracle's.  Exercises the guarded launch-configuration caches (scan_kernels.cu)."""
    import threading
    import random
    import oracle
    import sassy_b200
    rng = random.Random(91)
    n = 300_000
    base = bytes(rng.choice(b"ACGT") for _ in range(n))
    jobs = []
    for m, k, alphabet in [(20, 2, "dna"), (100, 8, "dna"), (23, 3, "iupac"), (64, 4, "dna"), (200, 6, "iupac"), (12, 1, "dna")]:
        p = bytes(rng.choice(b"ACGT") for _ in range(m))
        t = bytearray(base)
        for _ in range(5):
            pos = rng.randrange(0, n - m)
            t[pos:pos + m] = p
        t = bytes(t)
        want = [(x.text_start, x.text_end, x.cost, x.strand, x.cigar) for x in oracle.search(alphabet, p, t, k, rc=True)]
        jobs.append((alphabet, p, t, k, want))
    errors = []

    def worker(my_jobs):
        try:
            for rep in range(3):
                for alphabet, p, t, k, want in my_jobs:
                    s = sassy_b200.Searcher(alphabet, rc=True)
                    got = [(x.text_start, x.text_end, x.cost, x.strand, x.cigar) for x in s.search(p, t, k)]
                    if got != want:
                        errors.append((alphabet, len(p), k, len(got), len(want)))
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(jobs[i::2],)) for i in range(2)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
