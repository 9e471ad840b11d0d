"""Oracle under the Searcher options (without_trace, only_best_match, max_n_frac, PAM end filter,
search_many), pinned against the known answers the reference's own tests hold for them."""
import random

import oracle
from oracle import USIZE_MAX
from tests.test_oracle_props import planted, rand_seq


def test_n_filter_complex_example():
    """reference src/n_filter.rs:84-106 (n_filter_complex_example)."""
    p = b"ACGTACGTACGT"
    t = b"NNNNNNNNNNNNNAAAAAAAAAAAAAAAAAANNNNNNNGTACGT"
    assert [m.text_end for m in oracle.search("iupac", p, t, 1, all_minima=True)] == [11, 12, 13, 14, 43, 44]
    got = oracle.search("iupac", p, t, 1, all_minima=True, max_n_frac=0.5)
    assert [m.text_end for m in got] == [44]


def test_filter_fn_simple_as_pam():
    """reference src/search.rs:2545-2564 (test_filter_fn_simple): two exact copies; without a filter
    both are found at 10 and 50.  The PAM form of the end filter keeps both (the last characters
    of the pattern match themselves) and rejects a PAM that does not occur."""
    p = b"ATCGATCA"
    t = bytearray(b"G" * 100)
    t[10:10] = p
    t[50:50] = p
    t = bytes(t)
    assert [m.text_start for m in oracle.search("dna", p, t, 0)] == [10, 50]
    assert [m.text_start for m in oracle.search("dna", p, t, 0, pam=b"TCA")] == [10, 50]
    assert oracle.search("dna", p, t, 0, pam=b"TCC") == []


def test_filter_fn_rc_as_pam():
    """reference src/search.rs:2583-2607 (test_filter_fn_rc): the closure sees the complemented
    text on the rc strand; both the forward copy at 10 and the reverse-complement copy at 50
    pass a filter that compares with the pattern's own suffix."""
    p = b"ATCGATCA"
    t = bytearray(b"G" * 100)
    t[10:10] = p
    t[50:50] = oracle.reverse_complement("dna", p)
    ms = oracle.search("dna", p, bytes(t), 0, rc=True, pam=p)
    assert [(m.text_start, m.strand) for m in ms] == [(10, "+"), (50, "-")]


def test_without_trace_fields():
    """reference src/search.rs:1464-1475 (untraced record) and :859-872 (rc mapping)."""
    p = b"ATCGATCA"
    t = bytearray(b"G" * 100)
    t[10:10] = p
    t[58:58] = oracle.reverse_complement("dna", p)
    ms = oracle.search("dna", p, bytes(t), 0, rc=True, without_trace=True)
    assert [(m.text_start, m.text_end, m.pattern_start, m.pattern_end, m.cost, m.strand, m.cigar) for m in ms] == [
        (USIZE_MAX, 18, USIZE_MAX, 8, 0, "+", ""), (58, USIZE_MAX, USIZE_MAX, 8, 0, "-", "")]


def test_only_best_is_rightmost_minimum():
    rng = random.Random(5)
    for _ in range(200):
        m = rng.randrange(4, 30)
        n = rng.randrange(1, 400)
        k = rng.randrange(0, m // 3 + 1)
        p, t = planted(rng, m, max(n, m + k + 2), k)
        t = t[:n]
        for allm in (False, True):
            full = oracle.search("dna", p, t, k, rc=True, all_minima=allm)
            best = oracle.search("dna", p, t, k, rc=True, all_minima=allm, only_best=True)
            for strand in "+-":
                f = [x for x in full if x.strand == strand]
                b = [x for x in best if x.strand == strand]
                if not f:
                    assert not b
                    continue
                c = min(x.cost for x in f)
                # rightmost in scan direction: largest end on "+", smallest start on "-"
                want = max((x for x in f if x.cost == c), key=lambda x: x.text_end if strand == "+" else -x.text_start)
                assert b == [want]


def test_search_many_equals_single_searches():
    rng = random.Random(6)
    for _ in range(30):
        pats = [rand_seq(rng, rng.choice([5, 5, 9, 20])) for _ in range(rng.randrange(1, 6))]
        texts = [rand_seq(rng, rng.randrange(0, 300)) for _ in range(rng.randrange(1, 6))]
        k = rng.randrange(0, 3)
        got = oracle.search_many("dna", pats, texts, k, rc=True)
        want = []
        for pi, p in enumerate(pats):
            for ti, t in enumerate(texts):
                for x in oracle.search("dna", p, t, k, rc=True):
                    want.append((pi, ti, x.text_start, x.text_end, x.cost, x.strand, x.cigar))
        assert [(x.pattern_idx, x.text_idx, x.text_start, x.text_end, x.cost, x.strand, x.cigar) for x in got] == want
