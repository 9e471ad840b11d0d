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


# ---- overhang (alpha): known answers of the reference's own tests ------------------------------

def _ends(ms):
    """(end position incl. overshoot, cost) of untraced matches: end = text_end + (m - pattern_end)."""
    return [(m.text_end, m.pattern_end, m.cost) for m in ms]


def test_overhang_reference_kats():
    S = lambda p, t, k, **kw: oracle.search("iupac", p, t, k, alpha=0.5, **kw)
    # src/search.rs:2372-2398 overshoot_simple_prefix: end position 3 with cost <= 2
    assert (3, 8, 2) in _ends(S(b"AAAAGGGG", b"GGGGTTTTTTTTTTTTTTTT", 2, all_minima=True, without_trace=True))
    # :2400-2428 overshoot_simple_suffix: end position 24 = text end 20 + 4 pattern characters beyond it
    assert (20, 4, 2) in _ends(S(b"GGGGAAAA", b"TTTTTTTTTTTTTTTTGGGG", 2, all_minima=True, without_trace=True))
    # :2430-2456 overshoot_simple_suffix_local_minima
    ms = S(b"GGGGAAAA", b"TTTTTTTTTTTTTTTTGGGG", 4)
    assert len(ms) == 2 and any(m.text_end == 20 and m.pattern_end == 3 and m.cost == 2 for m in ms)
    # :2458-2490 overshoot_test_prefix_and_suffix: ends 3 and 13, cost 2 each
    e = _ends(S(b"AAAAGGGG", b"GGGGGAAAAA", 2, all_minima=True, without_trace=True))
    assert (3, 8, 2) in e and (10, 5, 2) in e
    # :2929-2942 test_pattern_trace_path_with_overhang_prefix: path (4,0) (5,1) (6,2) (7,3)
    m = S(b"ATCGATCG", b"ATCGGGGGGGGGG", 2)[0]
    assert (m.pattern_start, m.pattern_end, m.text_start, m.text_end, m.cost, m.cigar) == (4, 8, 0, 4, 2, "4=")
    # :2944-2958 ..._suffix: path (0,7) (1,8) (2,9) (3,10)
    m = S(b"ATCGATCG", b"GGGGGGGATCG", 2)[0]
    assert (m.pattern_start, m.pattern_end, m.text_start, m.text_end, m.cost, m.cigar) == (0, 4, 7, 11, 2, "4=")
    # :3022-3058 test_case4: a match ending at 1 with cost 1, in both modes
    assert any(m.text_end == 1 and m.cost == 1 for m in S(b"ATC", b"CGGGGGG", 3))
    assert any(m.text_end == 1 and m.cost == 1 for m in S(b"ATC", b"CGGGGGG", 3, all_minima=True))
    # src/n_filter.rs:66-82 n_filter_full_overhang_match: the overhang's wildcards do not count as N
    assert len(S(b"AAAA", b"GGGGGG", 2, all_minima=True, max_n_frac=0.0)) == 4
    # src/search.rs:2344-2348: Dna has no overhang
    try:
        oracle.search("dna", b"ACGT", b"ACGT", 1, alpha=0.5)
    except oracle.OracleError:
        pass
    else:
        raise AssertionError("overhang must be rejected for dna")


def test_overhang_is_plain_search_away_from_the_ends():
    """Matches whose alignment stays m+k characters away from both text ends do not change."""
    rng = random.Random(9)
    for _ in range(60):
        m = rng.randrange(6, 30)
        k = rng.randrange(0, m // 3 + 1)
        n = rng.randrange(4 * m, 600)
        p, t = planted(rng, m, n, k)
        t = t[:n]
        a = oracle.search("iupac", p, t, k, rc=True, all_minima=True)
        b = oracle.search("iupac", p, t, k, rc=True, all_minima=True, alpha=rng.choice([0.0, 0.5, 1.0]))
        inner = lambda ms: [x for x in ms if x.text_start > m + k and x.text_end < n - (m + k) and x.pattern_start == 0
                            and x.pattern_end == m]
        assert inner(a) == inner(b)


def test_concatenated_texts_argument():
    """The argument behind the many-text route of the GPU engine (Engine::search_texts):
    scanning texts as ONE concatenation gives, beyond the first m + k end positions of a text (in scan
    direction), exactly the end positions and costs <= k of a search of that text alone -- whatever
    precedes it -- and the first m + k end positions only need the text's own first m + k characters."""
    import random
    from tests.test_oracle_props import mutate, rand_seq
    rng = random.Random(123)
    seen_deep, seen_head, seen_rc = 0, 0, 0
    for alphabet in ("dna", "iupac"):
        for it in range(25):
            m = rng.randrange(4, 40)
            k = rng.randrange(0, max(1, m // 3))
            p = rand_seq(rng, m)
            texts = []
            for _ in range(rng.randrange(2, 9)):
                ln = rng.choice([0, 1, m - 1, m + k, m + k + 1, 3 * m, 200])
                t = bytearray(rand_seq(rng, ln))
                if ln >= m and rng.random() < 0.8:  # a copy at the very start / end / somewhere
                    q = mutate(rng, p, rng.randrange(0, k + 1))[:ln]
                    pos = rng.choice([0, max(0, ln - len(q)), rng.randrange(0, ln - len(q) + 1)])
                    t[pos:pos + len(q)] = q
                texts.append(bytes(t))
            # as the engine stages them: every text on a 16-byte boundary, zero bytes in between
            offs, cat = [], bytearray()
            for t in texts:
                offs.append(len(cat))
                cat += t + bytes((-len(t)) % 16)
            cat = bytes(cat)
            total = len(cat)
            if total == 0:
                continue

            def ends(ms, n):  # (strand, end position in scan direction, cost)
                return {(x.strand, x.text_end if x.strand == "+" else n - x.text_start, x.cost) for x in ms}

            # (costs and end positions only: the padding bytes are outside the alphabet, a traceback
            #  over them is the reference's "Trace failed" panic; the engine traces per text)
            whole = ends(oracle.search(alphabet, p, cat, k, rc=True, all_minima=True, without_trace=True), total)
            for t, off in zip(texts, offs):
                n = len(t)
                own = ends(oracle.search(alphabet, p, t, k, rc=True, all_minima=True, without_trace=True), n)
                # the concatenation's candidates that fall into this text, in text coordinates
                got = set()
                for strand, e, c in whole:
                    g = e - 1 if strand == "+" else total - e  # forward index of the last character consumed
                    if off <= g < off + n:
                        local = g - off + 1 if strand == "+" else off + n - g
                        if local > m + k:
                            got.add((strand, local, c))
                assert got == {x for x in own if x[1] > m + k}, (alphabet, p, texts, k)
                seen_deep += len(got)
                seen_rc += sum(1 for x in got if x[0] == "-")
                # the first m + k end positions from the text's own first (scan direction) m + k characters
                head_f = ends(oracle.search(alphabet, p, t[:m + k], k, rc=False, all_minima=True, without_trace=True),
                              min(n, m + k))
                want_f = {x for x in own if x[0] == "+" and x[1] <= m + k}
                assert {x for x in head_f if x[1] <= m + k} == want_f, (alphabet, p, t, k)
                seen_head += len(want_f)
    assert seen_deep > 100 and seen_head > 20, (seen_deep, seen_head, seen_rc)
