"""The CPU baseline (oracle/sassy_cpu_port.c, a SIMD restatement of the reference's v1 engine)
must find exactly the oracle's end positions before its timings mean anything."""
import random

import pytest

import oracle
from oracle import cpu_port
from tests.test_oracle_props import planted


def ends_of(ms, n):
    out = []
    for m in ms:
        if m.strand == "+":
            out.append((m.text_end, m.cost, 0))
        else:
            out.append((n - m.text_start, m.cost, 1))
    return out


@pytest.mark.parametrize("alphabet", ["dna", "iupac"])
def test_cpu_port_equals_oracle(alphabet):
    rng = random.Random(31)
    for it in range(120):
        m = rng.choice([1, 3, 8, 20, 23, 40, 64, 65, 100, 150])
        n = rng.randrange(0, 3000)
        k = rng.randrange(0, max(1, m // 3) + 1)
        p, t = planted(rng, m, max(n, 1), k)
        t = t[:n]
        if alphabet == "iupac" and rng.random() < 0.3:
            t = bytes(c if rng.random() > 0.05 else ord(rng.choice("NRYSWKM")) for c in t)
        if rng.random() < 0.1:
            t, p = b"A" * n, b"A" * m
        for allm in (False, True):
            want = ends_of(oracle.search(alphabet, p, t, k, rc=True, all_minima=allm), n)
            got, _ = cpu_port.search_ends(alphabet, p, t, n, k, True, allm, threads=1)
            assert got == want, (alphabet, p, t, k, allm)


def test_cpu_port_threads():
    rng = random.Random(32)
    m, k, n = 20, 2, 1 << 20
    p, t = planted(rng, m, n, k)
    t = bytearray(t)
    for i in range(50):
        pos = rng.randrange(0, n - m)
        t[pos:pos + m] = p
    t = bytes(t)
    one, _ = cpu_port.search_ends("dna", p, t, n, k, True, False, threads=1)
    four, _ = cpu_port.search_ends("dna", p, t, n, k, True, False, threads=4)
    assert one == four and len(one) >= 50
    want = ends_of(oracle.search("dna", p, t, k, rc=True), n)
    assert one == want


@pytest.mark.parametrize("prefilter", [True, False])
def test_cpu_port_v2_equals_oracle(prefilter):
    """The pattern-tiled CPU baseline (u32 lanes, u16 suffix prefilter for 1 <= k <= 3) reports the
    oracle's end positions, costs and traced start positions of search_encoded_patterns."""
    rng = random.Random(41)
    for it in range(60):
        m = rng.choice([17, 20, 23, 32, 8, 16])
        n = rng.randrange(1, 3000)
        k = rng.randrange(0, 5)
        P = rng.randrange(1, 70)
        pats = []
        t = bytearray(rand_text(rng, n))
        for _ in range(P):
            p = bytes(rng.choice(b"ACGT") for _ in range(m - 3)) + (b"NGG" if rng.random() < 0.7 else b"ACG")
            pats.append(p)
            q = bytearray(p.replace(b"N", b"A"))
            for _ in range(rng.randrange(0, k + 1)):
                q[rng.randrange(len(q))] = rng.choice(b"ACGT")
            pos = rng.randrange(0, max(1, n - m))
            t[pos:pos + len(q)] = q
        if rng.random() < 0.3:
            t = bytearray(c if rng.random() > 0.03 else ord(rng.choice("NRYK")) for c in t)
        t = bytes(t[:n])
        for allm in (False, True):
            want = sorted((x.pattern_idx, x.text_end, x.cost, x.text_start)
                          for x in oracle.search_encoded("iupac", pats, t, k, rc=False, all_minima=allm))
            got, _ = cpu_port.search_batch(pats, t, len(t), k, all_minima=allm, prefilter=prefilter,
                                           threads=rng.choice([1, 3]))
            assert sorted(got) == want, (pats, t, k, allm)


def rand_text(rng, n):
    return bytes(rng.choice(b"ACGT") for _ in range(n))
