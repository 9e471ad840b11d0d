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
