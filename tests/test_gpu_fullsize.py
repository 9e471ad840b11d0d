"""Size-independent properties at BASELINE.json's full text size (3 GB), where the O(m n) oracle
cannot follow: (1) every planted copy is found with at most its number of edits, and nothing the
oracle would reject is reported -- every reported match is re-checked by the oracle on its own
window; (2) sharding: the matches of the whole text equal the union of the matches of two
overlapping halves (every end position is decided by the m+k characters before it),
the property the multi-GPU text sharding relies on; (3) resident, host-pointer (packed
transport) and sharded searches agree."""
import random

import pytest

import oracle

pytestmark = pytest.mark.gpu

N = 3_000_000_000


def key(m):
    return (m.text_start, m.text_end, m.cost, m.strand, m.cigar)


@pytest.fixture(scope="module")
def world():
    import argparse
    import torch
    import bench
    import sassy_b200
    if torch.cuda.get_device_properties(0).total_memory < 40e9:
        pytest.skip("needs a GPU with room for the 3 GB text and its shards")
    dev = torch.device("cuda", 0)
    m, k = 20, 2
    args = argparse.Namespace(c5_patterns=2048)
    # bench.py's text: uniform ACGT with the c2 / c4 patterns planted 64 times and the c3 / c5
    # guides twice (0..k edits each)
    text = bench.build_window(torch, args, N, 1, 0, N, dev)
    pat = bench.workload_patterns("c2", 1)[0]
    plants = [(pos, q) for pos, q in bench.slab_plants(0, N, bench.planted_set(args)[:1])]
    # also plant reverse complements and copies right at the ends and at the cut
    rng = random.Random(7)
    cut = 1_400_000_123
    extra = [(0, pat), (N - m, pat), (cut - 7, pat), (cut + 1, oracle.reverse_complement("dna", pat))]
    for i in range(16):
        extra.append((rng.randrange(1000, N - 1000) // 64 * 64 + 61, oracle.reverse_complement("dna", pat)))
    for pos, q in extra:
        text[pos:pos + len(q)] = torch.tensor(list(q), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    s = sassy_b200.Searcher("dna", rc=True, device=0)
    return dict(torch=torch, s=s, text=text, pat=pat, m=m, k=k, plants=plants + extra, cut=cut, args=args)


def test_fullsize_planted_and_sharding(world):
    torch, s, text, p, m, k, cut = (world[x] for x in ("torch", "s", "text", "pat", "m", "k", "cut"))
    dt = s.text_from_device(text.data_ptr(), N)
    full = s.search_all(p, dt, k)
    assert s.stats()["filter_words"] > 0  # the production route (prefilter + re-scan)
    got = sorted(map(key, full))
    # (1a) every planted copy shows up: some match overlaps it with cost <= k
    starts = sorted((x.text_start, x.text_end) for x in full)
    import bisect
    for pos, q in world["plants"]:
        i = bisect.bisect_left(starts, (pos - m - k, 0))
        assert any(a < pos + len(q) and b > pos for a, b in starts[i:i + 200]), pos
    # (1b) every reported match is what the oracle reports on the match's own neighbourhood
    sample = full[:200] + full[-200:]
    for x in sample:
        lo = max(0, x.text_start - 2 * (m + k))
        hi = min(N, x.text_end + 2 * (m + k))
        window = bytes(text[lo:hi].cpu().numpy().tobytes())
        want = [(w.text_start + lo, w.text_end + lo, w.cost, w.strand, w.cigar)
                for w in oracle.search("dna", p, window, k, rc=True, all_minima=True)]
        if lo > 0:  # windows cut out of the text: ignore ends inside the first m+k characters
            want = [w for w in want if (w[1] if w[3] == "+" else hi - (w[0] - lo)) > 0]
        assert key(x) in want, (key(x), want[:5])
    # (2) sharding property on end positions
    halo = m + k
    left = s.text_from_device(text.data_ptr(), cut + halo)           # rc ends near the cut need the halo
    right = s.text_from_device(text.data_ptr() + cut - halo, N - cut + halo)
    lm = [key(x) for x in s.search_all(p, left, k)]
    rm = [(a + cut - halo, b + cut - halo, c, d, e) for a, b, c, d, e in map(key, s.search_all(p, right, k))]
    # forward matches are owned by the shard holding their END, rc matches (scanned right to left) by
    # the shard holding their START
    def owner_left(t):
        return (t[1] <= cut) if t[3] == "+" else (t[0] < cut)
    union = sorted([t for t in lm if owner_left(t)] + [t for t in rm if not owner_left(t)])
    assert union == got
    assert len(got) >= 64
    # local-minima mode on the whole text is a subset of the all-positions mode
    loc = sorted(map(key, s.search(p, dt, k)))
    assert set(loc) <= set(got) and len(loc) >= 64
    world["full_local"] = loc
    for t in (left, right, dt):
        t.free()


def test_fullsize_host_pointer_equals_resident(world):
    torch, s, text, p, k = (world[x] for x in ("torch", "s", "text", "pat", "k"))
    host = torch.empty(N, dtype=torch.uint8, pin_memory=True)
    host.copy_(text)
    torch.cuda.synchronize()
    got = sorted(map(key, s.search(p, (host.data_ptr(), N), k)))
    assert s.stats()["transfer_packed"] == 1  # crossed PCIe at 2 bits per character
    dt = s.text_from_device(text.data_ptr(), N)
    want = sorted(map(key, s.search(p, dt, k)))
    assert got == want and len(got) >= 64


# ---- the other BASELINE shapes at full size --------------------------------------------------
# c4 (Dna, m = 100, k = 8, multi-word), c3 (Iupac batch, m = 23, k = 4) and c5 (k = 3): the GPU end
# positions and costs over the whole 3 GB equal the CPU port's (itself pinned to the oracle by
# tests/test_cpu_port.py), every planted copy is found, and a sample of the reported matches is
# re-derived by the oracle (coordinates, cost and CIGAR) on the match's own neighbourhood.

def _host_copy(world):
    torch, text = world["torch"], world["text"]
    if "host" not in world:
        host = torch.empty(N, dtype=torch.uint8, pin_memory=True)
        host.copy_(text)
        torch.cuda.synchronize()
        world["host"] = host
    return world["host"]


def _oracle_recheck(profile, pats, text, ms, m, k, encoded, limit=300):
    step = max(1, len(ms) // limit)
    for x in list(ms)[::step][:limit]:
        lo = max(0, x.text_start - 2 * (m + k))
        hi = min(N, x.text_end + 2 * (m + k))
        window = bytes(text[lo:hi].cpu().numpy().tobytes())
        p = pats[x.pattern_idx]
        if encoded:
            want = oracle.search_encoded(profile, [p], window, k, rc=False, all_minima=True)
        else:
            want = oracle.search(profile, p, window, k, rc=False, all_minima=True)
        want = [(w.text_start + lo, w.text_end + lo, w.cost, w.cigar) for w in want]
        assert (x.text_start, x.text_end, x.cost, x.cigar) in want, (x, want[:4])


def test_fullsize_c4_shape(world):
    import bench
    import sassy_b200
    from oracle import cpu_port
    text = world["text"]
    host = _host_copy(world)
    p = bench.workload_patterns("c4", 1)[0]
    m, k = 100, 8
    s = sassy_b200.Searcher("dna", rc=False, device=0)
    dt = s.text_from_device(text.data_ptr(), N)
    got = s.search(p, dt, k)
    assert s.stats()["filter_words"] > 0 and not s.stats()["filter_fallback"]  # the production route
    ends, _ = cpu_port.search_ends("dna", p, host.data_ptr(), N, k, False, False, threads=16)
    assert sorted((x.text_end, x.cost) for x in got) == sorted((e, c) for e, c, _ in ends)
    assert len(got) >= 64
    _oracle_recheck("dna", [p], text, got, m, k, encoded=False)
    # search_all: superset of search, and every position the CPU port reports
    got_all = s.search_all(p, dt, k)
    ends_all, _ = cpu_port.search_ends("dna", p, host.data_ptr(), N, k, False, True, threads=16, cap=1 << 22)
    assert sorted((x.text_end, x.cost) for x in got_all) == sorted((e, c) for e, c, _ in ends_all)
    dt.free()


@pytest.mark.parametrize("k,npat", [(4, 12), (3, 24)])
def test_fullsize_guide_batches(world, k, npat):
    """c3 (k = 4) and c5 (k = 3) shapes: encoded Iupac guides (20 nt + NGG) over the whole text."""
    import bench
    import sassy_b200
    from oracle import cpu_port
    text = world["text"]
    host = _host_copy(world)
    pats = bench.workload_patterns("c3", npat)
    m = 23
    s = sassy_b200.Searcher("iupac", rc=False, device=0)
    dt = s.text_from_device(text.data_ptr(), N)
    got = s.search_encoded_patterns(s.encode_patterns(pats), dt, k)
    want = []
    for pi, p in enumerate(pats):
        ends, _ = cpu_port.search_ends("iupac", p, host.data_ptr(), N, k, False, False, threads=16)
        want += [(pi, e, c) for e, c, _ in ends]
    assert sorted((x.pattern_idx, x.text_end, x.cost) for x in got) == sorted(want)
    assert len(got) >= 2 * npat  # every guide was planted twice
    _oracle_recheck("iupac", pats, text, got, m, k, encoded=True)
    dt.free()
