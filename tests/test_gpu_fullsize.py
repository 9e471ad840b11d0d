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
    import torch
    import bench
    import sassy_b200
    if torch.cuda.get_device_properties(0).total_memory < 40e9:
        pytest.skip("needs a GPU with room for the 3 GB text and its shards")
    dev = torch.device("cuda", 0)
    m, k = 20, 2
    pats = bench.make_patterns("dna", 1, m)
    text = bench.synth_text_device(torch, N, 42, dev)
    plants = bench.plant_list(pats, N, k, 64, seed=44)
    # also plant reverse complements and copies right at the ends and at the cut
    rng = random.Random(7)
    cut = 1_400_000_123
    extra = [(0, pats[0]), (N - m, pats[0]), (cut - 7, pats[0]), (cut + 1, oracle.reverse_complement("dna", pats[0]))]
    for i in range(16):
        extra.append((rng.randrange(1000, N - 1000) // 64 * 64 + 61, oracle.reverse_complement("dna", pats[0])))
    for pos, q in plants + extra:
        text[pos:pos + len(q)] = torch.tensor(list(q), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    s = sassy_b200.Searcher("dna", rc=True, device=0)
    return dict(torch=torch, s=s, text=text, pat=pats[0], m=m, k=k, plants=plants + extra, cut=cut)


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
