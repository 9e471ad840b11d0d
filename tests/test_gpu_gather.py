"""The fused peer-memory gather (csrc/peer_gather.cu) behind the search: world size 1 on any
GPU box, world size 2 (two processes, CUDA IPC + NVLink stores) when the box has >= 2 GPUs."""
import os
import random
import socket

import pytest

import oracle
from tests.test_oracle_props import planted, rand_seq

pytestmark = pytest.mark.gpu


def key(m):
    return (m.pattern_idx, m.text_idx, m.text_start, m.text_end, m.cost, m.strand, m.cigar)


def _text_with_plants(rng, n, pats, k):
    t = bytearray(rand_seq(rng, n))
    for i, p in enumerate(pats):
        for j in range(3):
            a = rng.randrange(0, n - len(p))
            q = bytearray(p.replace(b"N", b"A"))
            if j:
                q[rng.randrange(len(q) - 3)] = ord("A")
            t[a:a + len(q)] = q
    return bytes(t)


def test_gather_world1_matches_plain_search():
    import sassy_b200
    from sassy_b200 import dist as sd
    rng = random.Random(41)
    s = sassy_b200.Searcher("dna", rc=True)
    p = rand_seq(rng, 20)
    t = _text_with_plants(rng, 300_000, [p], 2)
    dt = s.upload_text(t)
    pg = sd.PeerGather(s, max_ops=20 + 2 + 1)
    for _ in range(3):  # both parities of the receive buffer
        got = pg.search(p, dt, 2)
        want = s.search(p, dt, 2)
        assert list(map(key, got)) == list(map(key, want)) and len(got) >= 3
    assert list(map(key, want)) == [(x.pattern_idx, 0, x.text_start, x.text_end, x.cost, x.strand, x.cigar)
                                    for x in oracle.search("dna", p, t, 2, rc=True)]
    # a result that does not fit the exchange (every position matches): the local list comes back
    dense = s.upload_text(b"A" * 50_000)
    got = pg.search(b"AAAA", dense, 0, all_minima=True)
    assert len(got) == 50_000 - 3  # forward strand only: complement(AAAA) never occurs
    # encoded patterns
    si = sassy_b200.Searcher("iupac", rc=True)
    pats = [rand_seq(rng, 20) + b"NGG" for _ in range(6)]
    t = _text_with_plants(rng, 200_000, pats, 3)
    dti = si.upload_text(t)
    enc = si.encode_patterns(pats)
    pgi = sd.PeerGather(si, max_ops=23 + 3 + 1)
    got = pgi.search_encoded(enc, dti, 3)
    want = si.search_encoded_patterns(enc, dti, 3)
    assert sorted(map(key, got)) == sorted(map(key, want)) and len(got) >= 6
    pg.close()
    pgi.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import sassy_b200
    from sassy_b200 import dist as sd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    rng = random.Random(100)          # same pattern on every rank
    p = rand_seq(rng, 24)
    trng = random.Random(200 + rank)  # own text shard
    t = _text_with_plants(trng, 400_000 + 1000 * rank, [p], 3)
    s = sassy_b200.Searcher("dna", rc=True, device=rank)
    dt = s.upload_text(t)
    pg = sd.PeerGather(s, max_ops=24 + 3 + 1)
    out = []
    for step in range(4):
        out.append(sorted(map(key, pg.search(p, dt, 3))))
    # step 5: rank 1 produces more matches than fit -> every rank must take the fall-back
    dense = s.upload_text(b"A" * (30_000 if rank == 1 else 3_000))
    fb = pg.search(b"AAAAAA", dense, 0, all_minima=True)
    after = sorted(map(key, pg.search(p, dt, 3)))
    local = sorted((x.pattern_idx, rank, x.text_start, x.text_end, x.cost, x.strand, x.cigar)
                   for x in oracle.search("dna", p, t, 3, rc=True))
    q.put((rank, out, local, len(fb), pg.fallbacks, after))
    dist.barrier()
    pg.close()
    dist.destroy_process_group()


def test_gather_world2_peer_memory():
    import sassy_b200
    if sassy_b200.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    want = sorted(res[0][2] + res[1][2])
    assert len(want) >= 6
    for rank, out, local, nfb, fallbacks, after in res:
        for step_result in out:
            assert step_result == want
        assert after == want
        assert nfb == (30_000 - 5) + (3_000 - 5)
        assert fallbacks == 1


def test_sharded_world1_pipelined_equals_plain_search():
    """One slab = the whole text: the pipelined exchange returns the result of the previous call,
    flush the last one; all equal the plain search."""
    import sassy_b200
    from sassy_b200 import dist as sd
    rng = random.Random(43)
    s = sassy_b200.Searcher("dna", rc=True)
    p = rand_seq(rng, 20)
    t = _text_with_plants(rng, 400_000, [p], 2)
    dt = s.upload_text(t)
    want = list(map(key, s.search(p, dt, 2)))
    layout = sd.slab_layout(len(t), 1, 20, 2)
    for pipelined in (False, True):
        pg = sd.PeerGather(s, max_ops=20 + 2 + 1, pipelined=pipelined)
        got = [pg.search_sharded(p, dt, 2, layout, len(t)) for _ in range(4)]
        if pipelined:
            assert got[0] is None
            got = got[1:] + [pg.flush_sharded(20, layout, len(t))]
            assert pg.flush_sharded(20, layout, len(t)) is not None  # idempotent: the same last result
        for g in got:
            assert list(map(key, g)) == want and len(want) >= 3
        pg.close()


def _shard_worker(rank, world, port, q, pipelined):
    import torch
    import torch.distributed as dist
    import sassy_b200
    from sassy_b200 import dist as sd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["SASSY_B200_GATHER_TIMEOUT_S"] = "20"
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    rng = random.Random(300)  # the same global text on every rank
    m, k = 20, 2
    p = rand_seq(rng, m)
    n = 600_001
    t = bytearray(_text_with_plants(rng, n, [p], k))
    cut = -(-n // world)
    rcp = oracle.reverse_complement("dna", p)
    t[cut - 7:cut - 7 + m] = p            # copies across the slab border, both strands
    t[cut + 40 - 3:cut + 40 - 3 + m] = rcp
    t[cut + 200 - m:cut + 200] = p[:10] + bytes([p[10]]) * 1 + p[10:19]  # a plateau next to the border
    t = bytes(t)
    layout = sd.slab_layout(n, world, m, k)
    wlo, whi, lo, hi = layout[rank]
    s = sassy_b200.Searcher("dna", rc=True, device=rank)
    dt = s.upload_text(t[wlo:whi])
    pg = sd.PeerGather(s, max_ops=m + k + 1, pipelined=pipelined)
    outs = []
    for step in range(4):
        r = sd.search_text_sharded(s, p, dt, k, n, peer_gather=pg)
        if r is not None:
            outs.append(list(map(key, r)))
    if pipelined:
        outs.append(list(map(key, pg.flush_sharded(m, layout, n))))
    nccl = list(map(key, sd.search_text_sharded(s, p, dt, k, n)))  # all-gather route
    want = [(x.pattern_idx, 0, x.text_start, x.text_end, x.cost, x.strand, x.cigar)
            for x in oracle.search("dna", p, t, k, rc=True)]
    q.put((rank, outs, nccl, want))
    dist.barrier()
    pg.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("pipelined", [False, True], ids=["lockstep", "pipelined"])
def test_text_sharded_world2_equals_unsharded(pipelined):
    """ONE text cut into 2 slabs on 2 GPUs (halo, fused gather, merged local-minima rule) == the
    oracle's search of the whole text, with matches and a plateau across the cut."""
    import sassy_b200
    if sassy_b200.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q, pipelined)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, outs, nccl, want in res:
        assert len(want) >= 5 and len(outs) == 4
        for o in outs:
            assert o == want
        assert nccl == want
