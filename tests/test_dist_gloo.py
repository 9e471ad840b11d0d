"""World-size-2 check (gloo, CPU) of the multi-GPU plumbing: round-robin sharding and the
final all-gather of match records give the same multiset as the unsharded search."""
import os
import random
import socket

import torch.multiprocessing as mp

import oracle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, patterns, text, k, q):
    import torch.distributed as dist
    from sassy_b200 import dist as sd
    from sassy_b200.searcher import Match
    from tests.emu_backend import EmuBackend
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sd.shard_indices(len(patterns), rank, world)
    local = []
    if mine:
        res = EmuBackend().search_encoded("iupac", [patterns[i] for i in mine], text, k, rc=True)
        for m in res:
            ops = "".join(ch * int(cnt) for cnt, ch in __import__("re").findall(r"(\d+)([=XID])", m.cigar))
            local.append(Match(mine[m.pattern_idx], 0, m.text_start, m.text_end, m.pattern_start, m.pattern_end,
                               m.cost, m.strand, ops))
    allm = sd.gather_matches(local, max_ops=len(patterns[0]) + k + 1)
    if rank == 0:
        q.put(sorted((m.pattern_idx, m.text_start, m.text_end, m.cost, m.strand, m.cigar) for m in allm))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_search_equals_unsharded():
    rng = random.Random(7)
    m, n, k = 23, 4000, 3
    text = bytearray(rng.choice(b"ACGT") for _ in range(n))
    patterns = []
    for i in range(5):
        p = bytes(rng.choice(b"ACGT") for _ in range(20)) + b"NGG"
        patterns.append(p)
        pos = rng.randrange(0, n - m)
        text[pos:pos + m] = p[:20] + b"AGG"
    text = bytes(text)
    want = sorted((x.pattern_idx, x.text_start, x.text_end, x.cost, x.strand, x.cigar)
                  for x in oracle.search_encoded("iupac", patterns, text, k, rc=True))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, patterns, text, k, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == want and len(want) >= 5


def test_pack_roundtrip():
    from sassy_b200 import dist as sd
    from sassy_b200.searcher import Match
    ms = [Match(3, 1, 10, 33, 0, 23, 2, "-", "=" * 10 + "X" + "D" + "=" * 11 + "I"),
          Match(0, 0, 2**35, 2**35 + 70, 0, 64, 0, "+", "=" * 64 + "DDDDDD")]
    t = sd.pack_matches(ms, 3)
    assert sd.unpack_matches(t) == ms


def _worker_ml(rank, world, port, q):
    """gather_matches on MatchList inputs (the form the GPU path returns) incl. a growing block."""
    import numpy as np
    import torch.distributed as dist
    from sassy_b200 import dist as sd
    from sassy_b200.searcher import MatchList, _REC_DTYPE
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = []
    for n in (3, 0 if rank == 0 else 700, 5):  # 700 > the initial capacity of 256: the block grows
        recs = np.zeros(n, dtype=_REC_DTYPE)
        ops = []
        off = 0
        for i in range(n):
            o = "=" * (10 + (i + rank) % 7) + "X"
            recs[i] = (i, 0, 100 * i + rank, 100 * i + rank + len(o), 0, len(o), 1, rank, (0, 0, 0), len(o), 0, off)
            ops.append(o)
            off += len(o)
        ml = sd.tag_rank(MatchList(recs, "".join(ops).encode()), rank)
        g = sd.gather_matches(ml, max_ops=24)
        out.append(sorted((m.pattern_idx, m.text_idx, m.text_start, m.text_end, m.cost, m.strand, m.cigar) for m in g))
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_gather_matchlists_and_rank_tags():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_ml, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for got, sizes in zip(out, ((3, 3), (0, 700), (5, 5))):
        assert len(got) == sum(sizes)
        for rank, n in enumerate(sizes):
            mine = [g for g in got if g[1] == rank]  # text_idx = source rank
            assert len(mine) == n
            for i, g in enumerate(sorted(mine)):
                o = 10 + (i + rank) % 7
                assert g == (i, rank, 100 * i + rank, 100 * i + rank + o + 1, 1, "-" if rank else "+", f"{o}=1X")
