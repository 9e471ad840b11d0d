"""Prints the engine statistics (route, hits, confirmed entries, kernel times) of one Dna search on the
bench text: python tools/route_stats.py M K [--rc]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import sassy_b200

ap = argparse.ArgumentParser()
ap.add_argument("m", type=int)
ap.add_argument("k", type=int)
ap.add_argument("--rc", action="store_true")
ap.add_argument("--n", type=int, default=3_000_000_000)
a = ap.parse_args()
dev = torch.device("cuda:0")
args = argparse.Namespace(c5_patterns=2048)
t = bench.build_window(torch, args, a.n, 1, 0, a.n, dev)
s = sassy_b200.Searcher("dna", rc=a.rc)
dt = s.text_from_device(t.data_ptr(), a.n)
p = bench.make_patterns("dna", 1, a.m, seed=bench.PATTERN_SEED["dna"] + a.m)[0]
if a.m not in (20, 100):  # plant a few copies with edits
    import random
    rng = random.Random(5)
    host = bytearray(p)
    for j in range(16):
        q = bytearray(p)
        for _ in range(rng.randrange(0, a.k + 1)):
            q[rng.randrange(a.m)] = rng.choice(b"ACGT")
        pos = (j + 1) * (a.n // 20)
        t[pos:pos + a.m] = torch.tensor(list(q), dtype=torch.uint8, device=dev)
for _ in range(3):
    ms = s.search(p, dt, a.k)
print(len(ms), "matches", s.stats())
