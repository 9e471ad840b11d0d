"""search_many on the nanopore-barcode shape (bench.py run_nanopore) with host phase timers:
SASSY_B200_HOST_TIMING=1 python tools/nanopore_probe.py"""
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import sassy_b200

rng = random.Random(47)
n = 400_000_000
host = np.frombuffer(np.random.default_rng(1).integers(0, 4, n, dtype=np.uint8).tobytes().translate(
    bytes(b"ACGT"[c & 3] for c in range(256))), dtype=np.uint8)
barcodes = [bytes(rng.choice(b"ACGT") for _ in range(24)) for _ in range(96)]
total, reads = 0, []
while total < 334_294_335:
    ln = rng.randrange(2000, 12000)
    a0 = rng.randrange(0, n - ln)
    reads.append(host[a0:a0 + ln].tobytes())
    total += ln
s = sassy_b200.Searcher("iupac", rc=True)
for _ in range(2):
    ms = s.search_many(barcodes, reads, 3)
ts = []
for _ in range(int(os.environ.get("PROBE_CALLS", "30"))):
    t0 = time.perf_counter()
    ms = s.search_many(barcodes, reads, 3)
    ts.append(time.perf_counter() - t0)
print(len(reads), "reads", total, "bp", len(ms), "matches", "ms per call", sorted(round(t * 1e3, 1) for t in ts)[::6])
