"""Times the host-pointer search of the c2 workload (3 GB pinned Dna text, m = 20, k = 2) under the
transport tuning knobs of csrc/engine.cu: one fresh process per setting (the knobs are read once).

    python tools/e2e_probe.py                      # sweep
    python tools/e2e_probe.py --one                # this process, current environment
"""
import argparse
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def one(n, steps):
    import numpy as np
    import torch
    import bench
    import sassy_b200
    dev = torch.device("cuda:0")
    args = argparse.Namespace(c5_patterns=2048)
    text = bench.build_window(torch, args, n, 1, 0, n, dev)
    host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    host.copy_(text)
    torch.cuda.synchronize()
    del text
    pats = bench.workload_patterns("c2", 1)
    s = sassy_b200.Searcher("dna", rc=False)
    buf = (host.data_ptr(), n)
    for _ in range(2):
        ms = s.search(pats[0], buf, 2)
    walls, xfer = [], []
    for _ in range(steps):
        t0 = time.perf_counter()
        ms = s.search(pats[0], buf, 2)
        walls.append(time.perf_counter() - t0)
        xfer.append(s.stats()["transfer_ms"])
    st = s.stats()
    walls.sort()
    print(json.dumps({"gbps_median": n / walls[len(walls) // 2] / 1e9, "gbps_best": n / walls[0] / 1e9,
                      "wall_ms": [round(w * 1e3, 2) for w in walls], "transfer_ms": round(sorted(xfer)[len(xfer) // 2], 2),
                      "h2d_bytes": st["transfer_bytes"], "matches": len(ms)}))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--one", action="store_true")
    ap.add_argument("--n", type=int, default=3_000_000_000)
    ap.add_argument("--steps", type=int, default=7)
    a = ap.parse_args()
    if a.one:
        one(a.n, a.steps)
        sys.exit(0)
    settings = [
        {},
        {"SASSY_B200_PACK_MERGE": "1"},
        {"SASSY_B200_PACK_MERGE": "4"},
        {"SASSY_B200_PACK_MERGE": "16"},
        {"SASSY_B200_PACK_THREADS": "15"},
        {"SASSY_B200_PACK_THREADS": "14"},
        {"SASSY_B200_PACK_THREADS": "12"},
        {"SASSY_B200_PACK_RING": "32"},
        {"SASSY_B200_PACK_RING": "64"},
        {"SASSY_B200_PACK_CHUNK": str(1 << 20), "SASSY_B200_PACK_RING": "96", "SASSY_B200_PACK_MERGE": "16"},
        {"SASSY_B200_PACK_CHUNK": str(1 << 19), "SASSY_B200_PACK_RING": "192", "SASSY_B200_PACK_MERGE": "32"},
        {"SASSY_B200_PACK_AHEAD": "8192"},
        {"SASSY_B200_PACK_THREADS": "15", "SASSY_B200_PACK_AHEAD": "8192"},
        {"SASSY_B200_PACK_RING": "0", "SASSY_B200_PACK_CHUNK": str(8 << 20)},
    ]
    for env in settings:
        e = dict(os.environ)
        e.update(env)
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", "--n", str(a.n), "--steps", str(a.steps)],
                             env=e, capture_output=True, text=True)
        line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:]
        print(json.dumps(env), line, flush=True)
