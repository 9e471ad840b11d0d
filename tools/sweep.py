"""Pattern-length x k sweep on one GPU (north star: query len 20/100/1000, k = 1..8, 3 GB synthetic DNA).

    python tools/sweep.py [--text-bytes N] [--steps K] > profiles/<name>.md

The text is generated once in HBM (bench.py's generator, seed 42) with 64 planted copies per
pattern; every (m, k) cell reports whole-search throughput (text resident, local-minima mode with
traceback), matches per second, the route the engine took (exact piece prefilter + re-scan, or
full bit-parallel scan), and the dominant kernel's algorithmic bytes / time as a fraction of the
measured HBM peak (MEASURED_PEAKS.json)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--text-bytes", type=int, default=3_000_000_000)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--lens", default="20,100,1000")
    ap.add_argument("--ks", default="1,2,3,4,5,6,7,8")
    ap.add_argument("--rc", action="store_true")
    args = ap.parse_args()
    import torch
    import bench
    import sassy_b200
    dev = torch.device("cuda", 0)
    n = args.text_bytes
    lens = [int(x) for x in args.lens.split(",")]
    ks = [int(x) for x in args.ks.split(",")]
    text = bench.synth_text(torch, n, 0, 42, dev)
    pats = {m: bench.make_patterns("dna", 1, m, seed=43 + m)[0] for m in lens}
    # plant for the largest k: 64 copies per pattern with 0..8 edits (slots of 4 KB: m <= 1000)
    bench.PLANT_SLOT = 4096
    bench.apply_plants(torch, text, 0, bench.slab_plants(0, n, [(pats[m], max(ks), 64) for m in lens]))
    torch.cuda.synchronize()
    s = sassy_b200.Searcher("dna", rc=args.rc, device=0)
    dt = s.text_from_device(text.data_ptr(), n)
    del text
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    print(f"# Sweep: Dna, 1 pattern, {n / 1e9:.1f} GB synthetic ACGT resident in HBM, {'both strands' if args.rc else 'forward strand'}, "
          f"{args.steps} steps per cell, 1 x B200 (HBM peak used: {peak:.0f} GB/s)\n")
    print("| m | k | GB/s text | ms/search | matches | matches/s | route | dominant kernel ms | of HBM peak |")
    print("|---|---|---|---|---|---|---|---|---|")
    for m in lens:
        dense = False
        for k in ks:
            if k >= m:
                continue
            if dense:  # random text matches by the million: the cell measures result handling, not the search
                print(f"| {m} | {k} | - | - | > 2 M chance matches at k - 1 | - | skipped | - | - |", flush=True)
                continue
            t0 = time.perf_counter()
            ms = s.search(pats[m], dt, k)
            once = time.perf_counter() - t0
            steps = max(2, min(args.steps, int(1.0 / max(once, 1e-4))))
            for _ in range(2 if once < 0.1 else 0):
                ms = s.search(pats[m], dt, k)
            dense = len(ms) > 2_000_000
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            kern = []
            for _ in range(steps):
                ms = s.search(pats[m], dt, k)
                kern.append(s.stats()["scan_ms"])
            el = (time.perf_counter() - t0) / steps
            st = s.stats()
            if st["filter_words"] and not st["filter_fallback"]:
                route = (f"q-gram bitmap (q = {st['filter_len']}), {st['hits']} hits confirmed + re-scanned" if st["filter_kind"] == 2
                         else f"piece automaton ({st['filter_words']} words, pieces >= {st['filter_len']}) + re-scan of {st['hits']} hits")
            else:
                route = f"full scan ({st['words']} words)"
            kms = sum(kern) / len(kern)
            print(f"| {m} | {k} | {n / el / 1e9:.0f} | {el * 1e3:.3f} | {len(ms)} | {len(ms) / el:.0f} | {route} | "
                  f"{kms:.3f} | {n / (kms * 1e-3) / 1e9 / peak * 100:.1f} % |", flush=True)


if __name__ == "__main__":
    main()
