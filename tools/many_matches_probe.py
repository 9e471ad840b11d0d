import sys, time, argparse
sys.path.insert(0, "/root/repo")
import torch, bench, sassy_b200
dev = torch.device("cuda:0")
args = argparse.Namespace(c5_patterns=2048)
n = 3_000_000_000
t = bench.build_window(torch, args, n, 1, 0, n, dev)
s = sassy_b200.Searcher("dna", rc=False)
dt = s.text_from_device(t.data_ptr(), n)
p = bench.workload_patterns("c2", 1)[0]
for k in (4, 5, 6):
    for _ in range(2):
        ms = s.search(p, dt, k)
    t0 = time.perf_counter(); ms = s.search(p, dt, k); wall = time.perf_counter() - t0
    st = s.stats()
    print(k, len(ms), "wall ms", round(wall * 1e3, 2), "device total", round(st["total_ms"], 2), "scan", round(st["scan_ms"], 2), "cands", st["candidates"])
