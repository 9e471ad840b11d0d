#!/usr/bin/env python
"""Benchmark of the Searcher::search hot path on B200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|c3|c4]

One "step" = one search of the pattern (batch) over the rank's whole text.  Default
workload = BASELINE.json configs[1]: Dna profile, one pattern of length 20, k=2, 3 GB of
synthetic ACGT per GPU (weak scaling: every rank scans its own 3 GB shard; the match
records are all-gathered over NCCL at the end of every step).

Prints ONE JSON line (rank 0).  `value` = text bytes scanned per second with the text
resident in HBM; `e2e` = the same search through the host-pointer C-ABI entry point
(host->device copy of the text and device->host copy of the matches inside the timed
region).  `--impl reference` times the CPU restatement of the reference's algorithm
(oracle/, the reference itself is Rust and cannot be built here) on the host cores.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import random
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (profile, n_patterns, m, k, text_bytes, description)
    "c1": ("dna", 1, 20, 1, 1 << 20, "Dna, 1 pattern len=20, k=1, 1 MB random ACGT (BASELINE configs[0])"),
    "c2": ("dna", 1, 20, 2, 3_000_000_000, "Dna, 1 pattern len=20, k=2, 3 GB synthetic ACGT (BASELINE configs[1])"),
    "c3": ("iupac", 1024, 23, 4, 3_000_000_000, "Iupac, 1024 encoded patterns len=23 (20nt+NGG), k=4, 3 GB (BASELINE configs[2])"),
    "c4": ("dna", 1, 100, 8, 3_000_000_000, "Dna, 1 pattern len=100, k=8, 3 GB synthetic ACGT (BASELINE configs[3])"),
    # pattern shards: every rank holds the whole text and searches patterns rank, rank + N, ...;
    # 100 000 x 3 GB is 3e14 character-pattern steps (minutes per step): use --patterns to bound a run
    "c5": ("iupac", 100_000, 23, 3, 3_000_000_000, "Iupac, 100k encoded patterns len=23 (20nt+NGG), k=3, 3 GB, patterns sharded over the GPUs (BASELINE configs[4])"),
}
PATTERN_SHARDED = {"c5"}
METRIC = "GB/s text scanned"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--text-bytes", type=int, default=0, help="override the text size (debug)")
    ap.add_argument("--patterns", type=int, default=0, help="override the number of patterns (debug)")
    ap.add_argument("--variant", default="tma", choices=["tma", "ldg"])
    ap.add_argument("--filter", default="auto", choices=["off", "auto", "force"],
                    help="exact piece prefilter in front of the scan (result-neutral)")
    ap.add_argument("--rc", action="store_true", help="search both strands (default: forward only, as the reference's evals)")
    ap.add_argument("--transport", default="packed", choices=["packed", "bytes"],
                    help="host->device transport of Dna texts in the e2e leg (2 bits per character, or bytes)")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="N > 1: fused peer-memory exchange of the match records (default) or NCCL all-gather")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------
# synthetic workload (same on every leg): uniform ACGT text, uniform ACGT patterns, planted
# copies with 0..k edits (seeded), cf. BASELINE.md section 2.

def make_patterns(profile, n_patterns, m, seed=43):
    rng = random.Random(seed)
    pats = []
    for _ in range(n_patterns):
        if profile == "iupac":
            pats.append(bytes(rng.choice(b"ACGT") for _ in range(m - 3)) + b"NGG")
        else:
            pats.append(bytes(rng.choice(b"ACGT") for _ in range(m)))
    return pats


def mutate(rng, p, edits):
    p = bytearray(p)
    for _ in range(edits):
        op = rng.randrange(3)
        pos = rng.randrange(len(p))
        if op == 0:
            p[pos] = rng.choice(b"ACGT")
        elif op == 1:
            p.insert(pos, rng.choice(b"ACGT"))
        elif len(p) > 1:
            del p[pos]
    return bytes(p)


def plant_list(pats, n, k, copies, seed=44):
    """[(pos, bytes)] non-overlapping planted copies (concrete ACGT, N -> A)."""
    rng = random.Random(seed)
    out = []
    m = len(pats[0])
    slots = max(1, n // (4 * (m + k + 1)))
    used = set()
    for p in pats:
        for _ in range(copies):
            s = rng.randrange(slots)
            while s in used:
                s = rng.randrange(slots)
            used.add(s)
            q = mutate(rng, p.replace(b"N", b"A"), rng.randrange(0, k + 1))
            pos = s * 4 * (m + k + 1)
            if pos + len(q) <= n:
                out.append((pos, q))
    return out


def synth_text_device(torch, n, seed, device):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty(n, dtype=torch.uint8, device=device)
    ch = 1 << 28
    for off in range(0, n, ch):
        mlen = min(ch, n - off)
        r = torch.randint(0, 4, (mlen,), dtype=torch.uint8, device=device, generator=g)
        # 0,1,2,3 -> 'A','C','G','T' (65,67,71,84)
        o = r * 2 + 65
        o += (r == 2).to(torch.uint8) * 2
        o += (r == 3).to(torch.uint8) * 13
        out[off:off + mlen] = o
        del r, o
    return out


class ClockSampler:
    """SM clock / throttle-reason samples taken DURING the timed region through NVML
    (nvidia-ml-py), every `every` steps from inside the step loop; falls back to one
    `nvidia-smi` query if NVML is unavailable."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, uuid: str, every: int = 8):
        self.every = every
        self.sm = []
        self.power = []
        self.reasons = set()
        self.sm_max = None
        self.h = None
        self.uuid = uuid
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:
            self.h = None
            self.err = str(e)

    def sample(self, step: int = 0):
        # NVML queries cost up to a millisecond: at most one sample per 50 ms of wall time
        now = time.perf_counter()
        if self.h is None or now - getattr(self, "_last", 0.0) < 0.05:
            return
        self._last = now
        nv = self.nv
        try:
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
            self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            try:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in self.REASONS.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception as e:
            self.err = str(e)

    def result(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "source": "nvml"}
        if self.sm:
            out["sm_mhz"] = statistics.median(self.sm)
            out["samples"] = len(self.sm)
            out["power_w_max"] = max(self.power) if self.power else None
            return out
        try:  # fallback: one nvidia-smi query (not under load)
            q = "clocks.sm,clocks.max.sm,power.draw"
            txt = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", self.uuid],
                                 capture_output=True, text=True, timeout=20).stdout.strip().split(", ")
            out.update({"sm_mhz": float(txt[0]), "sm_max_mhz": float(txt[1]), "power_w_max": float(txt[2]),
                        "samples": 1, "source": "nvidia-smi after the timed region (NVML unavailable: %s)" % getattr(self, "err", "?")})
        except Exception as e:
            out["error"] = str(e)
        return out


# ----------------------------------------------------------------------------------------
# CPU leg (oracle/ is test + baseline infrastructure; never on the product path)

def cpu_port_run(profile, pats, k, rc, text_addr, nbytes, threads):
    """Runs the CPU restatement of the reference's search over text[0:nbytes].
    Returns (seconds, matches, kind_description)."""
    from oracle import cpu_port
    return cpu_port.search_timed(profile, pats, k, rc, text_addr, nbytes, threads)


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (restated, oracle/) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    profile, n_patterns, m, k, n, desc = WORKLOADS[args.workload]
    if args.text_bytes:
        n = args.text_bytes
    pats = make_patterns(profile, n_patterns, m)
    cores = os.cpu_count() or 1
    # bounded sample of the same workload: a prefix of the text (and of the pattern set)
    from oracle import cpu_port
    rate = cpu_port.calibrate(profile, pats[:min(len(pats), 32)], k, args.rc)  # bytes*patterns/s/thread
    sample_pats = pats[:min(len(pats), 64)]
    # each step is a bounded sample; the whole --steps/--warmup run stays within a few minutes
    budget = min(args.cpu_seconds, 150.0 / max(1, args.steps + args.warmup))
    sample_n = int(min(n, max(1 << 24, rate * cores * budget / max(1, len(sample_pats)))))
    rng = np.random.default_rng(42)
    text = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=sample_n, dtype=np.uint8)]
    for pos, q in plant_list(pats, sample_n, k, 64 if n_patterns == 1 else 2):
        text[pos:pos + len(q)] = np.frombuffer(q, dtype=np.uint8)
    text = np.ascontiguousarray(text)
    times = []
    nm = 0
    for it in range(args.warmup + args.steps):
        sec, nm, kind = cpu_port.search_timed(profile, sample_pats, k, args.rc, text.ctypes.data, sample_n, cores)
        if it >= args.warmup:
            times.append(sec)
    sec = sum(times) / len(times)
    # throughput of the FULL workload's metric: text bytes/s for the whole pattern set
    scale = len(sample_pats) / n_patterns
    value = sample_n / sec / 1e9 * scale
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": desc, "rc": bool(args.rc), "sample_bytes": sample_n, "sample_patterns": len(sample_pats)},
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": cores, "kind": "port",
                         "sample": f"{sample_n} text bytes x {len(sample_pats)} patterns per step ({kind})"},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "matches": nm,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------

def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    # the contract is ONE JSON line on stdout: libraries that write there (NCCL prints its version
    # banner at communicator creation) go to stderr while the benchmark runs
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    import sassy_b200
    from sassy_b200 import dist as sdist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    profile, n_patterns, m, k, n, desc = WORKLOADS[args.workload]
    if args.text_bytes:
        n = args.text_bytes
    if args.patterns:
        n_patterns = args.patterns
    all_pats = make_patterns(profile, n_patterns, m)
    copies = 64 if n_patterns == 1 else 2
    pshard = args.workload in PATTERN_SHARDED
    # pattern-sharded workloads: the same text on every rank, this rank's share of the patterns
    pats = [all_pats[i] for i in sdist.shard_indices(len(all_pats), rank, world)] if pshard else all_pats
    if not pats:
        raise SystemExit("fewer patterns than ranks")

    # ---- inputs: text (shard) of this rank, generated in HBM ------------------------------
    text_dev = synth_text_device(torch, n, 42 + (0 if pshard else rank), dev)
    for pos, q in plant_list(all_pats[:4096], n, k, copies, seed=44 + (0 if pshard else rank)):
        text_dev[pos:pos + len(q)] = torch.tensor(list(q), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()

    s = sassy_b200.Searcher(profile, rc=args.rc, device=local_rank)
    s.set_variant(args.variant)
    s.set_filter(args.filter)
    s.set_transport(args.transport)
    dt = s.text_from_device(text_dev.data_ptr(), n)
    enc = s.encode_patterns(pats) if n_patterns > 1 else None

    host = None
    if not args.no_e2e or not args.no_cpu:
        host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        host.copy_(text_dev)
        torch.cuda.synchronize()
    del text_dev
    torch.cuda.empty_cache()

    # N > 1: the match records of all ranks reach every rank through the fused peer-memory
    # exchange behind the traceback (csrc/peer_gather.cu); --gather nccl uses the host-staged
    # NCCL all-gather instead (A/B)
    pg = None
    if world > 1 and args.gather == "peer":
        pg = sdist.PeerGather.create_or_none(s, max_ops=m + k + 1, device=dev)

    def step_resident():
        if pg is not None:
            return pg.search_encoded(enc, dt, k) if enc is not None else pg.search(pats[0], dt, k)
        if enc is not None:
            ms = s.search_encoded_patterns(enc, dt, k)
        else:
            ms = s.search(pats[0], dt, k)
        if world > 1:
            ms = sdist.gather_matches(sdist.tag_rank(ms, rank), max_ops=m + k + 1, device=dev)
        return ms

    def step_e2e():
        buf = (host.data_ptr(), n)
        if enc is not None:
            ms = s.search_encoded_patterns(enc, buf, k)
        else:
            ms = s.search(pats[0], buf, k)
        if world > 1:
            ms = sdist.gather_matches(sdist.tag_rank(ms, rank), max_ops=m + k + 1, device=dev)
        return ms

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step_wall = []  # host wall time of every timed step (diagnostics: p50 / max next to the mean)

    def timed(fn, steps, warmup, sampler=None):
        for _ in range(warmup):
            ms = fn()
        barrier()
        scan_ms, total_ms, launches = [], [], 0
        step_wall.clear()
        t0 = time.perf_counter()
        tp = t0
        for it in range(steps):
            ms = fn()
            if sampler is not None:
                sampler.sample(it)
            tn = time.perf_counter()
            step_wall.append((tn - tp) * 1e3)
            tp = tn
            st = s.stats()
            scan_ms.append(st["scan_ms"])
            total_ms.append(st["total_ms"])
            launches += st["scan_launches"] + st["aux_launches"]
        barrier()
        el = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([el], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            el = float(t.item())
        return el, ms, scan_ms, total_ms, launches, s.stats()

    uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
    if not uuid.startswith("GPU-"):
        uuid = "GPU-" + uuid
    sampler = ClockSampler(uuid) if rank == 0 else None
    if sampler is not None:  # the first NVML queries of a process can take long: not inside the timed region
        sampler.sample(0)
        sampler.sm.clear(), sampler.power.clear(), sampler.reasons.clear()
        sampler._last = 0.0
    el, matches, scan_ms, total_ms, launches, st = timed(step_resident, args.steps, args.warmup, sampler)
    clocks = sampler.result() if rank == 0 else None
    wall_sorted = sorted(step_wall)
    step_stats = {"p50": wall_sorted[len(wall_sorted) // 2], "max": wall_sorted[-1],
                  "p90": wall_sorted[int(len(wall_sorted) * 0.9)]} if wall_sorted else None

    # text shards: N x n bytes scanned per step (weak scaling); pattern shards: the job is
    # "all patterns over the n-byte text", the same total work for every N (strong scaling)
    total_bytes = n if pshard else n * world
    total_patterns = len(all_pats)
    value = total_bytes * args.steps / el / 1e9
    ms_per_step = el / args.steps * 1e3

    e2e = None
    if not args.no_e2e:
        e_steps = max(2, min(args.steps, 5))
        e_el, e_matches, _, _, _, _ = timed(step_e2e, e_steps, 2)  # 2 warm-ups: buffers, pool and the pack-rate estimate settle
        if hasattr(e_matches, "records") and hasattr(matches, "records"):
            import numpy as np
            same = np.array_equal(e_matches.records, matches.records) and e_matches._ops == matches._ops
        else:
            same = sorted(x._key() for x in e_matches) == sorted(x._key() for x in matches)
        assert same, "host-pointer path and resident path disagree"
        tables = s.stats()["words"] * 4 * {"dna": 4, "iupac": 32, "ascii": 256}[profile] * len(pats) * (2 if args.rc else 1)
        est = s.stats()
        packed = bool(est["transfer_packed"])
        e2e = {"value": total_bytes * e_steps / e_el / 1e9, "unit": "GB/s",
               "h2d_bytes_per_step": est["transfer_bytes"] + tables + len(pats) * m,
               "transport": ("head of the text at 2 bits per character (packed by host threads inside the timed "
                             "region), tail as bytes, split so that packing and PCIe finish together") if packed else "bytes",
               "transfer_ms": est["transfer_ms"],
               "d2h_bytes_per_step": len(e_matches) // max(1, world) * (24 + 4 * ((m + k + 1 + 15) // 16)) + 16,
               "ms_per_step": e_el / e_steps * 1e3, "steps": e_steps}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    scan_avg_ms = sum(scan_ms) / len(scan_ms)
    achieved = n / (scan_avg_ms * 1e-3) / 1e9  # algorithmic bytes of one scan launch set = the text, read once
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(
            args.workload + ("_filter" if st["filter_words"] else "_scan"))
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                "kernel": "filter_kernel" if st["filter_words"] else "scan_kernel", "kernel_ms": scan_avg_ms,
                "verify_kernel_ms": st["verify_ms"], "prefilter_hits": st["hits"],
                "kernel_share_of_step": scan_avg_ms / ms_per_step,
                "algorithmic_bytes_per_launch": n,
                "lane_steps_per_s": n * len(pats) * (2 if args.rc else 1) / (scan_avg_ms * 1e-3)}

    cpu = None
    if not args.no_cpu:
        try:
            from oracle import cpu_port
            cores = os.cpu_count() or 1
            sample_pats = all_pats[:min(len(all_pats), 64)]
            rate = cpu_port.calibrate(profile, sample_pats[:32], k, args.rc)
            sample_n = int(min(n, max(1 << 24, rate * cores * args.cpu_seconds / len(sample_pats))))
            sec, nm, kind = cpu_port.search_timed(profile, sample_pats, k, args.rc, host.data_ptr(), sample_n, cores)
            cpu = {"value": sample_n / sec / 1e9 * len(sample_pats) / total_patterns, "unit": "GB/s", "cores": cores,
                   "kind": "port", "sample": f"first {sample_n} text bytes x {len(sample_pats)} patterns, {sec:.2f} s ({kind})",
                   "matches_in_sample": nm}
        except Exception as e:
            cpu = {"value": None, "unit": "GB/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}

    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong" if pshard else "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": desc, "text_bytes_per_gpu": n, "patterns": total_patterns,
                   "patterns_per_gpu": len(pats), "pattern_len": m, "k": k,
                   "rc": bool(args.rc), "mode": "search (local minima) + traceback",
                   "l2": "text (3 GB) is larger than L2 (126 MB); no flush needed", "variant": args.variant,
                   "row_bytes": st["row_bytes"], "rows": st["rows"], "blocks_per_sm": st["blocks_per_sm"],
                   "prefilter": {"mode": args.filter, "words": st["filter_words"], "piece_len": st["filter_len"],
                                 "fallback": st["filter_fallback"]},
                   "sharding": (("pattern shards (round-robin), text replicated; " if pshard else "text shards, one per rank; ")
                                + "match records of all ranks exchanged per step by "
                                + ("peer-memory stores over NVLink fused behind the traceback (%d NCCL fall-backs)" % pg.fallbacks
                                   if pg is not None else "a host-staged NCCL all-gather")) if world > 1 else "single GPU"},
        "matches": len(matches), "matches_per_s": len(matches) * args.steps / el,
        "gchar_pattern_per_s": (n * total_patterns if pshard else total_bytes * len(pats)) * args.steps / el / 1e9,
        "device_ms_per_step": sum(total_ms) / len(total_ms), "step_wall_ms": step_stats,
        "gpu_launches": launches, "clocks": clocks, "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu,
    }
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
