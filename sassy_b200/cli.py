"""`sassy search | filter | crispr` on the GPU path: FASTA/FASTQ ingestion, record batching,
ordered TSV / filtered-record output.

Mirrors the callers of the hot path in the reference's CLI (bin/grep.rs:28-143 arguments,
:390-586 batch loop, :586-760 output; bin/crispr.rs:15-262; bin/input_iterator.rs:76-201
batching; TSV columns README.md:199-253).  The coloured `grep` pretty printer is not part
of this path.  Every search goes through sassy_b200.Searcher (libsassy_b200.so); nothing
here computes alignments on the CPU.

    python -m sassy_b200.cli search -p ACGT -k 1 reads.fa [--filter out.fa]
    python -m sassy_b200.cli filter -f patterns.fa -k 2 reads.fq.gz [--search hits.tsv] [-v]
    python -m sassy_b200.cli crispr -g guides.txt -k 3 --max-n-frac 0.1 genome.fa
"""
from __future__ import annotations

import argparse
import gzip
import io
import sys
from typing import Callable, Iterable, Iterator, List, Optional, Sequence, Tuple

DEFAULT_BATCH_BYTES = 1024 * 1024   # bin/input_iterator.rs:6-7
DEFAULT_BATCH_PATTERNS = 64         # bin/input_iterator.rs:8-9
TSV_HEADER = "pat_id\ttext_id\tcost\tstrand\tstart\tend\tmatch_region\tcigar\n"          # bin/grep.rs:462-466
CRISPR_HEADER = "guide\ttext_id\tcost\tstrand\tstart\tend\tmatch_region\tcigar\n"        # bin/crispr.rs:157-160

# Profile::reverse_complement tables (src/profiles/dna.rs:121-133: upper case only;
# src/profiles/iupac.rs:235-278: both cases), bytes not in the table stay as they are.
_RC_DNA = bytes.maketrans(b"ACGT", b"TGCA")
_RC_IUPAC = bytes.maketrans(b"ACTGRYSWKMBDHVNXactgryswkmbdhvnx", b"TGACYRSWMKVHDBNXtgacyrswmkvhdbnx")


def reverse_complement(seq: bytes, alphabet: str) -> bytes:
    return seq.translate(_RC_DNA if alphabet == "dna" else _RC_IUPAC)[::-1]


class Record:
    __slots__ = ("id", "seq", "qual")

    def __init__(self, rid: str, seq: bytes, qual: bytes = b""):
        self.id, self.seq, self.qual = rid, seq, qual


def _open(path: str):
    if path in ("", "-"):
        raw = sys.stdin.buffer
    else:
        raw = open(path, "rb")
    head = raw.peek(2)[:2] if hasattr(raw, "peek") else b""
    if head == b"\x1f\x8b":
        return gzip.open(raw, "rb")
    return raw


def read_fastx(path: str) -> Iterator[Record]:
    """FASTA (multi-line) and FASTQ (4-line) records, optionally gzipped; the id is the whole
    header line, as needletail's `id()` gives it to the reference (bin/input_iterator.rs:128)."""
    with _open(path) as f:
        line = f.readline()
        while line:
            line = line.rstrip(b"\r\n")
            if not line:
                line = f.readline()
                continue
            if line[:1] == b">":
                rid = line[1:].decode("utf-8", "replace")
                parts = []
                line = f.readline()
                while line and line[:1] != b">":
                    parts.append(line.rstrip(b"\r\n"))
                    line = f.readline()
                yield Record(rid, b"".join(parts))
            elif line[:1] == b"@":
                rid = line[1:].decode("utf-8", "replace")
                seq = f.readline().rstrip(b"\r\n")
                f.readline()  # '+'
                qual = f.readline().rstrip(b"\r\n")
                yield Record(rid, seq, qual)
                line = f.readline()
            else:
                raise ValueError(f"{path}: not a FASTA/FASTQ record: {line[:40]!r}")


def text_batches(paths: Sequence[str], byte_limit: int = DEFAULT_BATCH_BYTES) -> Iterator[Tuple[str, List[Record]]]:
    """Batches of whole records of about byte_limit bytes (bin/input_iterator.rs:120-160)."""
    for path in paths:
        batch, size = [], 0
        for rec in read_fastx(path):
            batch.append(rec)
            size += len(rec.seq)
            if size >= byte_limit:
                yield path, batch
                batch, size = [], 0
        if batch:
            yield path, batch


def get_patterns(args) -> List[Record]:
    """bin/grep.rs:623-662."""
    if args.pattern is not None:
        return [Record("pattern", args.pattern.encode())]
    if args.pattern_file is not None:
        with open(args.pattern_file, "rb") as f:
            # BufReader::lines(): split at \n, a trailing \r is dropped, the last line needs no newline
            lines = f.read().split(b"\n")
            if lines and lines[-1] == b"":
                lines.pop()
            return [Record(str(i + 1), line[:-1] if line.endswith(b"\r") else line) for i, line in enumerate(lines)]
    if args.pattern_fasta is not None:
        return list(read_fastx(args.pattern_fasta))
    raise SystemExit("No --pattern, --pattern-file, or --pattern-fasta provided!")


def format_match_region(slice_: bytes, strand: str, alphabet: str, sam: bool) -> bytes:
    """bin/grep.rs:741-750."""
    if strand == "-" and not sam:
        return reverse_complement(slice_, alphabet)
    return slice_


def format_cigar(m, sam: bool) -> str:
    """bin/grep.rs:752-760: in SAM mode the CIGAR of an rc match is reversed (text direction)."""
    if m.strand == "-" and sam:
        from .searcher import _rle
        return _rle(m._ops[::-1])
    return m.cigar


def tsv_line(pat_id: str, text: Record, m, alphabet: str, sam: bool) -> str:
    """bin/grep.rs:714-739."""
    region = format_match_region(text.seq[m.text_start:m.text_end], m.strand, alphabet, sam)
    return (f"{pat_id}\t{text.id}\t{m.cost}\t{m.strand}\t{m.text_start}\t{m.text_end}\t"
            f"{region.decode('utf-8', 'replace')}\t{format_cigar(m, sam)}\n")


def write_record(rec: Record, out) -> None:
    """bin/grep.rs:664-674."""
    if rec.qual:
        out.write(f"@{rec.id}\n{rec.seq.decode('utf-8', 'replace')}\n+\n{rec.qual.decode('utf-8', 'replace')}\n")
    else:
        out.write(f">{rec.id}\n{rec.seq.decode('utf-8', 'replace')}\n")


def _default_searcher(alphabet: str, rc: bool, max_n_frac: Optional[float], alpha: Optional[float] = None):
    import sassy_b200
    return sassy_b200.Searcher(alphabet, rc=rc, alpha=alpha, max_n_frac=max_n_frac)


def run_search(args, match_out=None, filter_out=None, make_searcher: Callable = _default_searcher) -> List[int]:
    """The batch loop of bin/grep.rs:390-586 (`search` and `filter`).  Returns the cost histogram."""
    patterns = get_patterns(args)
    if not patterns:
        raise SystemExit("No pattern sequences found")
    k = args.k
    rc = not args.no_rc
    # only the Iupac searcher gets the N filter (bin/grep.rs:489-497)
    searcher = make_searcher(args.alphabet, rc, args.max_n_frac if args.alphabet == "iupac" else None,
                             **({"alpha": args.overhang} if args.overhang is not None else {}))
    hist = [0] * (k + 1)
    if match_out is not None:
        match_out.write(TSV_HEADER)
    pbs = args.pattern_batch_size or DEFAULT_BATCH_PATTERNS
    pattern_batches = [patterns[i:i + pbs] for i in range(0, len(patterns), pbs)]
    encoded = None
    if args.v2:
        encoded = [searcher.encode_patterns([p.seq for p in pb]) for pb in pattern_batches]
    for path, texts in text_batches(args.paths or [""]):
        for bi, pb in enumerate(pattern_batches):
            per_text: List[list] = [[] for _ in texts]
            if args.v2:  # bin/grep.rs:358-388
                for ti, t in enumerate(texts):
                    for m in searcher.search_encoded_patterns(encoded[bi], t.seq, k):
                        per_text[ti].append((pb[m.pattern_idx], m))
            else:        # every pattern of the batch against every record: one search_many call
                for m in searcher.search_many([p.seq for p in pb], [t.seq for t in texts], k):
                    per_text[m.text_idx].append((pb[m.pattern_idx], m))
            for ti, t in enumerate(texts):
                ms = per_text[ti]
                # search_many is pattern-major per text, like the reference's loop (:520-536)
                ms.sort(key=lambda pm: pm[1].text_start)  # stable, bin/grep.rs:596
                for _, m in ms:
                    hist[m.cost] += 1
                if filter_out is not None and (bool(ms) != bool(args.invert)):
                    write_record(t, filter_out)
                if match_out is not None:
                    for p, m in ms:
                        match_out.write(tsv_line(p.id, t, m, args.alphabet, args.sam))
    return hist


def read_guides(path: str) -> List[bytes]:
    """bin/crispr.rs:125-134: one guide per line, empty lines dropped."""
    with open(path, "rb") as f:
        return [ln for ln in (x.rstrip(b"\r\n") for x in f.read().split(b"\n")) if ln]


def run_crispr(args, out, make_searcher: Callable = _default_searcher, log=sys.stderr) -> int:
    """bin/crispr.rs:145-262: IUPAC, rc unless --no-rc, N filter, PAM (the last pam_length
    characters of every guide) has to match exactly unless --allow-pam-edits."""
    guides = read_guides(args.guide)
    if not (0.0 <= args.max_n_frac <= 1.0):
        raise SystemExit("[N-chars] Error: max_n_frac must be between 0 and 1.0")
    if not guides:
        raise SystemExit("[PAM] Error: No guide sequences provided, please check your input file "
                         "(one guide sequence per line)")
    pam = guides[0][len(guides[0]) - args.pam_length:]
    for g in guides:
        if g[len(g) - args.pam_length:] != pam:
            raise SystemExit("[PAM] One of the guide sequences has a PAM different than the provided PAM")
    print(f"[PAM] PAM used to filter: {pam.decode()}", file=log)
    print(f"[PAM] Edits in PAM are allowed: {str(bool(args.allow_pam_edits)).lower()}", file=log)
    searcher = make_searcher("iupac", not args.no_rc, args.max_n_frac)
    out.write(CRISPR_HEADER)
    total = 0
    for path, texts in text_batches([args.path]):
        for t in texts:
            dt = searcher.upload_text(t.seq) if hasattr(searcher, "upload_text") else t.seq
            for g in guides:
                if args.allow_pam_edits:
                    ms = searcher.search_all(g, dt, args.k)
                else:
                    ms = searcher.search_with_pam(g, dt, args.k, pam, all_minima=True)
                total += len(ms)
                gs = g.decode("utf-8", "replace")
                for m in ms:
                    out.write(tsv_line(gs, t, m, "iupac", False))
            if hasattr(dt, "free"):
                dt.free()
    print(f"  Total targets found:   {total}", file=log)
    return total


def run_selfcheck(out=sys.stderr) -> int:
    """`sassy test` (src/lib.rs:187-281 prints CPU features and a throughput figure): the device the
    library would use, the reference's doctest as a known answer, and search throughput."""
    import ctypes
    import random
    import time
    from . import _native, Searcher
    lib = _native.load()
    n_dev = lib.sassy_gpu_device_count()
    print(f"libsassy_b200: {_native.LIB_PATH}", file=out)
    print(f"CUDA devices: {n_dev}", file=out)
    if n_dev == 0:
        print("no CUDA device: this library has no CPU fallback", file=out)
        return 1
    for d in range(n_dev):
        name = ctypes.create_string_buffer(128)
        sms, mhz, cmaj, cmin = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        mem = ctypes.c_size_t()
        lib.sassy_gpu_device_info(d, name, 128, ctypes.byref(sms), ctypes.byref(mhz), ctypes.byref(mem),
                                  ctypes.byref(cmaj), ctypes.byref(cmin))
        ok = "+" if cmaj.value >= 10 else "- (needs sm_100a)"
        print(f"  [{d}] {name.value.decode()}  sm_{cmaj.value}{cmin.value} {ok}  {sms.value} SMs  {mhz.value} MHz  "
              f"{mem.value / 2**30:.0f} GiB", file=out)
    s = Searcher("dna", rc=True)
    got = [(m.text_start, m.text_end, m.cost, m.strand, m.cigar) for m in s.search(b"ATCG", b"CCCATCACCC", 1)]
    want = [(3, 7, 1, "+", "3=1X"), (1, 5, 1, "-", "2=1X1=")]  # src/lib.rs:72-107
    print(f"known answer (src/lib.rs doctest): {'ok' if got == want else 'MISMATCH ' + repr(got)}", file=out)
    rng = random.Random(1)
    n = 1 << 28
    block = bytes(rng.choice(b"ACGT") for _ in range(1 << 16))
    text = block * (n // len(block))
    pattern = bytes(rng.choice(b"ACGT") for _ in range(23))
    si = Searcher("iupac", rc=False)
    dt = si.upload_text(text)
    si.search(pattern, dt, 1)
    t0 = time.perf_counter()
    for _ in range(10):
        si.search(pattern, dt, 1)
    el = (time.perf_counter() - t0) / 10
    print(f"23 bp pattern, k = 1, {n >> 20} MiB text resident in HBM: {n / el / 1e9:.0f} GB/s", file=out)
    t0 = time.perf_counter()
    si.search(pattern, text, 1)
    el = time.perf_counter() - t0
    print(f"same search from a host buffer (copy included): {n / el / 1e9:.1f} GB/s", file=out)
    return 0 if got == want else 1


def _base_args(p: argparse.ArgumentParser) -> None:
    p.add_argument("-p", "--pattern")
    p.add_argument("-l", "--pattern-file")
    p.add_argument("-f", "--pattern-fasta")
    p.add_argument("--pattern-batch-size", type=int)
    p.add_argument("-k", type=int, required=True)
    p.add_argument("-a", "--alphabet", choices=["dna", "iupac"], default="iupac")
    p.add_argument("--overhang", type=float)
    p.add_argument("--no-rc", action="store_true")
    p.add_argument("--max-n-frac", type=float, default=0.2)
    p.add_argument("--v2", action="store_true")
    p.add_argument("-j", "--threads", type=int, help="accepted for compatibility; the GPU does the work")
    p.add_argument("-v", "--invert", action="store_true")
    p.add_argument("--sam", action="store_true")
    p.add_argument("paths", nargs="*")


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(prog="sassy_b200", description=__doc__.split("\n\n")[0])
    sub = ap.add_subparsers(dest="cmd", required=True)
    s = sub.add_parser("search", help="TSV of all matches on stdout")
    _base_args(s)
    s.add_argument("--filter", nargs="?", const="-", help="also write the matching records here")
    f = sub.add_parser("filter", help="matching (or with -v non-matching) records on stdout")
    _base_args(f)
    f.add_argument("--search", nargs="?", const="-", help="also write the TSV of all matches here")
    sub.add_parser("test", help="self-check: device, known answer, throughput")
    c = sub.add_parser("crispr", help="CRISPR off-target search")
    c.add_argument("-g", "--guide", required=True)
    c.add_argument("-k", "--k", type=int, required=True)
    c.add_argument("-o", "--output")
    c.add_argument("--max-n-frac", type=float, required=True)
    c.add_argument("-j", "--threads", type=int)
    c.add_argument("--pam-length", type=int, default=3)
    c.add_argument("--allow-pam-edits", action="store_true")
    c.add_argument("--no-rc", action="store_true")
    c.add_argument("path")
    return ap


def _writer(path: Optional[str]):
    if path is None:
        return None, False
    if path in ("", "-"):
        return sys.stdout, False
    return open(path, "w"), True


def main(argv: Optional[Sequence[str]] = None, make_searcher: Callable = _default_searcher) -> int:
    args = build_parser().parse_args(argv)
    if args.cmd == "test":
        return run_selfcheck()
    if args.cmd == "crispr":
        out, close = _writer(args.output or "-")
        try:
            run_crispr(args, out, make_searcher)
        finally:
            if close:
                out.close()
        return 0
    if args.cmd == "search":
        match_out, c1 = sys.stdout, False
        filter_out, c2 = _writer(args.filter)
    else:
        filter_out, c2 = sys.stdout, False
        match_out, c1 = _writer(args.search)
    try:
        hist = run_search(args, match_out, filter_out, make_searcher)
    finally:
        if c1:
            match_out.close()
        if c2:
            filter_out.close()
    print("\nStatistics", file=sys.stderr)
    for d, c in enumerate(hist):
        print(f"  dist {d}: {c}", file=sys.stderr)
    print(f"  total: {sum(hist)}", file=sys.stderr)
    return 0


if __name__ == "__main__":
    sys.exit(main())
