"""Host-side logic of the CLI (sassy_b200/cli.py): FASTA/FASTQ ingestion, batching, output order
and TSV formatting, driven on the CPU through an oracle-backed searcher; the reference's own CLI
tests (bin/grep.rs:795-813 sam_output, bin/crispr.rs:264-362 test_crispr) are the known answers."""
import contextlib
import gzip
import io
import random

import oracle
from sassy_b200 import cli
from tests import cli_backend
from tests.test_oracle_props import rand_seq


def run(argv, make=cli_backend.make):
    out = io.StringIO()
    err = io.StringIO()
    with contextlib.redirect_stdout(out), contextlib.redirect_stderr(err):
        cli.main(argv, make_searcher=make)
    return out.getvalue()


def test_sam_output_formatting():
    """bin/grep.rs:795-813."""
    assert cli.format_match_region(b"AAGT", "-", "dna", False) == oracle.reverse_complement("dna", b"AAGT")
    assert cli.format_match_region(b"AAGT", "-", "dna", True) == b"AAGT"
    from sassy_b200.searcher import Match
    ops = "==XDDD"  # 2=1X3D
    rc = Match(0, 0, 0, 0, 0, 0, 0, "-", ops)
    fw = Match(0, 0, 0, 0, 0, 0, 0, "+", ops)
    assert cli.format_cigar(rc, False) == "2=1X3D"
    assert cli.format_cigar(rc, True) == "3D1X2="
    assert cli.format_cigar(fw, True) == "2=1X3D"
    assert cli.format_cigar(fw, False) == "2=1X3D"


def test_fastx_reader(tmp_path):
    fa = tmp_path / "a.fa"
    fa.write_text(">r1 first record\nACGT\nAC\n\n>r2\nGG\n")
    recs = list(cli.read_fastx(str(fa)))
    assert [(r.id, r.seq, r.qual) for r in recs] == [("r1 first record", b"ACGTAC", b""), ("r2", b"GG", b"")]
    fq = tmp_path / "a.fq.gz"
    with gzip.open(fq, "wb") as f:
        f.write(b"@q1\nACGT\n+\nIIII\n@q2\nTT\n+\n!!\n")
    recs = list(cli.read_fastx(str(fq)))
    assert [(r.id, r.seq, r.qual) for r in recs] == [("q1", b"ACGT", b"IIII"), ("q2", b"TT", b"!!")]
    # batching: whole records, about byte_limit bytes per batch
    big = tmp_path / "b.fa"
    big.write_text("".join(f">t{i}\n{'A' * 400}\n" for i in range(10)))
    sizes = [len(b) for _, b in cli.text_batches([str(big)], byte_limit=1000)]
    assert sizes == [3, 3, 3, 1]


def _expected_tsv(pats, recs, k, alphabet, rc, sam, max_n_frac):
    lines = [cli.TSV_HEADER]
    for rid, seq in recs:
        ms = []
        for pid, p in pats:
            for m in oracle.search(alphabet, p, seq, k, rc=rc, max_n_frac=max_n_frac):
                ms.append((pid, m))
        ms.sort(key=lambda pm: pm[1].text_start)
        for pid, m in ms:
            region = seq[m.text_start:m.text_end]
            cigar = m.cigar
            if m.strand == "-":
                if sam:
                    import re
                    cigar = "".join(f"{c}{o}" for c, o in reversed(re.findall(r"(\d+)([=XID])", m.cigar)))
                else:
                    region = oracle.reverse_complement(alphabet, region)
            lines.append(f"{pid}\t{rid}\t{m.cost}\t{m.strand}\t{m.text_start}\t{m.text_end}\t{region.decode()}\t{cigar}\n")
    return "".join(lines)


def make_inputs(tmp_path, rng, n_rec=12, n_pat=5, m=16, n_prob=0.3, tag=""):
    pats = [(f"p{i}", rand_seq(rng, m)) for i in range(n_pat)]
    recs = []
    for i in range(n_rec):
        t = bytearray(rand_seq(rng, rng.randrange(50, 600)))
        for _ in range(rng.randrange(0, 3)):
            pid, p = pats[rng.randrange(n_pat)]
            q = bytearray(p if rng.random() < 0.5 else oracle.reverse_complement("dna", p))
            if rng.random() < 0.5:
                q[rng.randrange(m)] = ord("N") if rng.random() < n_prob else ord("A")
            a = rng.randrange(0, len(t) - m)
            t[a:a + m] = q
        recs.append((f"rec{i} len={len(t)}", bytes(t)))
    fa = tmp_path / f"texts{tag}.fa"
    fa.write_text("".join(f">{rid}\n{seq.decode()}\n" for rid, seq in recs))
    pf = tmp_path / f"pats{tag}.fa"
    pf.write_text("".join(f">{pid}\n{p.decode()}\n" for pid, p in pats))
    return pats, recs, str(fa), str(pf)


def test_search_tsv_matches_reference_loop(tmp_path):
    rng = random.Random(51)
    pats, recs, fa, pf = make_inputs(tmp_path, rng)
    # Dna texts must stay inside ACGT (an N ends in the reference's "Trace failed" panic)
    dpats, drecs, dfa, dpf = make_inputs(tmp_path, rng, n_prob=0.0, tag="_dna")
    for sam in (False, True):
        argv = ["search", "-f", pf, "-k", "2", "-a", "iupac", fa] + (["--sam"] if sam else [])
        assert run(argv) == _expected_tsv(pats, recs, 2, "iupac", True, sam, 0.2)
        argv = ["search", "-f", dpf, "-k", "2", "-a", "dna", dfa] + (["--sam"] if sam else [])
        assert run(argv) == _expected_tsv(dpats, drecs, 2, "dna", True, sam, None)
    # single inline pattern, forward only, tiny pattern batches give the same lines per record
    argv = ["search", "-p", pats[0][1].decode(), "-k", "1", "--no-rc", fa]
    assert run(argv) == _expected_tsv([("pattern", pats[0][1])], recs, 1, "iupac", False, False, 0.2)


def test_filter_and_invert(tmp_path):
    rng = random.Random(52)
    pats, recs, fa, pf = make_inputs(tmp_path, rng)
    hit = [bool(_expected_tsv(pats, [r], 1, "iupac", True, False, 0.2) != cli.TSV_HEADER) for r in recs]
    assert any(hit) and not all(hit)
    out = run(["filter", "-f", pf, "-k", "1", fa])
    assert out == "".join(f">{rid}\n{seq.decode()}\n" for (rid, seq), h in zip(recs, hit) if h)
    out = run(["filter", "-f", pf, "-k", "1", "-v", fa])
    assert out == "".join(f">{rid}\n{seq.decode()}\n" for (rid, seq), h in zip(recs, hit) if not h)


def test_v2_flag_uses_encoded_search(tmp_path):
    rng = random.Random(53)
    pats, recs, fa, pf = make_inputs(tmp_path, rng, n_pat=9)
    out = run(["search", "-f", pf, "-k", "2", "--v2", fa])
    lines = [cli.TSV_HEADER]
    for rid, seq in recs:
        ms = oracle.search_encoded("iupac", [p for _, p in pats], seq, 2, rc=True, max_n_frac=0.2)
        ms = sorted(ms, key=lambda m: m.text_start)
        for m in ms:
            region = seq[m.text_start:m.text_end]
            if m.strand == "-":
                region = oracle.reverse_complement("iupac", region)
            lines.append(f"{pats[m.pattern_idx][0]}\t{rid}\t{m.cost}\t{m.strand}\t{m.text_start}\t{m.text_end}\t"
                         f"{region.decode()}\t{m.cigar}\n")
    assert sorted(out.splitlines()) == sorted("".join(lines).splitlines())


def _crispr_counts(tmp_path, **kw):
    """bin/crispr.rs:264-362 (test_crispr)."""
    g = tmp_path / "guides.txt"
    g.write_text("TAGCATCAGCTACGNGG\n")
    t = tmp_path / "targets.fa"
    t.write_text(">exact\nTAGCATCAGCTACGAGG\n>pam_mutated\nTAGCATCAGCTACGACG\n>n_frac\nTNNNATCAGCTACGAGG\n")
    argv = ["crispr", "-g", str(g), "-k", "1", "--max-n-frac", str(kw.get("max_n_frac", 1.0)), str(t)]
    if kw.get("allow_pam_edits"):
        argv.append("--allow-pam-edits")
    out = run(argv, kw.get("make", cli_backend.make))
    lines = out.splitlines()
    assert lines[0] + "\n" == cli.CRISPR_HEADER
    counts = {"exact": 0, "pam_mutated": 0, "n_frac": 0}
    for ln in lines[1:]:
        counts[ln.split("\t")[1]] += 1
    return counts


def test_crispr_reference_counts(tmp_path):
    assert _crispr_counts(tmp_path, allow_pam_edits=True) == {"exact": 2, "pam_mutated": 1, "n_frac": 2}
    assert _crispr_counts(tmp_path) == {"exact": 1, "pam_mutated": 0, "n_frac": 1}
    thr = 3.0 / 17.0
    assert _crispr_counts(tmp_path, max_n_frac=thr + 0.01)["n_frac"] == 1
    assert _crispr_counts(tmp_path, max_n_frac=thr - 0.01)["n_frac"] == 0


def test_overhang_flag(tmp_path):
    """--overhang (bin/grep.rs:76-78) reaches the searcher: a pattern hanging over the end of a read."""
    fa = tmp_path / "reads.fa"
    fa.write_text(">r1\nGGGGGGGATCG\n>r2\nATCGGGGGGGGGG\n")
    out = run(["search", "-p", "ATCGATCG", "-k", "2", "--no-rc", "--overhang", "0.5", str(fa)])
    assert out == (cli.TSV_HEADER + "pattern\tr1\t2\t+\t7\t11\tATCG\t4=\n" + "pattern\tr2\t2\t+\t0\t4\tATCG\t4=\n")
    assert run(["search", "-p", "ATCGATCG", "-k", "2", "--no-rc", str(fa)]) == cli.TSV_HEADER
