"""The reference's own C caller (c/example.c, unmodified, compiled by __graft_entry__.build() in the
build container where the reference is mounted) runs against libsassy_b200.so and prints what the
oracle computes for its inputs.  Only the prebuilt binary oracle/_ref/example_c is used here."""
import os
import subprocess

import pytest

import oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "example_c")


def test_reference_c_example_output():
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/example_c was not built (no /root/reference where build() ran)")
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    # inputs of c/example.c:13-24: dna, rc = true, k = 1
    want = oracle.search("dna", b"AAGGGGA", b"CCCCCCCCCAAGGGGACCCCCAAGGCGACCCCCCCCC", 1, rc=True)
    lines = [f"Found {len(want)} match(es):"]
    for i, m in enumerate(want):  # print_match, c/example.c:7-10
        lines.append(f"#{i}  pat[{m.pattern_start}-{m.pattern_end}]  txt[{m.text_start}-{m.text_end}]  "
                     f"cost={m.cost}  strand={m.strand}")
    assert out.stdout.strip().split("\n") == lines
    assert len(want) == 2
