"""Shared driver for the known-answer vectors in tests/golden/kat.json.

`backend` is any object with
    search(alphabet, pattern, text, k, rc, all_minima) -> [Match]
    search_encoded(alphabet, patterns, text, k, rc, all_minima) -> [Match]
where Match has the reference's field names (src/search.rs:35-62) and
`strand` in {"+", "-"}, `cigar` as the run-length string.
"""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def load_cases():
    with open(os.path.join(HERE, "golden", "kat.json")) as f:
        return json.load(f)["cases"]


def build_text(case) -> bytes:
    if "text" in case:
        return case["text"].encode()
    e = case["text_expr"]
    t = bytearray(e["fill"].encode() * e["len"])
    for pos, s in e.get("splice", []):
        t[pos:pos] = s.encode()  # Vec::splice(pos..pos, ..) inserts
    for pos, s in e.get("overwrite", []):
        t[pos:pos + len(s)] = s.encode()  # Vec::splice(pos..pos+len, ..) replaces
    return bytes(t)


def run(backend, case):
    text = build_text(case)
    api = case["api"]
    if api in ("search", "search_all"):
        return backend.search(case["alphabet"], case["pattern"].encode(), text, case["k"],
                              rc=case["rc"], all_minima=(api == "search_all"))
    pats = [p.encode() for p in case["patterns"]]
    return backend.search_encoded(case["alphabet"], pats, text, case["k"], rc=case["rc"],
                                  all_minima=(api == "encoded_all"))


def to_path(m):
    """Match::to_path, src/search.rs:83-104."""
    import re
    if m.strand == "-":
        pos = [m.pattern_start, m.text_end - 1]
        sign = -1
    else:
        pos = [m.pattern_start, m.text_start]
        sign = 1
    path = [list(pos)]
    for cnt, op in re.findall(r"(\d+)([=XID])", m.cigar):
        for _ in range(int(cnt)):
            if op in "=X":
                pos = [pos[0] + 1, pos[1] + sign]
            elif op == "I":
                pos = [pos[0] + 1, pos[1]]
            else:
                pos = [pos[0], pos[1] + sign]
            path.append(list(pos))
    path.pop()
    return path


def _fields_match(m, exp):
    for key, val in exp.items():
        if key == "path":
            if to_path(m) != val:
                return False
        elif getattr(m, key) != val:
            return False
    return True


def check(backend, case):
    ms = run(backend, case)
    mode = case["expect_mode"]
    if "expect_count" in case:
        assert len(ms) == case["expect_count"], (case["source"], ms)
    if mode == "count":
        return
    if mode == "exact_list":
        assert len(ms) == len(case["expect"]), (case["source"], ms)
        for m, exp in zip(ms, case["expect"]):
            assert _fields_match(m, exp), (case["source"], m, exp)
    elif mode == "first":
        assert ms and _fields_match(ms[0], case["expect"][0]), (case["source"], ms)
    elif mode == "contains":
        for exp in case["expect"]:
            assert any(_fields_match(m, exp) for m in ms), (case["source"], exp, ms)
    elif mode == "encoded_by_pattern":
        for exp in case["expect"]:
            cand = [m for m in ms if m.pattern_idx == exp["pattern_idx"]]
            assert len(cand) == 1 and _fields_match(cand[0], exp), (case["source"], exp, ms)
    elif mode == "ends":
        assert sorted(m.text_end for m in ms) == case["expect_ends"], (case["source"], ms)
    elif mode == "same_first_cigar_as":
        other = run(backend, case["other"])
        assert ms and other and ms[0].cigar == other[0].cigar, (case["source"], ms, other)
    elif mode == "first_cigar_differs_from":
        other = run(backend, case["other"])
        assert ms and other and ms[0].cigar != other[0].cigar, (case["source"], ms, other)
    elif mode == "count_differs_from":
        other = run(backend, case["other"])
        assert len(ms) != len(other), (case["source"], ms, other)
    else:
        raise AssertionError(f"unknown expect_mode {mode}")
