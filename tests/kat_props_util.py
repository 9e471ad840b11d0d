"""Driver for tests/golden/kat_props.json: assertions of the reference's own tests restated as data
(tools/make_kat_props.py).  A runner is a callable
    run(alphabet, rc, api, k, pattern | patterns, text, alpha=None, max_n_frac=None, pam=None) -> [Match]
with Match fields as the reference's (src/search.rs:35-62), strand "+" / "-", cigar run-length."""
import json
import os

from tests.kat_util import build_text, to_path

HERE = os.path.dirname(os.path.abspath(__file__))


def load_cases():
    with open(os.path.join(HERE, "golden", "kat_props.json")) as f:
        return json.load(f)["cases"]


def _run(runner, case, **over):
    c = dict(case)
    c.update(over)
    text = c["text"].encode() if isinstance(c.get("text"), str) else build_text(c)
    kw = dict(alpha=c.get("alpha"), max_n_frac=c.get("max_n_frac"), pam=c["pam"].encode() if c.get("pam") else None)
    if c["api"].startswith("encoded"):
        return runner(c["alphabet"], c["rc"], c["api"], c["k"], [p.encode() for p in c["patterns"]], text, **kw)
    return runner(c["alphabet"], c["rc"], c["api"], c["k"], c["pattern"].encode(), text, **kw)


def _matches(m, exp):
    for key, val in exp.items():
        if key == "cost_le":
            if m.cost > val:
                return False
        elif key == "path":
            if to_path(m) != val:
                return False
        elif getattr(m, key) != val:
            return False
    return True


def check(runner, case):
    ms = _run(runner, case)
    src = case["source"]
    for key, val in case["expect"].items():
        if key == "nonempty":
            assert len(ms) > 0, src
        elif key == "empty":
            assert len(ms) == 0, (src, ms)
        elif key == "count":
            assert len(ms) == val, (src, ms)
        elif key == "exists_text_start_within":
            assert any(abs(m.text_start - val[0]) <= val[1] for m in ms), (src, ms)
        elif key == "none_text_start_within":
            assert not any(abs(m.text_start - val[0]) <= val[1] for m in ms), (src, ms)
        elif key == "all_text_start_ge":
            assert all(m.text_start >= val for m in ms), (src, ms)
        elif key == "all_text_start_lt":
            assert all(m.text_start < val for m in ms), (src, ms)
        elif key == "all_cost_le":
            assert all(m.cost <= val for m in ms), (src, ms)
        elif key == "text_ends":
            assert [m.text_end for m in ms] == val, (src, ms)
        elif key == "text_starts":
            assert [m.text_start for m in ms] == val, (src, ms)
        elif key == "starts_strands":
            assert [[m.text_start, m.strand] for m in ms] == val, (src, ms)
        elif key == "contains":
            for exp in val:
                assert any(_matches(m, exp) for m in ms), (src, exp, ms)
        elif key == "first":
            assert ms and _matches(ms[0], val), (src, ms)
        elif key == "same_count_as":
            other = _run(runner, case, **val)
            assert len(ms) == len(other), (src, ms, other)
        elif key == "same_coords_as":
            other = _run(runner, case, **val)
            a = sorted((m.text_start, m.text_end, m.cost) for m in ms)
            b = sorted((m.text_start, m.text_end, m.cost) for m in other)
            assert a == b, (src, a, b)
        elif key == "first_cigar_equals":
            other = _run(runner, case, **val)
            assert ms and other and ms[0].cigar == other[0].cigar, (src, ms, other)
        else:
            raise AssertionError(f"unknown expectation {key}")


# ---- runners -----------------------------------------------------------------------------------

def oracle_runner(alphabet, rc, api, k, pat, text, alpha=None, max_n_frac=None, pam=None):
    import oracle
    if api.startswith("encoded"):
        return oracle.search_encoded(alphabet, pat, text, k, rc=rc, all_minima=(api == "encoded_all"),
                                     max_n_frac=max_n_frac, alpha=alpha)
    return oracle.search(alphabet, pat, text, k, rc=rc, all_minima=(api == "search_all"), max_n_frac=max_n_frac,
                         pam=pam, alpha=alpha)


def emu_runner(alphabet, rc, api, k, pat, text, alpha=None, max_n_frac=None, pam=None):
    from tests.emu_backend import EmuBackend
    b = EmuBackend(use_filter=-1)
    if api.startswith("encoded"):
        return b.search_encoded(alphabet, pat, text, k, rc=rc, all_minima=(api == "encoded_all"), alpha=alpha,
                                max_n_frac=max_n_frac)
    return b.search_opts(alphabet, pat, text, k, rc=rc, all_minima=(api == "search_all"), max_n_frac=max_n_frac,
                         pam=pam, alpha=alpha)


def gpu_runner(alphabet, rc, api, k, pat, text, alpha=None, max_n_frac=None, pam=None):
    import sassy_b200
    s = sassy_b200.Searcher(alphabet, rc=rc, alpha=alpha, max_n_frac=max_n_frac)
    allm = api in ("search_all", "encoded_all")
    if api.startswith("encoded"):
        enc = s.encode_patterns(pat)
        return (s.search_all_encoded_patterns if allm else s.search_encoded_patterns)(enc, text, k)
    if pam:
        return s.search_with_pam(pat, text, k, pam, all_minima=allm)
    return (s.search_all if allm else s.search)(pat, text, k)
