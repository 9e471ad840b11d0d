"""The CLI on the CUDA path: same bytes on stdout as with the oracle-backed searcher."""
import random

import pytest

from tests import cli_backend
from tests.test_cli import _crispr_counts, make_inputs, run

pytestmark = pytest.mark.gpu


def gpu_make(alphabet, rc, max_n_frac, alpha=None):
    import sassy_b200
    return sassy_b200.Searcher(alphabet, rc=rc, alpha=alpha, max_n_frac=max_n_frac)


def test_cli_search_filter_on_gpu(tmp_path):
    rng = random.Random(61)
    pats, recs, fa, pf = make_inputs(tmp_path, rng, n_rec=40, n_pat=7)
    dpats, drecs, dfa, dpf = make_inputs(tmp_path, rng, n_rec=40, n_pat=7, n_prob=0.0, tag="_dna")
    for argv in (["search", "-f", pf, "-k", "2", fa],
                 ["search", "-f", dpf, "-k", "2", "-a", "dna", "--sam", dfa],
                 ["search", "-f", pf, "-k", "1", "--no-rc", "--pattern-batch-size", "3", fa],
                 ["filter", "-f", pf, "-k", "1", "-v", fa],
                 ["search", "-f", pf, "-k", "3", "--overhang", "0.5", fa],
                 ["search", "-p", pats[2][1].decode(), "-k", "3", "--max-n-frac", "0.0", fa]):
        assert run(argv, gpu_make) == run(argv, cli_backend.make), argv
    a = run(["search", "-f", pf, "-k", "2", "--v2", fa], gpu_make)
    b = run(["search", "-f", pf, "-k", "2", "--v2", fa], cli_backend.make)
    assert sorted(a.splitlines()) == sorted(b.splitlines()) and len(a.splitlines()) > 5


def test_cli_crispr_on_gpu(tmp_path):
    """bin/crispr.rs:264-362 through the CUDA path."""
    assert _crispr_counts(tmp_path, allow_pam_edits=True, make=gpu_make) == {"exact": 2, "pam_mutated": 1, "n_frac": 2}
    assert _crispr_counts(tmp_path, make=gpu_make) == {"exact": 1, "pam_mutated": 0, "n_frac": 1}
    thr = 3.0 / 17.0
    assert _crispr_counts(tmp_path, max_n_frac=thr + 0.01, make=gpu_make)["n_frac"] == 1
    assert _crispr_counts(tmp_path, max_n_frac=thr - 0.01, make=gpu_make)["n_frac"] == 0
