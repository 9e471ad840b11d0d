"""Stand-in for sassy_b200.Searcher backed by the CPU oracle -- TEST INFRASTRUCTURE ONLY.
Lets the host-side logic of sassy_b200/cli.py (ingestion, batching, ordering, TSV formatting)
run without a GPU, and gives the GPU test of the CLI its expected output."""
import re

import oracle
from sassy_b200.searcher import Match


def _conv(ms):
    out = []
    for x in ms:
        ops = "".join(ch * int(cnt) for cnt, ch in re.findall(r"(\d+)([=XID])", x.cigar))
        out.append(Match(x.pattern_idx, x.text_idx, x.text_start, x.text_end, x.pattern_start, x.pattern_end,
                         x.cost, x.strand, ops))
    return out


class OracleSearcher:
    def __init__(self, alphabet, rc=True, max_n_frac=None, alpha=None):
        self.alphabet, self.rc, self.max_n_frac, self.alpha = alphabet, rc, max_n_frac, alpha

    def search(self, p, t, k):
        return _conv(oracle.search(self.alphabet, p, t, k, rc=self.rc, max_n_frac=self.max_n_frac, alpha=self.alpha))

    def search_all(self, p, t, k):
        return _conv(oracle.search(self.alphabet, p, t, k, rc=self.rc, all_minima=True, max_n_frac=self.max_n_frac, alpha=self.alpha))

    def search_with_pam(self, p, t, k, pam, all_minima=True):
        return _conv(oracle.search(self.alphabet, p, t, k, rc=self.rc, all_minima=all_minima, pam=pam,
                                   max_n_frac=self.max_n_frac))

    def search_many(self, pats, texts, k, threads=0, mode="single"):
        return _conv(oracle.search_many(self.alphabet, pats, texts, k, rc=self.rc, max_n_frac=self.max_n_frac, alpha=self.alpha))

    def encode_patterns(self, pats):
        return list(pats)

    def search_encoded_patterns(self, enc, t, k):
        return _conv(oracle.search_encoded(self.alphabet, enc, t, k, rc=self.rc, max_n_frac=self.max_n_frac))


def make(alphabet, rc, max_n_frac, alpha=None):
    return OracleSearcher(alphabet, rc, max_n_frac, alpha)
