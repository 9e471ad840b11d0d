"""Adapter giving sassy_b200.Searcher (the CUDA path behind the C ABI) the same call
signature as the oracle, so the same known-answer and differential tests drive both."""
from __future__ import annotations

from typing import Sequence

from oracle import Match, OracleError
import sassy_b200


class GpuBackend:
    def __init__(self, variant: str = "tma", filter_mode: str = "auto"):
        self.variant = variant
        self.filter_mode = filter_mode
        self._cache = {}

    def _searcher(self, alphabet: str, rc: bool) -> sassy_b200.Searcher:
        key = (alphabet.lower(), bool(rc))
        if key not in self._cache:
            s = sassy_b200.Searcher(alphabet, rc=rc)
            s.set_variant(self.variant)
            s.set_filter(self.filter_mode)
            self._cache[key] = s
        return self._cache[key]

    @staticmethod
    def _conv(ms):
        return [Match(m.pattern_idx, m.text_start, m.text_end, m.pattern_start, m.pattern_end, m.cost, m.strand,
                      m.cigar) for m in ms]

    def search(self, alphabet, pattern: bytes, text: bytes, k: int, rc: bool = False, all_minima: bool = False):
        s = self._searcher(alphabet, rc)
        try:
            ms = s.search_all(pattern, text, k) if all_minima else s.search(pattern, text, k)
        except ValueError as e:
            raise OracleError(str(e))
        return self._conv(ms)

    def search_encoded(self, alphabet, patterns: Sequence[bytes], text: bytes, k: int, rc: bool = False,
                       all_minima: bool = False):
        s = self._searcher(alphabet, rc)
        try:
            enc = s.encode_patterns(list(patterns))
        except ValueError as e:
            raise OracleError(str(e))
        ms = s.search_all_encoded_patterns(enc, text, k) if all_minima else s.search_encoded_patterns(enc, text, k)
        return self._conv(ms)
