"""Pin the CPU oracle against the reference's own known-answer vectors."""
import pytest

import oracle
from tests import kat_util


@pytest.mark.parametrize("case", kat_util.load_cases(), ids=lambda c: c["source"][:60])
def test_oracle_kat(case):
    kat_util.check(oracle, case)
