"""The reference's own assertions about the hot path (tests/golden/kat_props.json, 50+ cases from
src/search.rs, src/n_filter.rs, src/pattern_tiling/search.rs and src/profiles/*.rs), checked against
the oracle, the host emulation of the kernels' per-thread code, and -- on a B200 -- the CUDA path."""
import pytest

from tests import kat_props_util as kp

CASES = kp.load_cases()
IDS = [c["source"][:70] for c in CASES]


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_oracle_kat_props(case):
    kp.check(kp.oracle_runner, case)


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_emu_kat_props(case):
    kp.check(kp.emu_runner, case)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_gpu_kat_props(case):
    kp.check(kp.gpu_runner, case)


def test_enough_reference_vectors():
    from tests import kat_util
    assert len(CASES) + len(kat_util.load_cases()) >= 60
