"""CPU suite: the product's host logic (grid, priors, hyper-grid, lowering of the transition-model tree, wave /
averaging bookkeeping, accessors) driven through the C ABI into the CPU oracle, compared against golden vectors
generated from the unmodified reference (oracle/make_golden.py).  Also pins the oracle itself: every fixture passes
through oracle/blgrid_oracle.c here."""
import numpy as np
import pytest

import cases
import parity
from conftest import load_golden


@pytest.mark.parametrize('name', sorted(cases.CASES))
def test_case_matches_reference_golden(name, use_oracle):
    import bayesloop_b200 as bl
    S, got = parity.run_case(name, bl)
    parity.compare(name, got, load_golden(name), rtol=2e-9, atol_post=1e-13)


@pytest.mark.parametrize('name', sorted(cases.REFERENCE_PINNED_LOGE))
def test_golden_matches_reference_own_tests(name):
    """The fixtures themselves reproduce the numbers hard-coded in the reference's test-suite (App. B)."""
    want = cases.REFERENCE_PINNED_LOGE[name]
    got = float(load_golden(name)['logEvidence'])
    assert abs(got - want) <= 5e-6 * max(1.0, abs(want)), (name, got, want)


def test_golden_matches_reference_docs():
    for name, want in cases.REFERENCE_PINNED_LOG10E.items():
        got = float(load_golden(name)['logEvidence']) / np.log(10)
        assert abs(got - want) < 1e-5
