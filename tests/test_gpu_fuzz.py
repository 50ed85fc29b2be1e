"""Random operator programs on the B200 against the CPU oracle: the case generator of tests/fuzz_cases.py (random
trees of Combined / Serial models with change- and break-points, hyper-parameter lists, missing data, all four study
types) run through the CUDA library and through the oracle, same host logic on both sides.

300 cases by default under `-m gpu` (BLG_TEST_FUZZ=<n> changes the count): a few seconds on the B200 box."""
import contextlib
import io
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_cases(first, second, n, seed0=0):
    """Runs n random cases on both engines; returns the list of (seed, label, reason) that disagree."""
    import fuzz_cases as F
    import bayesloop_b200 as bl
    from bayesloop_b200 import engine

    def run(eng, build):
        engine.set_default_engine(eng)
        sink = io.StringIO()
        try:
            with contextlib.redirect_stdout(sink), contextlib.redirect_stderr(sink), np.errstate(all='ignore'):
                return ('ok', F.extract(build(bl)))
        except Exception as e:  # noqa: BLE001 -- configuration errors must be the same on both sides
            return ('exc', type(e).__name__)
        finally:
            engine.set_default_engine(None)

    bad = []
    for seed in range(seed0, seed0 + n):
        build, label = F.draw_case(seed)
        a, b = run(first, build), run(second, build)
        if a[0] != b[0]:
            bad.append((seed, label, 'outcome %s vs %s' % (a, b[:1])))
        elif a[0] == 'exc':
            if a[1] != b[1]:
                bad.append((seed, label, 'exceptions %s vs %s' % (a[1], b[1])))
        else:
            for x, y in zip(a[1], b[1]):
                scale = max(1.0, float(np.nanmax(np.abs(y[np.isfinite(y)]), initial=0.0)))
                if x.shape != y.shape or not np.allclose(x, y, rtol=1e-6, atol=1e-11 * scale, equal_nan=True):
                    bad.append((seed, label, 'values differ'))
                    break
    return bad


def test_harness_agrees_with_itself(oracle_engine):
    """The comparison harness on CPU: oracle against oracle over a few random cases (also keeps the generator alive)."""
    assert run_cases(oracle_engine, oracle_engine, 12) == []


@pytest.mark.gpu
def test_random_programs_cuda_vs_oracle(cuda_engine, oracle_engine):
    n = int(os.environ.get('BLG_TEST_FUZZ', '300'))
    assert run_cases(cuda_engine, oracle_engine, n) == []


@pytest.mark.gpu
def test_random_programs_on_the_stream_kernels(cuda_engine, oracle_engine):
    """The same generator with the global-memory stream kernels forced (plan option force_stream)."""
    with cuda_engine.options(force_stream=1):
        assert run_cases(cuda_engine, oracle_engine, 60, seed0=1000) == []
