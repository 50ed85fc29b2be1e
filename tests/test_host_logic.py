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


@pytest.mark.parametrize('name,wave', [('syn_hyper_poisson_sweep', 5), ('syn_hyper_poisson_sweep', 1),
                                       ('syn_hyper_dead_combo', 1), ('syn_cps_gauss_2d', 7), ('ref_cps_coal_all', 10)])
def test_sweep_in_several_waves_matches_the_golden(name, wave, use_oracle, monkeypatch):
    """The device-side averaging across waves (blg_wave_weights: weights relative to a running reference log-weight,
    re-base of the running sum when a later wave brings a larger evidence) reproduces the reference's
    np.logaddexp accumulation (core.py:1358-1366) whatever the wave size."""
    import bayesloop_b200 as bl
    from bayesloop_b200 import core
    plain = core.HyperStudy.__init__

    def capped(self, *a, **kw):
        plain(self, *a, **kw)
        self.maxWave = wave

    monkeypatch.setattr(core.HyperStudy, '__init__', capped)
    S, got = parity.run_case(name, bl)
    assert S.sweepStats['waves'] >= 2
    parity.compare(name, got, load_golden(name), rtol=2e-9, atol_post=1e-13)


def _online_study(bl):
    S = bl.OnlineStudy(storeHistory=False, silent=True)
    S.setOM(bl.om.ScaledAR1('rho', bl.oint(-1, 1, 24), 'sigma', bl.oint(0, 3, 28)), silent=True)
    S.add('normal', bl.tm.CombinedTransitionModel(bl.tm.GaussianRandomWalk('s1', bl.cint(0, 0.1, 3), target='rho'),
                                                  bl.tm.GaussianRandomWalk('s2', bl.cint(0, 0.2, 2), target='sigma')))
    S.add('chaotic', bl.tm.RegimeSwitch('p', bl.cint(-8, -3, 3)))
    S.add('indep', bl.tm.Independent())
    return S


def test_online_study_checkpoint_resume(use_oracle):
    """an OnlineStudy pickled between two steps (device state -> host) and resumed continues
    exactly like the uninterrupted one (reference: bl.save / bl.load, fileIO.py:10-37)."""
    import contextlib
    import io
    import pickle
    import bayesloop_b200 as bl
    rng = np.random.default_rng(21)
    x = np.zeros(40)
    for i in range(1, len(x)):
        x[i] = 0.5 * x[i - 1] + rng.normal()
    with contextlib.redirect_stdout(io.StringIO()):
        A, B = _online_study(bl), _online_study(bl)
        for d in x:
            A.step(d)
        for d in x[:17]:
            B.step(d)
        blob = pickle.dumps(B)
        del B
        C = pickle.loads(blob)
        np.testing.assert_allclose(C.marginalizedPosterior, pickle.loads(blob).marginalizedPosterior)  # usable at once
        for d in x[17:]:
            C.step(d)
    assert C.logEvidence == A.logEvidence
    np.testing.assert_array_equal(C.marginalizedPosterior, A.marginalizedPosterior)
    np.testing.assert_array_equal(C.transitionModelDistribution, A.transitionModelDistribution)
    for a, c in zip(A.parameterPosterior, C.parameterPosterior):
        np.testing.assert_array_equal(a, c)


def test_edge_cases_behave_like_the_reference(use_oracle):
    """Probed against the unmodified reference in the build container (it cannot travel): a series shorter than one
    segment gives empty sequences and logE = log(prod(latticeConstant)) (the loops of core.py:372-470 never run,
    :417 still does); a negative Poisson count raises ValueError (math.factorial, observationModels.py:502)."""
    import contextlib
    import io
    import bayesloop_b200 as bl
    with contextlib.redirect_stdout(io.StringIO()):
        S = bl.Study(silent=True)
        S.loadData(np.array([0.3]), silent=True)
        S.set(bl.om.ScaledAR1('rho', bl.oint(-1, 1, 20), 'sig', bl.oint(0, 2, 20)), bl.tm.Static(), silent=True)
        S.fit(silent=True)
        assert abs(S.logEvidence - (-4.702750514326955)) < 1e-12
        assert S.posteriorSequence.shape == (0, 20, 20) and S.posteriorMeanValues.shape == (2, 0)
        assert len(S.localEvidence) == 0

        P = bl.Study(silent=True)
        P.loadData(np.array([1., -2., 3.]), silent=True)
        P.set(bl.om.Poisson('r', bl.oint(0, 6, 50)), bl.tm.Static(), silent=True)
        with pytest.raises(ValueError):
            P.fit(silent=True)


def test_marginals_from_the_resident_sequence_equal_host_reductions(use_oracle):
    """while the fitted [T x G] sequence is still in engine memory, marginal distributions and the
    time average (core.py:886, :915, :979-980) are reduced there (blg_marginal / blg_time_average); once the attribute
    has been read they are NumPy reductions of the downloaded array.  Both paths must agree."""
    import bayesloop_b200 as bl
    S, _ = None, None
    rng = np.random.default_rng(3)
    S = bl.HyperStudy(silent=True)
    S.loadData(rng.normal(0.3, 1.0, 25), silent=True)
    S.set(bl.om.Gaussian('mean', bl.cint(-2, 2, 18), 'std', bl.oint(0, 3, 14)),
          bl.tm.GaussianRandomWalk('s', bl.cint(0, 0.3, 3), target='mean'), silent=True)
    S.fit(silent=True)
    assert S.__dict__['_postDev'] is not None  # nothing has been downloaded yet
    dev = {k: (S.getParameterDistributions(k)[1], S.getParameterDistribution(7, k)[1], S.getParameterDistribution('avg', k)[1])
           for k in ('mean', 'std')}
    assert S.__dict__['_postDev'] is not None
    seq = S.averagePosteriorSequence  # download
    assert S.__dict__['_postDev'] is None and seq.shape == (25, 18, 14)
    for k in ('mean', 'std'):
        host = (S.getParameterDistributions(k)[1], S.getParameterDistribution(7, k)[1], S.getParameterDistribution('avg', k)[1])
        for a, b in zip(dev[k], host):
            np.testing.assert_allclose(a, b, rtol=1e-13)
    assert dev['mean'][0].shape == (25, 18) and dev['std'][0].shape == (25, 14)
