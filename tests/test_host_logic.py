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


def test_array_prior_is_never_written_to_by_the_filter(use_oracle):
    """The reference's forward loop multiplies the likelihood INTO the array _computePrior returned (core.py:362,
    :385); for an ndarray prior that is the user's own array, so a ChangePoint / Independent reset
    (transitionModels.py:303-304, :352-353) restores whatever the filter wrote into it, and every later combination of
    a sweep starts from it (found by oracle/differential_probe.py).  The product normalises the array in place like
    core.py:212-213 and then leaves it alone: an array prior gives exactly what the same values give as a callable
    (core.py:224-235), which is also what the reference computes for the callable."""
    import bayesloop_b200 as bl
    rng = np.random.default_rng(5)
    rng.normal(0.3, 0.8, 15)  # the probe script draws its Gaussian series first
    counts = rng.poisson(3, 15)
    values = np.concatenate([np.zeros(10), np.ones(50)])

    def fit(prior, cls, tm):
        S = cls(silent=True)
        S.loadData(counts, silent=True)
        S.set(bl.om.Poisson('r', bl.oint(0, 8, 60), prior=prior), tm, silent=True)
        S.fit(silent=True)
        return S

    models = [(bl.Study, lambda: bl.tm.ChangePoint('t', 6)), (bl.Study, lambda: bl.tm.Independent()),
              (bl.HyperStudy, lambda: bl.tm.GaussianRandomWalk('s', [0.1, 0.3], target='r')),
              (bl.ChangepointStudy, lambda: bl.tm.CombinedTransitionModel(
                  bl.tm.ChangePoint('t', 'all'), bl.tm.GaussianRandomWalk('s', [0.1, 0.3], target='r')))]
    for cls, tm in models:
        array = values.copy()
        A = fit(array, cls, tm())
        C = fit(lambda r: values.copy(), cls, tm())
        lattice = 8. / 61.
        np.testing.assert_allclose(array, values / values.sum() / lattice, rtol=1e-14)  # normalised in place, nothing else
        assert abs(A.logEvidence - C.logEvidence) <= 1e-12 * abs(C.logEvidence)
        np.testing.assert_allclose(A.posteriorSequence, C.posteriorSequence, rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(A.localEvidence, C.localEvidence, rtol=1e-12)
    # the reference, given the callable, on the first model of the list (reference value computed in the build container
    # by oracle/differential_probe.py: 'prior array with a zero region and a change-point')
    assert abs(fit(values.copy(), bl.Study, bl.tm.ChangePoint('t', 6)).logEvidence - (-31.911321254878697)) < 1e-10


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


# ---------------------------------------------------------------------------------------------------------------------
# change-point prefix sharing

def _cp_study(bl, engine=None, cps=np.arange(3, 30, 3), hs=(3, 2), T=34, share=True, wave=None):
    rng = np.random.default_rng(8)
    x = np.concatenate([rng.normal(-1, 0.7, T // 2), rng.normal(1.2, 0.9, T - T // 2)])
    S = bl.ChangepointStudy(silent=True, engine=engine)
    S.loadData(x, silent=True)
    S.set(bl.om.Gaussian('mean', bl.cint(-3, 3, 24), 'std', bl.oint(0, 3, 20)),
          bl.tm.CombinedTransitionModel(bl.tm.ChangePoint('tChange', cps),
                                        bl.tm.GaussianRandomWalk('s_mean', bl.cint(0, 0.2, hs[0]), target='mean'),
                                        bl.tm.GaussianRandomWalk('s_std', bl.cint(0.02, 0.1, hs[1]), target='std')),
          silent=True)
    S.shareChangepoints = share
    S.maxWave = wave
    S.fit(silent=True)
    return S


@pytest.mark.parametrize('wave', [None, 12, 6])
def test_changepoint_prefix_sharing_equals_the_plain_sweep(wave, use_oracle):
    """The shared schedule (one change-point-free run per group, forward passes over the steps after the change-point,
    backward passes over the rows before it) against the plain sweep of the same study: identical results, about half
    the executed cell updates.  The reference's history erasure at tChange (transitionModels.py:300-312) is what makes
    the two equal; combos are enumerated as in core.py:1823-1842."""
    import bayesloop_b200 as bl
    plain = _cp_study(bl, share=False)
    shared = _cp_study(bl, share=True, wave=wave)
    assert not plain.sweepStats['shared'] and shared.sweepStats['shared']
    assert shared.sweepStats['nominal_updates'] == plain.sweepStats['executed_updates']
    ratio = shared.sweepStats['executed_updates'] / shared.sweepStats['nominal_updates']
    assert 0.5 < ratio < 0.68, ratio  # (T - c) + (c + 2) per combination + one shared run per group, out of 2 T
    if wave:
        assert shared.sweepStats['waves'] >= 2
    np.testing.assert_allclose(shared.logEvidenceList, plain.logEvidenceList, rtol=1e-12)
    assert abs(shared.logEvidence - plain.logEvidence) <= 1e-12 * abs(plain.logEvidence)
    np.testing.assert_allclose(shared.hyperParameterDistribution, plain.hyperParameterDistribution, rtol=1e-10)
    np.testing.assert_allclose(shared.posteriorMeanValues, plain.posteriorMeanValues, rtol=1e-10)
    np.testing.assert_allclose(shared.localEvidence, plain.localEvidence, rtol=1e-9)
    a, b = shared.posteriorSequence, plain.posteriorSequence
    top = b.reshape(len(b), -1).max(axis=1).reshape(-1, 1, 1)
    assert np.all(np.abs(a - b) <= 1e-9 * np.abs(b) + 1e-14 * top)


def test_prefix_sharing_only_where_the_history_is_erased(use_oracle):
    """Sweeps whose combinations differ in more than the step of ONE reset keep the plain schedule: break-points switch
    sub-models without erasing the history, Serial models with two points, Independent resets at every step."""
    import bayesloop_b200 as bl
    from bayesloop_b200.core import _share_structure

    def structure(T_model, data=np.arange(12.) % 5, cls=None):
        S = (cls or bl.ChangepointStudy)(silent=True)
        S.loadData(data, silent=True)
        S.set(bl.om.Poisson('rate', bl.oint(0, 8, 40)), T_model, silent=True)
        S._formatData()
        S._createHyperGrid(silent=True)
        ctx = S._lower(np.asarray(S.hyperGridValues, dtype=float).reshape(len(S.hyperGridValues), -1), S.formattedTimestamps)
        return _share_structure(ctx.ops, len(S.formattedData))

    grw = bl.tm.GaussianRandomWalk
    got = structure(bl.tm.CombinedTransitionModel(bl.tm.ChangePoint('t', 'all'), grw('s', [0.1, 0.2, 0.3], target='rate')))
    assert got is not None and (got['nG'], got['nC']) == (3, 11)
    got = structure(bl.tm.CombinedTransitionModel(grw('s', [0.1, 0.2], target='rate'), bl.tm.ChangePoint('t', [2, 5, 7])))
    assert got is not None and (got['nG'], got['nC']) == (2, 3)  # reset listed AFTER the random walk
    assert structure(bl.tm.SerialTransitionModel(bl.tm.Static(), bl.tm.BreakPoint('t', 'all'), grw('s', 0.2, target='rate'))) is None
    assert structure(bl.tm.CombinedTransitionModel(bl.tm.Independent(), grw('s', [0.1, 0.2, 0.3, 0.4], target='rate')),
                     cls=bl.HyperStudy) is None
    assert structure(grw('s', [0.1, 0.2, 0.3, 0.4], target='rate'), cls=bl.HyperStudy) is None
    # a change-point next to a Serial model whose segments are active on RANGES of steps: the passes of the shared
    # schedule run on windows of the sequence (step indices start at 0 again), so such a sweep keeps the plain schedule
    ranged = bl.tm.CombinedTransitionModel(
        bl.tm.ChangePoint('t', [2, 4]),
        bl.tm.SerialTransitionModel(bl.tm.NotEqual('q', -4.), bl.tm.BreakPoint('b', 3), bl.tm.RegimeSwitch('p', [-7., -4.])))
    assert structure(ranged, cls=bl.HyperStudy) is None


def test_shared_sweep_with_a_dead_combination_falls_back_to_the_plain_schedule(use_oracle):
    """A NotEqual model turns a flat distribution into 0/0: with a missing value right behind the change-point the
    combination dies (zero / NaN norm, core.py:388-400, :440-452).  The reference keeps the local evidences of the
    forward pass in the rows the backward pass never reached (core.py:1356); the plain sweep reproduces that, so a
    shared sweep in which a combination died is run again with the plain schedule."""
    import bayesloop_b200 as bl
    data = np.array([0.84, -0.15, 0.26, 0.53, 0.69, -0.89, 0.47, -0.02, -0.60, -0.68, -0.23, np.nan, -1.46, 0.06, 0.77, -0.24])

    def fit(share):
        S = bl.HyperStudy(silent=True)
        S.loadData(data, silent=True)
        S.set(bl.om.ScaledAR1('rho', bl.oint(-1, 1, 10), 'sig', bl.oint(0, 3, 10)),
              bl.tm.CombinedTransitionModel(bl.tm.NotEqual('q', [-3., -2., -1.]), bl.tm.ChangePoint('t', list(range(1, 15)))),
              silent=True)
        S.shareChangepoints = share
        S.fit(silent=True)
        return S

    shared, plain = fit(True), fit(False)
    dead = ~np.isfinite(np.asarray(plain.logEvidenceList))
    assert dead.any() and not dead.all()
    assert not shared.sweepStats['shared']
    np.testing.assert_array_equal(np.asarray(shared.logEvidenceList), np.asarray(plain.logEvidenceList))
    np.testing.assert_array_equal(np.asarray(shared.localEvidence), np.asarray(plain.localEvidence))
    np.testing.assert_allclose(shared.posteriorSequence, plain.posteriorSequence, rtol=1e-12)  # other summation order


def test_changepoint_sweep_next_to_a_serial_model_matches_the_reference(use_oracle):
    """Found by oracle/fuzz_lowering.py (seeds 22257, 22726, 25252 of 8 000): a HyperStudy over the time of a
    change-point combined with a Serial model (NotEqual before, RegimeSwitch after a FIXED break-point) was run with
    prefix sharing, whose windowed passes shifted the Serial model's segments.  Expected values: the unmodified
    reference in the build container (core.py:1349-1419 over transitionModels.py:289-314, :756-818)."""
    import bayesloop_b200 as bl
    S = bl.HyperStudy(silent=True)
    S.loadData(np.array([2., 4, 2, 3, 5, 1, 1, 2]), silent=True)
    S.set(bl.om.Poisson('r', bl.oint(0, 8, 12)),
          bl.tm.CombinedTransitionModel(
              bl.tm.ChangePoint('t', [2, 4]),
              bl.tm.SerialTransitionModel(bl.tm.NotEqual('q4', -4.), bl.tm.BreakPoint('b6', 3),
                                          bl.tm.RegimeSwitch('p5', [-7., -4.]))), silent=True)
    S.fit(silent=True)
    assert not S.sweepStats['shared']
    np.testing.assert_allclose(S.logEvidenceList, [-15.92795686903254, -15.927738811661808, -14.397293915094128,
                                                   -14.397330652637821], rtol=1e-10)
    assert abs(S.logEvidence - (-14.894547565491354)) < 1e-9
    np.testing.assert_allclose(S.posteriorMeanValues[0], [
        1.826368071727187, 5.256682591927492, 1.864359256320257, 4.250236184533213, 4.249751692782431, 1.707450557414314,
        1.7074487137590055, 1.707457673995852], rtol=1e-9)


def test_sm_assignment_table_of_a_c2_like_sweep(oracle_engine):
    """Program._assignment: the per-SM chain lists of the 1-D kernels (blg_program.sm_assign).  Every combo exactly once,
    at most 4 per SM, filled slots form a prefix of every list (the interleaved kernel reads them in order), and the
    predicted cost (chains + DMMA groups, engine.py) is level to a few per cent for the sigma sweep of BASELINE configs[1]."""
    import numpy as np
    from bayesloop_b200.engine import Program
    B, sms = 512, 148
    sigma_n = np.linspace(0, 0.2, B) / (12.0 / 1001)  # cint(0, 0.2, 512) on oint(0, 12, 1000)
    radius = (4.0 * sigma_n + 0.5).astype(np.int32)
    ops = [dict(kind=1, axis=0, param=sigma_n, radius=radius, window=np.tile(np.array([0, 1 << 30, 0, 1 << 30], np.int32), (B, 1)))]
    prog = Program(oracle_engine, ops, B)
    table = prog._assignment(0, B, sms).cpu().numpy()
    assert table.shape == (sms, 4)
    used = table[table >= 0]
    assert sorted(used.tolist()) == list(range(B))
    for row in table:
        filled = np.flatnonzero(row >= 0)
        assert list(filled) == list(range(len(filled)))
    groups = (2 * radius + 15 + (radius & 1)) // 8
    cost = np.array([sum(1.0 + groups[b] for b in row if b >= 0) for row in table])
    assert cost.max() <= 1.06 * cost.mean()
    assert prog._assignment(0, B, sms // 4) is None  # more combos than 4 per SM: the hardware places the CTAs
