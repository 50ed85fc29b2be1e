"""The reference's own known-answer tests, re-run through the product's public API and ACCESSORS.

Each test below restates one test of the reference's suite (file:line cited; `/root/reference/tests/...`) with the
numbers that suite hard-codes -- posterior slices through `getParameterDistributions`, means, evidences,
hyper-parameter / joint / duration distributions, transition-model probabilities, optimised hyper-parameters --
at the reference's own tolerances or tighter.  Every test runs twice: on CPU with the host logic driven through the
C ABI into the oracle (`-m "not gpu"`), and on the B200 through libblgrid.so (`-m gpu`).
"""
import contextlib
import io

import numpy as np
import pytest

D5 = np.array([1, 2, 3, 4, 5])


@pytest.fixture(params=['oracle', pytest.param('cuda', marks=pytest.mark.gpu)])
def bl(request):
    from bayesloop_b200 import engine
    eng = request.getfixturevalue('oracle_engine' if request.param == 'oracle' else 'cuda_engine')
    engine.set_default_engine(eng)
    import bayesloop_b200
    with contextlib.redirect_stdout(io.StringIO()), np.errstate(all='ignore'):
        yield bayesloop_b200
    engine.set_default_engine(None)


def gauss20(bl, prior=lambda m, s: 1 / s ** 3):
    return bl.om.Gaussian('mean', bl.cint(0, 6, 20), 'sigma', bl.oint(0, 2, 20), prior=prior)


def check(S, name, column, dist, means, logE, rtol=1e-2, decimal=2):
    np.testing.assert_allclose(S.getParameterDistributions(name, density=False)[1][:, column], dist, rtol=rtol)
    np.testing.assert_allclose(S.getParameterMeanValues(name), means, rtol=rtol)
    np.testing.assert_almost_equal(S.logEvidence, logE, decimal=decimal)


# ---------------------------------------------------------------------------------------- tests/test_study.py
def test_study_one_parameter_static(bl):  # test_study.py:10-30
    S = bl.Study()
    S.loadData(D5)
    S.setOM(bl.om.Poisson('rate'))
    S.setTM(bl.tm.Static())
    S.fit()
    check(S, 'rate', 250, [0.00034] * 5, [3.09761] * 5, -10.4463425036, rtol=1e-2)


def test_study_one_parameter_grw(bl):  # test_study.py:32-52
    S = bl.Study()
    S.loadData(D5)
    S.setOM(bl.om.Poisson('rate'))
    S.setTM(bl.tm.GaussianRandomWalk('sigma', 0.1, target='rate'))
    S.fit()
    check(S, 'rate', 250, [0.000417, 0.000386, 0.000356, 0.000336, 0.000332],
          [3.073534, 3.08179, 3.093091, 3.104016, 3.111173], -10.4337420351)


def test_study_one_parameter_optimize(bl):  # test_study.py:145-176
    import sympy.stats as stats
    S = bl.Study()
    S.loadData(D5)
    S.setOM(bl.om.Poisson('rate', bl.oint(0, 6, 1000), prior=stats.Exponential('expon', 1.)))
    S.setTM(bl.tm.CombinedTransitionModel(bl.tm.GaussianRandomWalk('sigma', 2.1, target='rate'),
                                          bl.tm.RegimeSwitch('log10pMin', -3)))
    S.optimize()
    check(S, 'rate', 250, [1.820641e-03, 2.083830e-03, 7.730833e-04, 1.977125e-04, 9.441302e-05],
          [1.015955, 2.291846, 3.36402, 4.113622, 4.390356], -9.47362827569)
    np.testing.assert_almost_equal(S.getHyperParameterValue('sigma'), 2.11216289063, decimal=2)
    np.testing.assert_almost_equal(S.getHyperParameterValue('log10pMin'), -3.0, decimal=3)


def test_study_two_parameter_optimize(bl):  # test_study.py:317-347
    S = bl.Study()
    S.loadData(D5)
    S.setOM(gauss20(bl))
    S.setTM(bl.tm.CombinedTransitionModel(bl.tm.GaussianRandomWalk('sigma', 1.07, target='mean'),
                                          bl.tm.RegimeSwitch('log10pMin', -3.90)))
    S.optimize()
    check(S, 'mean', 5, [9.903855e-03, 1.887901e-02, 8.257234e-05, 5.142727e-06, 2.950377e-06],
          [0.979099, 1.951689, 3.000075, 4.048376, 5.020886], -8.010466752050611)
    np.testing.assert_almost_equal(S.getHyperParameterValue('sigma'), 1.065854087589326, decimal=2)
    np.testing.assert_almost_equal(S.getHyperParameterValue('log10pMin'), -4.039735868499399, decimal=2)


def test_study_parameter_distribution_at_time_and_average(bl):
    """getParameterDistribution(t | 'avg') is a slice / the time average of getParameterDistributions
    (core.py:864-928)."""
    S = bl.Study()
    S.loadData(D5, timestamps=[10, 11, 12, 13, 14])
    S.setOM(gauss20(bl))
    S.setTM(bl.tm.GaussianRandomWalk('sigma', 0.2, target='mean'))
    S.fit()
    x, all_ = S.getPDs('sigma')
    np.testing.assert_array_equal(S.getPD(12, 'sigma')[1], all_[2])
    np.testing.assert_allclose(S.getPD('avg', 'sigma')[1], all_.mean(axis=0), rtol=1e-12)
    np.testing.assert_allclose(S.getPD(12, 'sigma', density=False)[1].sum(), 1.0, rtol=1e-10)
    with pytest.raises(bl.exceptions.PostProcessingError):
        S.getPD(3, 'sigma')
    with pytest.raises(bl.exceptions.PostProcessingError):
        S.getPDs('nope')


# ----------------------------------------------------------------------------------- tests/test_hyperstudy.py
def test_hyperstudy_two_hyper_parameters(bl):  # test_hyperstudy.py:60-103
    S = bl.HyperStudy()
    S.loadData(D5)
    S.setOM(gauss20(bl))
    S.setTM(bl.tm.CombinedTransitionModel(bl.tm.GaussianRandomWalk('sigma', bl.cint(0, 0.2, 2), target='mean'),
                                          bl.tm.RegimeSwitch('log10pMin', [-3, -1])))
    S.fit()
    check(S, 'mean', 5, [0.005589, 0.112966, 0.04335, 0.00976, 0.002909],
          [0.963756, 2.105838, 2.837739, 3.734359, 4.595412], -10.7601875492, rtol=1e-4, decimal=5)
    x, p = S.getHyperParameterDistribution('sigma')
    np.testing.assert_allclose(np.array([x, p]), [[0., 0.2], [0.48943645, 0.51056355]], rtol=1e-5)
    x, y, p = S.getJointHyperParameterDistribution(['log10pMin', 'sigma'])
    np.testing.assert_allclose(np.array([x, y]), [[-3., -1.], [0., 0.2]], rtol=1e-5)
    np.testing.assert_allclose(p, [[0.00701834, 0.0075608], [0.48241812, 0.50300274]], rtol=1e-5)


# ------------------------------------------------------------------------------ tests/test_changepointstudy.py
def _serial(bl, cp_prior=None, grw_values=None, grw_prior=None, bp_prior=None):
    return bl.tm.SerialTransitionModel(
        bl.tm.Static(),
        bl.tm.ChangePoint('ChangePoint', [0, 1], prior=cp_prior),
        bl.tm.CombinedTransitionModel(
            bl.tm.GaussianRandomWalk('sigma', grw_values, target='mean', prior=grw_prior),
            bl.tm.RegimeSwitch('log10pMin', [-3, -1])),
        bl.tm.BreakPoint('BreakPoint', 'all', prior=bp_prior),
        bl.tm.Static())


def test_changepointstudy_changepoint_breakpoint(bl):  # test_changepointstudy.py:10-54
    S = bl.ChangepointStudy()
    S.loadData(D5)
    S.setOM(gauss20(bl))
    S.setTM(_serial(bl, grw_values=bl.cint(0, 0.2, 2)))
    S.fit()
    check(S, 'mean', 5, [0.012437, 0.030168, 0.01761, 0.001731, 0.001731],
          [0.968022, 1.956517, 3.476958, 4.161028, 4.161028], -15.072007461556161, decimal=5)
    x, p = S.getHyperParameterDistribution('sigma')
    np.testing.assert_allclose(np.array([x, p]), [[0., 0.2], [0.4963324, 0.5036676]], rtol=1e-2)
    d, p = S.getDurationDistribution(['ChangePoint', 'BreakPoint'])
    np.testing.assert_allclose(np.array([d, p]), [[1., 2., 3.], [0.01039273, 0.49395867, 0.49564861]], rtol=1e-2)


def test_changepointstudy_hyperpriors(bl):  # test_changepointstudy.py:56-100
    import sympy.stats as stats
    S = bl.ChangepointStudy()
    S.loadData(D5)
    S.setOM(gauss20(bl))
    S.setTM(_serial(bl, cp_prior=np.array([0.3, 0.7]), grw_values=bl.oint(0, 0.2, 2), grw_prior=lambda s: 1. / s,
                    bp_prior=stats.Normal('Normal', 3., 1.)))
    S.fit()
    check(S, 'mean', 5, [0.033729, 0.050869, 0.020636, 0.001647, 0.001647],
          [0.98944, 1.927195, 3.349921, 4.213695, 4.213695], -15.709534690217343, decimal=5)
    x, p = S.getHyperParameterDistribution('sigma')
    np.testing.assert_allclose(np.array([x, p]), [[0.06666667, 0.13333333], [0.66515107, 0.33484893]], rtol=1e-2)
    d, p = S.getDurationDistribution(['ChangePoint', 'BreakPoint'])
    np.testing.assert_allclose(np.array([d, p]), [[1., 2., 3.], [0.00373717, 0.40402616, 0.59223667]], rtol=1e-2)


# ---------------------------------------------------------------------------------- tests/test_onlinestudy.py
def test_onlinestudy_static(bl):  # test_onlinestudy.py:10-32
    S = bl.OnlineStudy(storeHistory=True)
    S.setOM(gauss20(bl))
    S.setTM(bl.tm.Static())
    for d in D5:
        S.step(d)
    check(S, 'mean', 5, [0.0053811, 0.38690331, 0.16329865, 0.04887604, 0.01334921],
          [0.96310103, 1.5065597, 2.00218465, 2.500366, 3.], -16.1946904707, rtol=1e-5, decimal=5)


def _online_two_models(bl, storeHistory=True):  # test_onlinestudy.py:34-55
    import sympy.stats as stats
    S = bl.OnlineStudy(storeHistory=storeHistory)
    S.setOM(gauss20(bl, prior=lambda m, s: 1. / s))
    S.addTransitionModel('T1', bl.tm.CombinedTransitionModel(
        bl.tm.GaussianRandomWalk('s1', [0.25, 0.5], target='mean', prior=stats.Exponential('e', 0.5)),
        bl.tm.GaussianRandomWalk('s2', bl.cint(0, 0.2, 2), target='sigma', prior=np.array([0.2, 0.8]))))
    S.addTransitionModel('T2', bl.tm.Independent())
    S.setTransitionModelPrior([0.9, 0.1])
    for d in D5:
        S.step(d)
    return S


def test_onlinestudy_two_transition_models(bl):  # test_onlinestudy.py:57-84
    S = _online_two_models(bl)
    names, p = S.getCurrentTransitionModelDistribution(local=False)
    assert list(names) == ['T1', 'T2']
    np.testing.assert_allclose(p, [0.49402616, 0.50597384], rtol=1e-5)
    np.testing.assert_allclose(S.getCurrentTransitionModelDistribution(local=True)[1], [0.81739495, 0.18260505],
                               rtol=1e-5)
    np.testing.assert_allclose(S.getCurrentHyperParameterDistribution('s2')[1], [0.19047162, 0.80952838], rtol=1e-5)
    check(S, 'mean', 5, [0.05825921, 0.20129444, 0.07273516, 0.02125759, 0.0039255],
          [1.0771838, 1.71494272, 2.45992376, 3.34160617, 4.39337253], -9.46900822686, rtol=1e-5, decimal=5)


def test_onlinestudy_history_accessors_are_consistent(bl):
    """The time-indexed accessors of core.py:2231-2833 against each other and against the 'current' ones: the last
    entry of every history equals the value read from the device state after the last step."""
    S = _online_two_models(bl)
    last = S.formattedTimestamps[-1]

    x, now = S.getCurrentParameterDistribution('mean')
    np.testing.assert_allclose(S.getParameterDistribution(last, 'mean')[1], now, rtol=1e-12)
    np.testing.assert_allclose(S.getParameterDistributions('mean')[1][-1], now, rtol=1e-12)
    np.testing.assert_allclose(S.getPD('avg', 'mean')[1], S.getPDs('mean')[1].mean(axis=0), rtol=1e-12)
    np.testing.assert_allclose(S.getCPD('mean', density=False)[1].sum(), 1.0, rtol=1e-10)

    np.testing.assert_allclose(S.getCurrentParameterMeanValue('mean'), S.getParameterMeanValues('mean')[-1], rtol=1e-12)
    np.testing.assert_allclose(S.getParameterMeanValue(last, 'sigma'), S.getParameterMeanValues('sigma')[-1],
                               rtol=1e-12)

    names, seq = S.getTransitionModelDistributions()
    assert seq.shape == (5, 2)
    np.testing.assert_allclose(seq[-1], S.getCTMD()[1], rtol=1e-12)
    np.testing.assert_allclose(seq.sum(axis=1), 1.0, rtol=1e-12)
    np.testing.assert_allclose(S.getTMPs('T2'), seq[:, 1], rtol=0)
    np.testing.assert_allclose(S.getTMPs('T1', local=True)[-1], S.getCTMP('T1', local=True), rtol=1e-12)

    values, cur = S.getCHPD('s1')
    np.testing.assert_allclose(values, [0.25, 0.5])
    np.testing.assert_allclose(cur.sum(), 1.0, rtol=1e-10)
    vals, hist = S.getHPDs('s1')
    assert hist.shape == (5, 2)
    np.testing.assert_allclose(hist[-1], cur, rtol=1e-10)
    np.testing.assert_allclose(S.getHPD(last, 's1')[1] * np.prod(S.hyperGridConstants[0]), cur, rtol=1e-12)
    means = S.getHyperParameterMeanValues('s1')
    np.testing.assert_allclose(means[-1], np.sum(values * cur), rtol=1e-10)
    np.testing.assert_allclose(S.getHyperParameterMeanValue(last, 's1'), means[-1], rtol=1e-12)
    np.testing.assert_allclose(S.getCurrentHyperParameterMeanValue('s1'), means[-1], rtol=1e-12)
    assert np.all((means >= 0.25) & (means <= 0.5))

    with pytest.raises(bl.exceptions.PostProcessingError):
        S.getHPD(last, 'unknown')
    with pytest.raises(bl.exceptions.PostProcessingError):
        S.getPD(99, 'mean')
    with pytest.raises(NotImplementedError):
        S.getJointHyperParameterDistribution(['s1', 's2'])
    with pytest.raises(NotImplementedError):
        S.fit()


def test_onlinestudy_without_history_refuses_past_queries(bl):
    S = _online_two_models(bl, storeHistory=False)
    np.testing.assert_allclose(S.getCTMD()[1], [0.49402616, 0.50597384], rtol=1e-5)
    np.testing.assert_allclose(S.getCurrentParameterMeanValue('mean'), 4.39337253, rtol=1e-5)
    for call in (lambda: S.getPDs('mean'), lambda: S.getPD(4, 'mean'), lambda: S.getParameterMeanValues('mean'),
                 lambda: S.getParameterMeanValue(4, 'mean'), lambda: S.getTransitionModelDistributions(),
                 lambda: S.getTMPs('T1'), lambda: S.getHPDs('s1'), lambda: S.getHPD(4, 's1'),
                 lambda: S.getHyperParameterMeanValues('s1'), lambda: S.getHyperParameterMeanValue(4, 's1')):
        with pytest.raises(bl.exceptions.PostProcessingError):
            call()


# -------------------------------------------------------------------------------------- tests/test_fileio.py
def test_save_load(bl, tmp_path):  # test_fileio.py:8-17 (the prior is a lambda: the study must still be storable)
    S = bl.HyperStudy()
    S.loadData(D5)
    S.setOM(gauss20(bl))
    S.setTM(bl.tm.Static())
    S.fit()
    bl.save(str(tmp_path / 'study.bl'), S)
    R = bl.load(str(tmp_path / 'study.bl'))
    assert type(R) is type(S)
    assert R.logEvidence == S.logEvidence
    np.testing.assert_array_equal(R.posteriorSequence, S.posteriorSequence)
    np.testing.assert_array_equal(R.getParameterMeanValues('mean'), S.getParameterMeanValues('mean'))
    R.fit()  # a loaded study is a working study
    assert R.logEvidence == S.logEvidence
