"""The oracles against third-party ground truth and against each other (CPU only).

  * restated reflect filter (oracle/np_oracle.py) vs scipy.ndimage.gaussian_filter1d, incl. R >= n
  * C oracle (linear-space averaging, own filter) vs NumPy port (reference-style log-space logaddexp averaging,
    SciPy filter) on seeded synthetic sweeps
The goldens of the real reference are checked in test_host_logic.py (every fixture goes through the C oracle)."""
import numpy as np
import pytest
from scipy.ndimage import gaussian_filter1d

import helpers  # noqa: F401  (adds oracle/ to sys.path)
import np_oracle


@pytest.mark.parametrize('n,sigma', [(50, 0.3), (50, 2.0), (50, 20.0), (50, 100.0), (1000, 83.4), (7, 3.0), (5, 0.13)])
def test_restated_filter_matches_scipy(n, sigma):
    x = np.random.default_rng(n).random(n)
    np.testing.assert_allclose(np_oracle.gaussian_filter1d_reflect(x, sigma, 0), gaussian_filter1d(x, sigma),
                               rtol=1e-13, atol=1e-16)


def test_restated_filter_2d_axes():
    x = np.random.default_rng(0).random((12, 9))
    for axis, sigma in ((0, 5.0), (1, 0.8)):
        np.testing.assert_allclose(np_oracle.gaussian_filter1d_reflect(x, sigma, axis),
                                   gaussian_filter1d(x, sigma, axis=axis), rtol=1e-13)


def _poisson_study(bl, engine, B=6, T=80, G=120):
    rng = np.random.default_rng(21)
    S = bl.HyperStudy(silent=True, engine=engine)
    S.loadData(rng.poisson(3 + 2 * np.sin(np.arange(T) / 9.)).astype(float), silent=True)
    S.set(bl.om.Poisson('rate', bl.oint(0, 12, G)),
          bl.tm.CombinedTransitionModel(bl.tm.GaussianRandomWalk('sigma', bl.cint(0, 0.4, B), target='rate'),
                                        bl.tm.RegimeSwitch('p', -6)), silent=True)
    return S


def _gauss_study(bl, engine):
    rng = np.random.default_rng(22)
    S = bl.HyperStudy(silent=True, engine=engine)
    S.loadData(rng.normal(0.3, 1.1, 40), silent=True)
    S.set(bl.om.Gaussian('mean', bl.cint(-3, 3, 24), 'std', bl.oint(0, 3, 20)),
          bl.tm.SerialTransitionModel(
              bl.tm.CombinedTransitionModel(bl.tm.GaussianRandomWalk('a', [0.1, 0.3], target='mean'),
                                            bl.tm.GaussianRandomWalk('b', [0.05, 0.2], target='std')),
              bl.tm.ChangePoint('tc', [10, 20, 30]),
              bl.tm.Static()), silent=True)
    return S


@pytest.mark.parametrize('make', [_poisson_study, _gauss_study])
def test_c_oracle_matches_numpy_port(make, oracle_engine):
    import bayesloop_b200 as bl
    S = make(bl, oracle_engine)
    got = helpers.abi_sweep(oracle_engine, S)
    ops, hp, _ = helpers.lowered(S)
    pb = helpers.np_problem(S)
    ref = np_oracle.hyper_fit(pb, ops, hp, use_scipy=True)
    np.testing.assert_allclose(got['logE'], ref['logEvidenceList'], rtol=1e-11)
    avg, means = np_oracle.finish_average(pb, ref['logAverage'])
    np.testing.assert_allclose(got['avg'].reshape(avg.shape), avg, rtol=1e-8, atol=1e-15)
    np.testing.assert_allclose(got['means'], means, rtol=1e-10)
    np.testing.assert_allclose(got['local'], ref['localEvidenceList'], rtol=1e-9)
