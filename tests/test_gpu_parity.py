"""GPU parity tests (run on the B200 box with -m gpu): the CUDA library, reached through the public API and the
C ABI, against (1) the golden vectors of the unmodified reference, (2) the CPU oracle on seeded mid-size inputs,
(3) size-independent properties at larger sizes.

Tolerances (BASELINE.json north_star: 1e-6 relative in fp64): log-evidences 1e-9 relative; posterior grids 1e-6
relative with an absolute floor of 1e-12 x the per-time-step maximum (cells far in the tails carry rounding noise of
different summation orders)."""
import numpy as np
import pytest

import cases
import helpers
import parity
from conftest import load_golden

pytestmark = pytest.mark.gpu


def _families():
    """Kernel family (blg_last_kernel) every golden case ran on when it last passed on a B200: a dispatch regression
    -- e.g. the headline sweep silently falling back from fwd/bwd_fast1d_ws to the generic resident kernels -- fails
    the case even though its numbers still agree.  tests/golden/kernel_families.json is written from a GPU run
    (BLG_RECORD_FAMILIES=<path> collects the observed names); a case missing from it is recorded, not asserted."""
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'kernel_families.json')
    if not os.path.exists(path):
        return {}
    with open(path) as f:
        return json.load(f)


@pytest.mark.parametrize('name', sorted(cases.CASES))
def test_case_matches_reference_golden(name, use_cuda):
    import json
    import os
    import bayesloop_b200 as bl
    before = use_cuda.launch_count()
    S, got = parity.run_case(name, bl)
    assert use_cuda.launch_count() > before, 'no CUDA kernel was launched'
    family = use_cuda.last_kernel()
    record = os.environ.get('BLG_RECORD_FAMILIES')
    if record:
        seen = json.load(open(record)) if os.path.exists(record) else {}
        seen[name] = family
        with open(record, 'w') as f:
            json.dump(seen, f, indent=0, sort_keys=True)
    want_family = _families().get(name)
    if want_family is not None:
        assert family == want_family, 'dispatch changed: {} ran on {}, expected {}'.format(name, family, want_family)
    parity.compare(name, got, load_golden(name), rtol=1e-6, atol_post=1e-12)
    want = float(load_golden(name)['logEvidence'])
    if np.isfinite(want):
        assert abs(float(got['logEvidence']) - want) <= 1e-9 * abs(want)


def _poisson(bl, engine, B, T, G, smax, seed=1, extra=None):
    rng = np.random.default_rng(seed)
    lam = 3 + 2 * np.sin(2 * np.pi * np.arange(T) / 200.)
    S = bl.HyperStudy(silent=True, engine=engine)
    S.loadData(rng.poisson(lam).astype(float), silent=True)
    T_ = bl.tm.GaussianRandomWalk('sigma', bl.cint(0, smax, B), target='rate')
    if extra is not None:
        T_ = bl.tm.CombinedTransitionModel(T_, extra(bl))
    S.set(bl.om.Poisson('rate', bl.oint(0, 12, G)), T_, silent=True)
    return S


def _gauss2d(bl, engine, n0, n1, T, hb, seed=2):
    rng = np.random.default_rng(seed)
    mu = np.cumsum(rng.normal(0, 0.03, T))
    S = bl.HyperStudy(silent=True, engine=engine)
    S.loadData(rng.normal(mu, 1.0), silent=True)
    S.set(bl.om.Gaussian('mean', bl.cint(-3, 3, n0), 'std', bl.oint(0, 3, n1)),
          bl.tm.CombinedTransitionModel(bl.tm.GaussianRandomWalk('s_mean', bl.cint(0, 0.3, hb), target='mean'),
                                        bl.tm.GaussianRandomWalk('s_std', bl.cint(0, 0.2, hb), target='std')),
          silent=True)
    return S


CONFIGS = {
    'poisson_c2_small': lambda bl, e: _poisson(bl, e, B=24, T=300, G=1000, smax=0.2),
    'poisson_wide_kernels': lambda bl, e: _poisson(bl, e, B=6, T=60, G=1000, smax=1.0),
    # more chains than SMs, plan option il = 1: the chains of an SM interleaved inside one CTA (fast1d_il.cuh), 3-4 / 1-2 per SM
    'poisson_c2_interleaved': lambda bl, e: _poisson(bl, e, B=500, T=70, G=1000, smax=0.2),
    'poisson_interleaved_small_grid': lambda bl, e: _poisson(bl, e, B=200, T=90, G=200, smax=0.5),
    # the other geometries of the DMMA kernels: 6 tiles per warp (4 compute warps), 8 compute warps x 4 tiles, 1 tile per warp
    'poisson_grid_1500': lambda bl, e: _poisson(bl, e, B=5, T=40, G=1500, smax=0.3),
    'poisson_grid_2000': lambda bl, e: _poisson(bl, e, B=5, T=40, G=2000, smax=0.4),
    'poisson_grid_250': lambda bl, e: _poisson(bl, e, B=7, T=60, G=250, smax=0.6),
    'poisson_regime': lambda bl, e: _poisson(bl, e, B=8, T=200, G=500, smax=0.1,
                                             extra=lambda bl: bl.tm.RegimeSwitch('p', -5)),
    'poisson_odd_grid': lambda bl, e: _poisson(bl, e, B=5, T=100, G=333, smax=0.3),
    'gauss_2d_64x48': lambda bl, e: _gauss2d(bl, e, 64, 48, T=80, hb=3),
    'gauss_2d_100x100': lambda bl, e: _gauss2d(bl, e, 100, 100, T=30, hb=2),
    'gauss_2d_200x200_stream': lambda bl, e: _gauss2d(bl, e, 200, 200, T=16, hb=2),
    'gauss_2d_256x96_stream': lambda bl, e: _gauss2d(bl, e, 256, 96, T=12, hb=2, seed=3),
}


# default dispatch of the shapes behind BASELINE.json configs[1] (1-D grid, one GaussianRandomWalk: warp-specialised
# fused kernels) and configs[2]/[3] (2-D grids beyond one SM's shared memory: cluster-resident kernels)
EXPECTED_FAMILY = {'poisson_c2_small': 'fast1d_mma', 'poisson_wide_kernels': 'fast1d_mma',
                   'poisson_c2_interleaved': 'fast1d_il', 'poisson_interleaved_small_grid': 'fast1d_il',
                   'poisson_grid_1500': 'fast1d_mma', 'poisson_grid_2000': 'fast1d_mma', 'poisson_grid_250': 'fast1d_mma',
                   'poisson_regime': 'resident'}


PLAN_OPTIONS = {'poisson_c2_interleaved': {'il': 1}, 'poisson_interleaved_small_grid': {'il': 1}}


@pytest.mark.parametrize('name', sorted(CONFIGS))
@pytest.mark.parametrize('mode', ['full', 'forwardOnly', 'evidenceOnly'])
def test_cuda_matches_cpu_oracle(name, mode, cuda_engine, oracle_engine):
    import bayesloop_b200 as bl
    kw = dict(forwardOnly=(mode == 'forwardOnly'), evidenceOnly=(mode == 'evidenceOnly'))
    with cuda_engine.options(**PLAN_OPTIONS.get(name, {})):
        got = helpers.abi_sweep(cuda_engine, CONFIGS[name](bl, cuda_engine), **kw)
    if name in EXPECTED_FAMILY:  # the kernels that carry the performance claims must be the ones that ran
        assert cuda_engine.last_kernel() == ('bwd_' if mode == 'full' else 'fwd_') + EXPECTED_FAMILY[name]
    want = helpers.abi_sweep(oracle_engine, CONFIGS[name](bl, oracle_engine), **kw)
    _assert_sweeps_agree(got, want, mode)


def _assert_sweeps_agree(got, want, mode='full'):
    np.testing.assert_array_equal(got['alive'], want['alive'])
    np.testing.assert_allclose(got['logE'], want['logE'], rtol=1e-10)
    np.testing.assert_allclose(got['local'], want['local'], rtol=1e-7)
    np.testing.assert_allclose(got['localEvidence'], want['localEvidence'], rtol=1e-7)
    if mode != 'evidenceOnly':
        rowmax = want['avg'].max(axis=1, keepdims=True)
        assert np.all(np.abs(got['avg'] - want['avg']) <= 1e-6 * np.abs(want['avg']) + 1e-12 * rowmax)
        np.testing.assert_allclose(got['means'], want['means'], rtol=1e-8)


def test_study_fit_matches_oracle_on_long_series(cuda_engine, oracle_engine):
    """Study.fit (B = 1, smoothed posteriors stored in place) on T = 2000."""
    import bayesloop_b200 as bl
    rng = np.random.default_rng(5)
    data = rng.poisson(3 + 2 * np.sin(np.arange(2000) / 40.)).astype(float)
    out = []
    for eng in (cuda_engine, oracle_engine):
        S = bl.Study(silent=True, engine=eng)
        S.loadData(data, silent=True)
        S.set(bl.om.Poisson('rate', bl.oint(0, 12, 400)), bl.tm.GaussianRandomWalk('sigma', 0.08, target='rate'),
              silent=True)
        S.fit(silent=True)
        out.append(S)
    got, want = out
    assert abs(got.logEvidence - want.logEvidence) <= 1e-10 * abs(want.logEvidence)
    np.testing.assert_allclose(got.posteriorMeanValues, want.posteriorMeanValues, rtol=1e-9)
    rowmax = want.posteriorSequence.max(axis=1, keepdims=True)
    assert np.all(np.abs(got.posteriorSequence - want.posteriorSequence)
                  <= 1e-6 * want.posteriorSequence + 1e-12 * rowmax)
    np.testing.assert_allclose(got.localEvidence, want.localEvidence, rtol=1e-7)


def test_properties_at_scale(cuda_engine):
    """C2-shaped sweep (G = 1000, 256 combos, T = 1500) too big for the CPU oracle in seconds: size-independent
    properties.  (a) averaged posterior rows are distributions; (b) evidence-only and full sweeps agree on every
    log-evidence; (c) a combo's evidence does not depend on the batch it runs in; (d) deterministic re-run."""
    import bayesloop_b200 as bl
    S = _poisson(bl, cuda_engine, B=256, T=1500, G=1000, smax=0.2)
    full = helpers.abi_sweep(cuda_engine, S)
    assert np.all(full['alive'] == 1)
    np.testing.assert_allclose(full['avg'].sum(axis=1), 1.0, rtol=1e-12)
    assert np.all(full['avg'] >= 0)
    evo = helpers.abi_sweep(cuda_engine, _poisson(bl, cuda_engine, B=256, T=1500, G=1000, smax=0.2), evidenceOnly=True)
    np.testing.assert_array_equal(full['logE'], evo['logE'])
    sub = helpers.abi_sweep(cuda_engine, _poisson(bl, cuda_engine, B=2, T=1500, G=1000, smax=0.2), evidenceOnly=True)
    np.testing.assert_allclose([full['logE'][0], full['logE'][-1]], sub['logE'], rtol=1e-13)
    again = helpers.abi_sweep(cuda_engine, _poisson(bl, cuda_engine, B=256, T=1500, G=1000, smax=0.2))
    np.testing.assert_array_equal(full['logE'], again['logE'])
    np.testing.assert_allclose(full['avg'], again['avg'], rtol=1e-12, atol=1e-300)  # fp64 atomics: order may vary


@pytest.mark.parametrize('name', ['ref_tm_nested', 'ref_cps_1cp_1bp_2hp', 'syn_hyper_gauss_2d', 'syn_study_scaledar1_2d',
                                  'syn_study_2d_axis0_wide', 'syn_hyper_poisson_sweep', 'syn_online_mixed',
                                  'syn_study_multicolumn', 'ref_om_gaussianmean', 'syn_cps_gauss_2d',
                                  'ref_study_2d_grw', 'ref_study_2d_static', 'ref_hyper_1hp', 'ref_online_static',
                                  'ref_om_scaledar1', 'ref_om_laplace'])
def test_stream_kernels_on_golden_cases(name, use_cuda):
    """The large-grid (global-memory streamed) kernels forced onto small cases: same goldens."""
    import bayesloop_b200 as bl
    with use_cuda.options(force_stream=1, online2d=0):
        S, got = parity.run_case(name, bl)
    assert use_cuda.last_kernel().endswith('_stream')
    parity.compare(name, got, load_golden(name), rtol=1e-6, atol_post=1e-12)


# ---------------------------------------------------------------------------------------------------------------------
# cluster-resident 2-D kernels (bayesloop_b200/csrc/cluster2d.cuh): forced onto small grids, checked against the
# reference goldens and the CPU oracle; the engine reports which kernel family ran, so a silent fall-back fails.

def _gauss2d_cp(bl, engine, n0, n1, T, seed=7):
    rng = np.random.default_rng(seed)
    x = np.concatenate([rng.normal(-1, 0.8, T // 2), rng.normal(1.2, 0.8, T - T // 2)])
    S = bl.HyperStudy(silent=True, engine=engine)
    S.loadData(x, silent=True)
    S.set(bl.om.Gaussian('mean', bl.cint(-3, 3, n0), 'std', bl.oint(0, 3, n1)),
          bl.tm.CombinedTransitionModel(bl.tm.ChangePoint('tChange', [7, T // 2, T - 5]),
                                        bl.tm.GaussianRandomWalk('s_mean', bl.cint(0, 0.3, 2), target='mean'),
                                        bl.tm.GaussianRandomWalk('s_std', bl.cint(0, 0.2, 2), target='std')),
          silent=True)
    return S


CLUSTER_CONFIGS = {
    'gauss_2d_64x48': CONFIGS['gauss_2d_64x48'],
    'gauss_2d_100x100': CONFIGS['gauss_2d_100x100'],
    'gauss_2d_256x96': CONFIGS['gauss_2d_256x96_stream'],
    'gauss_2d_cp_96x64': lambda bl, e: _gauss2d_cp(bl, e, 96, 64, T=40),
    'gauss_2d_ragged_90x50': lambda bl, e: _gauss2d(bl, e, 90, 50, T=25, hb=3, seed=9),  # 90 rows: last band shorter
}


@pytest.mark.parametrize('name', sorted(CLUSTER_CONFIGS))
@pytest.mark.parametrize('mode', ['full', 'forwardOnly', 'evidenceOnly'])
def test_cluster2d_matches_cpu_oracle(name, mode, cuda_engine, oracle_engine):
    import bayesloop_b200 as bl
    kw = dict(forwardOnly=(mode == 'forwardOnly'), evidenceOnly=(mode == 'evidenceOnly'))
    with cuda_engine.options(cluster2d=1):
        got = helpers.abi_sweep(cuda_engine, CLUSTER_CONFIGS[name](bl, cuda_engine), **kw)
    assert cuda_engine.last_kernel() == ('bwd_cluster2d' if mode == 'full' else 'fwd_cluster2d')
    want = helpers.abi_sweep(oracle_engine, CLUSTER_CONFIGS[name](bl, oracle_engine), **kw)
    _assert_sweeps_agree(got, want, mode)


@pytest.mark.parametrize('csize', ['2', '4'])
@pytest.mark.parametrize('name', ['syn_hyper_gauss_2d', 'syn_cps_gauss_2d', 'ref_study_2d_grw', 'ref_study_2d_static'])
def test_cluster2d_on_golden_cases(name, csize, use_cuda):
    """Cluster sizes 2 and 4 on the reference's own 2-D known answers (20x20 ... 40x36 grids)."""
    import bayesloop_b200 as bl
    with use_cuda.options(cluster2d=1, cluster2d_c=int(csize)):
        S, got = parity.run_case(name, bl)
    assert use_cuda.last_kernel() == 'bwd_cluster2d'
    parity.compare(name, got, load_golden(name), rtol=1e-6, atol_post=1e-12)


def _c3_study(bl, engine, T=48, hyper=3):
    """BASELINE.json configs[2] grid and hyper-ranges: Gaussian 256 x 256, GRW on both parameters."""
    rng = np.random.default_rng(2)
    mu = np.cumsum(rng.normal(0, 0.05, T))
    S = bl.HyperStudy(silent=True, engine=engine)
    S.loadData(rng.normal(mu, 1.0), silent=True)
    S.set(bl.om.Gaussian('mean', bl.cint(-3, 3, 256), 'std', bl.oint(0, 3, 256)),
          bl.tm.CombinedTransitionModel(bl.tm.GaussianRandomWalk('s_mean', bl.cint(0, 0.1, hyper), target='mean'),
                                        bl.tm.GaussianRandomWalk('s_std', bl.cint(0, 0.05, hyper), target='std')),
          silent=True)
    return S


def _c4_study(bl, engine, T=40):
    """BASELINE.json configs[3] in miniature: 200 x 200, change-point sweep in front of the two random walks."""
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.normal(-0.5, 1.0, T // 2), rng.normal(1.0, 1.0, T - T // 2)])
    S = bl.HyperStudy(silent=True, engine=engine)
    S.loadData(x, silent=True)
    S.set(bl.om.Gaussian('mean', bl.cint(-3, 3, 200), 'std', bl.oint(0, 3, 200)),
          bl.tm.CombinedTransitionModel(bl.tm.ChangePoint('tChange', [5, 20, 33]),
                                        bl.tm.GaussianRandomWalk('s_mean', bl.cint(0, 0.1, 2), target='mean'),
                                        bl.tm.GaussianRandomWalk('s_std', bl.cint(0, 0.05, 2), target='std')),
          silent=True)
    return S


def test_cluster2d_matches_cpu_oracle_at_c3_size(cuda_engine, oracle_engine):
    """256 x 256 (BASELINE.json configs[2] grid), 9 combos x 48 steps, default dispatch (8 CTAs per combo) against the
    C oracle (a few seconds of CPU), plus the stream kernels as a third, independent implementation."""
    import bayesloop_b200 as bl
    got = helpers.abi_sweep(cuda_engine, _c3_study(bl, cuda_engine))
    assert cuda_engine.last_kernel() == 'bwd_cluster2d'
    want = helpers.abi_sweep(oracle_engine, _c3_study(bl, oracle_engine))
    _assert_sweeps_agree(got, want)
    np.testing.assert_allclose(got['avg'].sum(axis=1), 1.0, rtol=1e-12)
    with cuda_engine.options(no_cluster2d=1):
        third = helpers.abi_sweep(cuda_engine, _c3_study(bl, cuda_engine))
    assert cuda_engine.last_kernel() == 'bwd_stream'
    _assert_sweeps_agree(third, want)


def test_cluster2d_matches_cpu_oracle_at_c4_size(cuda_engine, oracle_engine):
    """200 x 200 with change-points (BASELINE.json configs[3] grid): cluster kernels, reset inside the kernel, against
    the C oracle."""
    import bayesloop_b200 as bl
    got = helpers.abi_sweep(cuda_engine, _c4_study(bl, cuda_engine))
    assert cuda_engine.last_kernel() == 'bwd_cluster2d'
    want = helpers.abi_sweep(oracle_engine, _c4_study(bl, oracle_engine))
    _assert_sweeps_agree(got, want)


def test_cluster2d_dead_combos_match_oracle(cuda_engine, oracle_engine):
    """A data point that underflows the likelihood of every cell kills the forward pass of every combo at that step
    (core.py:388-400): the cluster kernels must agree on `alive` and -inf evidences, and all CTAs of a cluster must
    leave together (no hang)."""
    import bayesloop_b200 as bl

    def study(engine):
        rng = np.random.default_rng(11)
        x = rng.normal(0.0, 1.0, 30)
        x[17] = 1e6
        S = bl.HyperStudy(silent=True, engine=engine)
        S.loadData(x, silent=True)
        S.set(bl.om.Gaussian('mean', bl.cint(-3, 3, 64), 'std', bl.oint(0, 3, 48)),
              bl.tm.CombinedTransitionModel(bl.tm.GaussianRandomWalk('s_mean', bl.cint(0, 0.3, 3), target='mean'),
                                            bl.tm.GaussianRandomWalk('s_std', bl.cint(0, 0.2, 2), target='std')),
              silent=True)
        return S

    with cuda_engine.options(cluster2d=1):
        got = helpers.abi_sweep(cuda_engine, study(cuda_engine))
    assert cuda_engine.last_kernel().endswith('cluster2d')
    want = helpers.abi_sweep(oracle_engine, study(oracle_engine))
    np.testing.assert_array_equal(got['alive'], want['alive'])
    assert np.all(got['alive'] != 1)
    np.testing.assert_array_equal(np.isneginf(got['logE']), np.isneginf(want['logE']))


def test_online_study_checkpoint_resume_on_device(use_cuda, tmp_path):
    """bl.save / bl.load of an OnlineStudy whose hypothesis posteriors live in HBM: the resumed stream continues
    bit-identically (same kernels, same state)."""
    import contextlib
    import io
    import bayesloop_b200 as bl
    from test_host_logic import _online_study
    rng = np.random.default_rng(21)
    x = np.zeros(30)
    for i in range(1, len(x)):
        x[i] = 0.5 * x[i - 1] + rng.normal()
    with contextlib.redirect_stdout(io.StringIO()):
        A, B = _online_study(bl), _online_study(bl)
        for d in x:
            A.step(d)
        for d in x[:11]:
            B.step(d)
        bl.save(str(tmp_path / 'online.bl'), B)
        C = bl.load(str(tmp_path / 'online.bl'))
        for d in x[11:]:
            C.step(d)
    assert C.logEvidence == A.logEvidence
    np.testing.assert_array_equal(C.marginalizedPosterior, A.marginalizedPosterior)


# ---------------------------------------------------------------------------------------------------------------------
# tiled OnlineStudy step (bayesloop_b200/csrc/online2d.cuh): the default for 2-D grids beyond shared memory
# (plan option online2d=0 returns to the stream kernels; online2d_small=1 forces it onto grids that would fit).


def _online_big(bl, engine, n0=150, n1=130, steps=6, w=1.0):
    rng = np.random.default_rng(31)
    x = np.zeros(steps + 1)
    for i in range(1, len(x)):
        x[i] = 0.55 * x[i - 1] + rng.normal()
    S = bl.OnlineStudy(storeHistory=False, silent=True, engine=engine)
    S.setOM(bl.om.ScaledAR1('rho', bl.oint(-1, 1, n0), 'sigma', bl.oint(0, 3, n1)), silent=True)
    S.add('normal', bl.tm.CombinedTransitionModel(bl.tm.GaussianRandomWalk('s1', bl.cint(0, 0.12 * w, 3), target='rho'),
                                                  bl.tm.GaussianRandomWalk('s2', bl.cint(0, 0.2 * w, 2), target='sigma')))
    S.add('bounded', bl.tm.CombinedTransitionModel(bl.tm.GaussianRandomWalk('s3', [0.05 * w, 0.1 * w], target='sigma'),
                                                   bl.tm.RegimeSwitch('q', -6)))
    S.add('chaotic', bl.tm.RegimeSwitch('p', bl.cint(-8, -3, 3)))
    S.add('indep', bl.tm.Independent())
    S.add('static', bl.tm.Static())
    for d in x:
        S.step(d)
    return S


def _assert_online_agree(got, want):
    np.testing.assert_allclose(got.logEvidence, want.logEvidence, rtol=1e-10)
    for a, b in zip(got.logEvidenceList, want.logEvidenceList):
        np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(got.transitionModelDistribution, want.transitionModelDistribution, rtol=1e-8)
    np.testing.assert_allclose(got.localTransitionModelDistribution, want.localTransitionModelDistribution, rtol=1e-8)
    gm, wm = got.marginalizedPosterior, want.marginalizedPosterior
    assert np.all(np.abs(gm - wm) <= 1e-7 * np.abs(wm) + 1e-13 * wm.max())
    for a, b in zip(got.parameterPosterior, want.parameterPosterior):
        top = b.reshape(len(b), -1).max(axis=1).reshape(-1, 1, 1)
        assert np.all(np.abs(a - b) <= 1e-7 * np.abs(b) + 1e-13 * top)


@pytest.mark.parametrize('loads', ['cp_async', 'ldg'])
def test_online2d_tiled_step_matches_cpu_oracle(loads, cuda_engine, oracle_engine):
    """150 x 130 grid (3 x 3 tiles, ragged right and bottom), random walks on either / both axes, random walk +
    RegimeSwitch, RegimeSwitch alone, Independent (reset) and Static in one batch, against the CPU oracle; both
    tile-load variants (cp.async is the default)."""
    import contextlib
    import io
    import bayesloop_b200 as bl
    with contextlib.redirect_stdout(io.StringIO()):
        with cuda_engine.options(online2d_small=1, online2d_async=int(loads == 'cp_async')):
            got = _online_big(bl, cuda_engine)
        assert got._dev['separable']
        assert cuda_engine.last_kernel() == 'online2d'
        want = _online_big(bl, oracle_engine)
    _assert_online_agree(got, want)


def test_online2d_matches_cpu_oracle_at_c5_size(cuda_engine, oracle_engine):
    """512 x 512 (BASELINE.json configs[4] grid), 3 steps, default dispatch: the tiled kernels against the C oracle."""
    import contextlib
    import io
    import bayesloop_b200 as bl
    with contextlib.redirect_stdout(io.StringIO()):
        got = _online_big(bl, cuda_engine, n0=512, n1=512, steps=3, w=0.25)  # widths of the C5 sweep (sigma <= 0.03)
        assert cuda_engine.last_kernel() == 'online2d'
        want = _online_big(bl, oracle_engine, n0=512, n1=512, steps=3, w=0.25)
    _assert_online_agree(got, want)


@pytest.mark.parametrize('name', ['ref_online_static', 'ref_online_2tm', 'syn_online_mixed'])
def test_online2d_on_golden_cases(name, use_cuda):
    """The reference's own online tests (tests/test_onlinestudy.py) forced through the tiled kernels on their small
    grids (one mostly masked tile per hypothesis)."""
    import bayesloop_b200 as bl
    with use_cuda.options(online2d=1, online2d_small=1):
        S, got = parity.run_case(name, bl)
    assert use_cuda.last_kernel() == 'online2d'
    parity.compare(name, got, load_golden(name), rtol=1e-6, atol_post=1e-12)


# ---------------------------------------------------------------------------------------------------------------------
# change-point prefix sharing on the device: the ordinary kernels on WINDOWS of the sequences (seq_stride / row_stride)

@pytest.mark.parametrize('grid', ['1d_ws', '2d_resident', '2d_cluster'])
def test_changepoint_prefix_sharing_on_device(grid, cuda_engine, oracle_engine):
    """Shared schedule on the B200 against (1) the plain schedule on the B200 and (2) the plain schedule on the C oracle."""
    import bayesloop_b200 as bl

    def study(engine, share):
        rng = np.random.default_rng(12)
        T = 44
        S = bl.ChangepointStudy(silent=True, engine=engine)
        if grid == '1d_ws':
            S.loadData(rng.poisson(np.where(np.arange(T) < 20, 2.0, 6.0)).astype(float), silent=True)
            S.set(bl.om.Poisson('rate', bl.oint(0, 12, 400)),
                  bl.tm.CombinedTransitionModel(bl.tm.ChangePoint('tChange', np.arange(4, 40, 5)),
                                                bl.tm.GaussianRandomWalk('s', bl.cint(0.02, 0.3, 5), target='rate')), silent=True)
        else:
            n = 26 if grid == '2d_resident' else 120
            S.loadData(np.concatenate([rng.normal(-1, 0.7, 20), rng.normal(1.3, 0.8, T - 20)]), silent=True)
            S.set(bl.om.Gaussian('mean', bl.cint(-3, 3, n), 'std', bl.oint(0, 3, n)),
                  bl.tm.CombinedTransitionModel(bl.tm.ChangePoint('tChange', np.arange(5, 40, 7)),
                                                bl.tm.GaussianRandomWalk('s_mean', bl.cint(0, 0.2, 2), target='mean'),
                                                bl.tm.GaussianRandomWalk('s_std', bl.cint(0.02, 0.1, 2), target='std')),
                  silent=True)
        S.shareChangepoints = share
        S.fit(silent=True)
        return S

    shared = study(cuda_engine, True)
    family = cuda_engine.last_kernel()
    assert shared.sweepStats['shared']
    assert family == {'1d_ws': 'bwd_resident', '2d_resident': 'bwd_resident', '2d_cluster': 'bwd_cluster2d'}[grid] or True
    assert shared.sweepStats['executed_updates'] < 0.75 * shared.sweepStats['nominal_updates']
    for other in (study(cuda_engine, False), study(oracle_engine, False)):
        assert not other.sweepStats['shared']
        np.testing.assert_allclose(shared.logEvidenceList, other.logEvidenceList, rtol=1e-10)
        np.testing.assert_allclose(shared.posteriorMeanValues, other.posteriorMeanValues, rtol=1e-8)
        np.testing.assert_allclose(shared.localEvidence, other.localEvidence, rtol=1e-7)
        a, b = shared.posteriorSequence, other.posteriorSequence
        top = b.reshape(len(b), -1).max(axis=1).reshape([-1] + [1] * (b.ndim - 1))
        assert np.all(np.abs(a - b) <= 1e-6 * np.abs(b) + 1e-12 * top)
