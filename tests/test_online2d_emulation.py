"""CPU check of the tiled OnlineStudy step (bayesloop_b200/csrc/online2d.cuh) without a GPU: the per-thread phases of
the kernel (online2d_phases.h, compiled here as plain C++) are run thread by thread, tile by tile, by
tools/emu/online2d_emu.cpp and compared with a direct whole-grid reflect convolution -- ragged tiles, radii of zero,
radii beyond the grid (multiple reflections), RegimeSwitch clamp, reset and pointwise hypotheses."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_kernel_phases_match_direct_convolution(tmp_path):
    exe = str(tmp_path / 'online2d_emu')
    subprocess.check_call(['g++', '-O2', '-std=c++17', os.path.join(ROOT, 'tools', 'emu', 'online2d_emu.cpp'), '-o', exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert 'all cases passed' in out.stdout


def test_separable_rows_promise(oracle_engine):
    """The host only promises BLG_F_SEPARABLE_ROWS for programs the tiled kernels can take."""
    import contextlib
    import io
    import bayesloop_b200 as bl

    def study(*models):
        S = bl.OnlineStudy(silent=True, engine=oracle_engine)
        S.setOM(bl.om.Gaussian('m', bl.cint(0, 6, 12), 's', bl.oint(0, 2, 10)), silent=True)
        for k, tm in enumerate(models):
            S.add('tm%d' % k, tm)
        for d in [1., 2., 3.]:
            S.step(d)
        return S._dev['separable']

    grw = bl.tm.GaussianRandomWalk
    with contextlib.redirect_stdout(io.StringIO()):
        assert study(bl.tm.Static())
        assert study(bl.tm.CombinedTransitionModel(grw('a', [0.1, 0.2], target='m'), grw('b', 0.1, target='s')),
                     bl.tm.RegimeSwitch('p', [-7, -5]), bl.tm.Independent())
        assert study(bl.tm.CombinedTransitionModel(grw('a', 0.2, target='s'), bl.tm.RegimeSwitch('p', -7)))
        assert not study(bl.tm.CombinedTransitionModel(bl.tm.RegimeSwitch('p', -7), grw('a', 0.2, target='m')))
        assert not study(bl.tm.CombinedTransitionModel(grw('a', 0.2, target='m'), grw('b', 0.1, target='m')))
        assert not study(bl.tm.NotEqual('q', -3))
        assert study(bl.tm.CombinedTransitionModel(grw('a', 0.0, target='m'), grw('b', 0.1, target='m')))  # width 0 = identity
