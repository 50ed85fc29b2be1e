"""The index arithmetic of the DMMA 1-D kernels (bayesloop_b200/csrc/fast1d_mma.cuh), restated in NumPy and checked on
the CPU against scipy.ndimage.gaussian_filter1d (what transitionModels.py:111 calls): the Toeplitz-block formulation of
the reflect convolution -- 8 x 8 output tiles, groups of 8 input offsets starting at the even offset s0 <= -R, the
zero-padded weight table, the mirrored halo -- must reproduce the filter for every radius, including the number of groups
the kernel loops over and the buffer extents it reads.  The GPU suite checks the kernels themselves; this test pins the
formulation they share with the LPT cost model of engine.py."""
import numpy as np
import pytest
from scipy.ndimage import gaussian_filter1d

W_PAD = 16  # kMmaWPad


def halo_of(R):  # mma_halo()
    return (R + 7 + 7) & ~7


def groups_of(R):  # mma_conv_body / engine.Program._assignment
    return (2 * R + 15 + (R & 1)) >> 3


def conv_by_tiles(x, sigma):
    """One convolution step exactly as the compute warps index it (fragment by fragment, no vectorisation)."""
    n = len(x)
    R = int(4.0 * sigma + 0.5)
    w = np.exp(-0.5 * (np.arange(-R, R + 1) / sigma) ** 2)
    w /= w.sum()
    halo = halo_of(R)
    ntiles = (n + 63) // 64
    pitch = halo + 64 * ntiles + halo + 8
    buf = np.zeros(pitch)  # cells beyond the mirrored halo stay zero (read with zero weights only)
    line = halo  # interior offset
    for g in range(n):
        buf[line + g] = x[g]
        if g < halo:
            buf[line - 1 - g] = x[g]
        if g >= n - halo:
            buf[line + 2 * n - 1 - g] = x[g]
    wz = np.zeros(2 * R + 1 + 2 * W_PAD)
    wz[W_PAD:W_PAD + 2 * R + 1] = w
    s0 = -(R + (R & 1))
    y = np.zeros(64 * ntiles)
    lo, hi = 0, 0
    for tile in range(0, 64 * ntiles, 64):
        Y = np.zeros((8, 8))
        for j in range(groups_of(R)):
            for odd in (0, 1):
                A = np.zeros((8, 4))
                B = np.zeros((4, 8))
                for a in range(8):
                    for u in range(4):
                        idx = line + tile + 8 * a + s0 + 8 * j + 2 * u + odd
                        lo, hi = min(lo, idx - line), max(hi, idx - line)
                        A[a, u] = buf[idx]
                for u in range(4):
                    for r in range(8):
                        B[u, r] = wz[s0 + 8 * j + 2 * u + odd - r + R + W_PAD]
                Y += A @ B
        y[tile:tile + 64] = Y.reshape(64)
    assert lo >= -halo and hi < pitch - halo, 'fragment loads leave the state buffer'
    return y[:n], R


@pytest.mark.parametrize('n,sigma', [(200, 0.8), (200, 3.3), (1000, 16.7), (1000, 8.4), (130, 4.0), (64, 1.9), (1000, 0.3)])
def test_toeplitz_blocks_reproduce_the_reflect_filter(n, sigma):
    rng = np.random.default_rng(int(n * 10 + sigma * 7))
    x = rng.random(n) ** 3
    got, R = conv_by_tiles(x, sigma)
    assert halo_of(R) <= n
    want = gaussian_filter1d(x, sigma)  # mode='reflect', truncate=4.0: SciPy's defaults, as in the reference
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-15)


def test_group_count_is_minimal_for_every_radius():
    for R in range(1, 400):
        s0 = -(R + (R & 1))
        g = groups_of(R)
        assert s0 % 2 == 0 and s0 <= -R
        assert s0 + 8 * g - 1 >= R + 7          # the 8 outputs of a fragment row need the offsets -R .. R+7
        assert s0 + 8 * (g - 1) - 1 < R + 7     # ... and one group fewer would not cover them
        assert g <= 2 * ((R + 7) // 8) + 1      # never more than groups aligned at multiples of 8
        # weight-table reads stay inside the zero padding
        assert s0 + 0 - 7 + R + W_PAD >= 0 and s0 + 8 * (g - 1) + 7 + R + W_PAD < 2 * R + 1 + 2 * W_PAD
