"""The zoom arithmetic that k_rescale mirrors (oracle/zoom_ref.py) == the installed scipy.ndimage.zoom, bit for bit;
the block means == zoom_array (oracle/pileup_oracle.py) up to the summation order of np.mean."""
import numpy as np
import pytest
from scipy.ndimage import zoom

from oracle.pileup_oracle import zoom_array
from oracle.zoom_ref import zoom_array_ref, zoom_linear_1d, zoom_linear_2d


@pytest.mark.parametrize("shape,out", [((7, 7), (9, 9)), ((3, 5), (11, 11)), ((1, 4), (9, 9)), ((13, 2), (18, 9)),
                                        ((40, 31), (45, 36)), ((25, 25), (27, 27)), ((2, 2), (15, 15)), ((9, 9), (9, 9)),
                                        ((64, 50), (99, 99)), ((100, 7), (198, 11))])
def test_linear_zoom_equals_scipy_bit_for_bit(shape, out):
    rng = np.random.default_rng(shape[0] * 131 + shape[1])
    D = rng.gamma(0.5, 3.0, size=shape)
    D[rng.random(shape) < 0.2] = 0.0
    want = zoom(D, np.array(out) / np.array(shape) + 0.0000001, order=1)
    assert want.shape == tuple(out)
    got = zoom_linear_2d(D, out)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("n,m", [(5, 9), (1, 7), (30, 33), (17, 99), (2, 11)])
def test_linear_zoom_1d_equals_scipy(n, m):
    v = np.random.default_rng(n + m).random(n) * 100
    want = zoom(v, np.array([m]) / np.array([n]) + 0.0000001, order=1)
    assert np.array_equal(zoom_linear_1d(v, m), want)


@pytest.mark.parametrize("shape,rs", [((7, 7), 9), ((30, 30), 9), ((100, 37), 11), ((19, 64), 9), ((250, 250), 99),
                                       ((95, 80), 9)])
def test_block_means_equal_zoom_array(shape, rs):
    rng = np.random.default_rng(shape[0] + rs)
    D = rng.gamma(0.5, 3.0, size=shape)
    want = zoom_array(D, (rs, rs))
    got = zoom_array_ref(D, rs)
    np.testing.assert_allclose(got, want, rtol=4e-16, atol=0)
    if -(-shape[1] // rs) < 8:  # np.mean stays sequential below 8 elements per block: identical bits
        assert np.array_equal(got, want)
