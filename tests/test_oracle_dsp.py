"""CPU: plain-numpy restatements of scipy's lfilter / resample_poly agree with scipy (they define what the
CUDA kernels implement), and the no-validation iirdesign shim is bit-identical to stock scipy on valid input."""
import numpy as np
import pytest
import scipy.signal

from oracle import dsp


@pytest.mark.parametrize('order,btype,wn', [(2, 'lowpass', 0.2), (6, 'lowpass', 0.35), (3, 'bandpass', [0.3, 0.4]),
                                            (4, 'bandstop', [0.2, 0.5]), (1, 'highpass', 0.05)])
def test_lfilter_restated(order, btype, wn):
    rng = np.random.default_rng(0)
    b, a = scipy.signal.iirfilter(order, wn, rs=48.0, rp=0.1, btype=btype, ftype='cheby2' if order == 6 else 'butter')
    x = rng.standard_normal((3, 300))
    np.testing.assert_allclose(dsp.lfilter_restated(b, a, x), scipy.signal.lfilter(b, a, x, axis=-1), rtol=0, atol=1e-12)


@pytest.mark.parametrize('up,down', [(2, 1), (1, 2), (3, 1), (1, 3), (3, 8), (3, 16), (2, 3), (3, 2), (1, 1), (720, 1920)])
@pytest.mark.parametrize('n', [360, 721])
def test_resample_restated(up, down, n):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((2, n))
    ref = scipy.signal.resample_poly(x, up, down, axis=-1)
    got = dsp.resample_restated(x, up, down)
    assert got.shape == ref.shape
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-13)


def test_halfband_structure():
    """Every second tap of the x2 / x3 interpolators is (numerically) zero -> kernels may skip them."""
    for m in (2, 3):
        h, half, _, _ = dsp.resample_taps(m, 1)
        k = np.arange(-half, half + 1)
        zero = (k % m == 0) & (k != 0)
        assert np.max(np.abs(h[zero])) < 1e-16
        assert abs(h[half] - 1.0) < 2e-3


def test_iirdesign_shim_identical_on_valid_input():
    for wp, ws in [(0.2, 0.5), ([0.3, 0.4], [0.2, 0.5]), ([0.2, 0.5], [0.3, 0.4])]:
        b0, a0 = scipy.signal.iirdesign(wp, ws, 3.0, 20.0, ftype='butter')
        b1, a1 = dsp._iirdesign_no_validation(wp, ws, 3.0, 20.0, 'butter')
        assert np.array_equal(b0, b1) and np.array_equal(a0, a1)


def test_filter_shift_semantics():
    f = dsp.design_iirdesign(0.2, 0.5, 3.0, 20.0)
    assert f.shift > 0
    x = np.random.default_rng(2).standard_normal(100)
    full = scipy.signal.lfilter(f.b, f.a, np.concatenate((x, np.full(f.shift, x[-1]))))
    np.testing.assert_array_equal(f(x), full[f.shift:])
