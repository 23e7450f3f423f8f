"""DSP kit of the oracle (test infrastructure only).

* ``Filt`` / ``design_*``  follow reference color_modem/utils.py:9-64 (FilterFunction and the
  iirfilter / iirdesign / iirdesign_wc / iirsplitter helpers).
* ``resample``            is scipy.signal.resample_poly applied along the last axis — the same
  third-party call the reference makes (qam.py:35,37,45,53-57 etc.).
* ``lfilter_restated`` / ``resample_restated`` restate scipy's ``lfilter`` (direct form II
  transposed) and ``resample_poly``/``upfirdn`` (Kaiser(5) windowed sinc, zero-padded, centred)
  in plain numpy.  They document exactly what the CUDA kernels implement and are checked
  against scipy in tests/test_oracle_dsp.py.
"""
import numpy as np
import scipy.signal
from scipy.signal import _filter_design as _fd


def _iirdesign_no_validation(wp, ws, gpass, gstop, ftype):
    """scipy.signal.iirdesign minus its argument validation.

    The reference clamps wp from below and ws only from above (utils.py:45-47), so for
    NTSC/NTSC-N/PAL-M/PAL-N the band-pass stop edge goes slightly negative; scipy >= 1.6
    rejects that in iirdesign, while buttord and iirfilter still accept it.
    On valid input this is bit-identical to scipy.signal.iirdesign.
    """
    wp = np.atleast_1d(wp)
    ws = np.atleast_1d(ws)
    ordfunc = _fd.filter_dict[ftype][1]
    band_type = 2 * (len(wp) - 1) + 1
    if wp[0] >= ws[0]:
        band_type += 1
    btype = {1: 'lowpass', 2: 'highpass', 3: 'bandstop', 4: 'bandpass'}[band_type]
    order, wn = ordfunc(wp, ws, gpass, gstop)
    return scipy.signal.iirfilter(order, wn, rp=gpass, rs=gstop, btype=btype, ftype=ftype, output='ba')


class Filt(object):
    """Causal IIR, zero initial state per line, integer group-delay compensation (utils.py:9-36)."""

    def __init__(self, b, a, wp, btype, shift):
        self.b = np.asarray(b, dtype=np.float64)
        self.a = np.asarray(a, dtype=np.float64)
        wp = np.atleast_1d(wp)
        stopish = btype.lower() in ('bs', 'bandstop', 'bands', 'stop')
        centre = float(np.average(wp)) if (len(wp) > 1 and not stopish) else 0.0   # utils.py:14-17
        if shift:
            gd = scipy.signal.group_delay((self.b, self.a), [centre], fs=2.0)[1]
            self.shift = int(np.round(gd[0]))                                       # utils.py:19-20
        else:
            self.shift = 0
        resp = scipy.signal.freqz(self.b, self.a, worN=[centre], fs=2.0)[1][0]
        self.phase_shift = (np.angle(resp) + self.shift * np.pi * centre) % (2.0 * np.pi)  # utils.py:24-26
        assert self.shift >= 0, 'negative group delay never occurs for the supported presets'

    def __call__(self, x):
        """utils.py:28-36 along the last axis: append shift copies of the last sample, filter, drop head."""
        x = np.asarray(x, dtype=np.float64)
        if self.shift == 0:
            return scipy.signal.lfilter(self.b, self.a, x, axis=-1)
        tail = np.repeat(x[..., -1:], self.shift, axis=-1)
        return scipy.signal.lfilter(self.b, self.a, np.concatenate((x, tail), axis=-1), axis=-1)[..., self.shift:]


def design_iirfilter(order, wn, rp=None, rs=None, btype='band', ftype='butter', shift=True):
    b, a = scipy.signal.iirfilter(order, wn, rp, rs, btype, ftype=ftype)            # utils.py:39-41
    return Filt(b, a, wn, btype, shift)


def design_iirdesign(wp, ws, gpass, gstop, ftype='butter', shift=True):
    lo = np.nextafter(0.0, 1.0)
    hi = np.nextafter(1.0, 0.0)
    b, a = _iirdesign_no_validation(np.maximum(wp, lo), np.minimum(ws, hi), gpass, gstop, ftype)  # utils.py:44-47
    btype = 'band'
    if len(np.atleast_1d(wp)) > 1 and len(np.atleast_1d(ws)) > 1 and ws[0] > wp[0]:
        btype = 'bandstop'                                                              # utils.py:48-50
    return Filt(b, a, wp, btype, shift)


def design_band(wc, wp, ws, gpass, gstop, ftype='butter', shift=True):
    return design_iirdesign([wc - wp, wc + wp], [wc - ws, wc + ws], gpass, gstop, ftype, shift)  # utils.py:54-55


def design_splitter(wc, wp, ws, gpass, gstop):
    """Band-pass / complementary band-stop pair, utils.py:58-64."""
    def inv_db(db):
        return -(20.0 * np.log10(1.0 - 10.0 ** (-db / 20.0)))
    return (design_band(wc, wp, ws, gpass, gstop),
            design_band(wc, ws, wp, inv_db(gstop), inv_db(gpass)))


def resample(x, up, down):
    return scipy.signal.resample_poly(np.asarray(x, dtype=np.float64), up=up, down=down, axis=-1)


def carrier_ramp(start, step, n):
    """``linspace(start, start + n*step, n, endpoint=False) % 2pi`` per line (qam.py:24-25,47-48).

    start: [L] array -> [L, n].
    """
    start = np.atleast_1d(np.asarray(start, dtype=np.float64))
    return np.linspace(start, start + n * step, num=n, endpoint=False, axis=-1) % (2.0 * np.pi)


# ------------------------------------------------------------------------------------------------
# plain-numpy restatements of the third-party kernels (what the CUDA side implements)
# ------------------------------------------------------------------------------------------------

def lfilter_restated(b, a, x):
    """Direct form II transposed, zero initial state (scipy.signal.lfilter's recursion)."""
    b = np.asarray(b, dtype=np.float64) / a[0]
    a = np.asarray(a, dtype=np.float64) / a[0]
    n = max(len(a), len(b))
    b = np.concatenate((b, np.zeros(n - len(b))))
    a = np.concatenate((a, np.zeros(n - len(a))))
    x = np.asarray(x, dtype=np.float64)
    z = np.zeros(x.shape[:-1] + (n,), dtype=np.float64)
    y = np.empty_like(x)
    for i in range(x.shape[-1]):
        xi = x[..., i]
        yi = b[0] * xi + z[..., 0]
        for k in range(1, n):
            z[..., k - 1] = b[k] * xi - a[k] * yi + z[..., k]
        y[..., i] = yi
    return y


def resample_taps(up, down):
    """Taps of resample_poly's default filter: firwin(2*half+1, 1/max, kaiser 5) * up, half = 10*max."""
    g = np.gcd(up, down)
    up, down = up // g, down // g
    m = max(up, down)
    half = 10 * m
    h = scipy.signal.firwin(2 * half + 1, 1.0 / m, window=('kaiser', 5.0)) * up
    return h, half, up, down


def resample_restated(x, up, down):
    """out[j] = sum_i x[i] * h[half + j*down - i*up], n_out = ceil(n*up/down); zero outside [0, n)."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[-1]
    if up == down:
        return x.copy()
    h, half, up, down = resample_taps(up, down)
    n_out = -(-n * up // down)
    out = np.zeros(x.shape[:-1] + (n_out,), dtype=np.float64)
    for j in range(n_out):
        c = half + j * down
        i_lo = max(0, -(-(c - 2 * half) // up))
        i_hi = min(n - 1, c // up)
        if i_hi < i_lo:
            continue
        idx = np.arange(i_lo, i_hi + 1)
        out[..., j] = x[..., idx] @ h[c - idx * up]
    return out
