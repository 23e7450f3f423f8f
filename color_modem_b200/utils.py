"""Host-side DSP design kit: drop-in for the *design* half of the reference's ``color_modem.utils`` (utils.py:9-88).

The reference designs its filters with scipy at construction time and then runs them line by line on the CPU.
Here the design stays on the host with the very same scipy calls (so coefficients cannot disagree with the
reference), is converted to second-order sections in float64 and shipped to the CUDA kernels through cm_desc;
the filtering itself only exists on the GPU (csrc/cm_iir.cuh).
"""
import fractions

import numpy
import scipy.signal
from scipy.signal import _filter_design as _fd

TWO_PI = 2.0 * numpy.pi


class FilterFunction(object):
    """One IIR use-site (reference utils.py:9-36): the design record the kernels are fed from, and — called — the
    reference's delay-compensated causal filter, run on the GPU (csrc/cm_util.cu: cm_filter_rows)."""

    def __init__(self, b, a, wp, btype, shift):
        self._b = numpy.asarray(b, dtype=numpy.float64)
        self._a = numpy.asarray(a, dtype=numpy.float64)
        wp = numpy.atleast_1d(wp)
        if len(wp) > 1 and btype.lower() not in ('bs', 'bandstop', 'bands', 'stop'):
            shiftfreq = float(numpy.average(wp))
        else:
            shiftfreq = 0.0
        if shift:
            delay = scipy.signal.group_delay((self._b, self._a), [shiftfreq], fs=2.0)[1]
            self._shift = int(numpy.round(delay[0]))
        else:
            self._shift = 0
        if self._shift < 0:
            raise ValueError('negative group delay compensation is not supported by the CUDA filter')
        response = scipy.signal.freqz(self._b, self._a, worN=[shiftfreq], fs=2.0)[1][0]
        self.phase_shift = (numpy.angle(response) + self._shift * numpy.pi * shiftfreq) % TWO_PI

    @property
    def shift(self):
        return self._shift

    @property
    def sos(self):
        """Second-order sections [nsec, 5] = b0 b1 b2 a1 a2 (a0 normalised to 1), float64."""
        sos = scipy.signal.tf2sos(self._b, self._a)
        sos = sos / sos[:, 3:4]
        return numpy.ascontiguousarray(sos[:, [0, 1, 2, 4, 5]])

    def __call__(self, x, precision='fp64'):
        """utils.py:28-36 along the last axis of ``x`` (zero initial state, ``shift`` copies of the last sample appended,
        the first ``shift`` outputs dropped).  numpy in, float64 numpy out; the recursion runs in ``precision``."""
        import ctypes as C
        import torch
        from . import _native as N
        if not torch.cuda.is_available():
            raise N.NativeUnavailable('no CUDA device: color_modem_b200 has no CPU path')
        x = numpy.asarray(x, dtype=numpy.float64)
        rows = numpy.ascontiguousarray(x.reshape(-1, x.shape[-1]), dtype=numpy.float32 if precision == 'fp32' else numpy.float64)
        if rows.shape[1] == 0 or rows.shape[0] == 0:
            return numpy.array(x)
        f = N.Filter()
        sos = self.sos
        if sos.shape[0] > N.MAX_SECTIONS:
            raise ValueError('filter order %d exceeds the CUDA cascade limit' % (2 * sos.shape[0]))
        f.nsec, f.shift, f.n, f.rate = sos.shape[0], self._shift, rows.shape[1], 1
        for s_ in range(sos.shape[0]):
            for k in range(5):
                f.sos[s_][k] = float(sos[s_, k])
        tin = torch.from_numpy(rows).cuda()
        tout = torch.empty_like(tin)
        N.check(N.load().cm_filter_rows(C.byref(f), N.FP32 if precision == 'fp32' else N.FP64, tin.data_ptr(),
                                        tout.data_ptr(), rows.shape[0], torch.cuda.current_stream().cuda_stream))
        return tout.double().cpu().numpy().reshape(x.shape)


def _scipy_iirdesign(wp, ws, gpass, gstop, ftype):
    """scipy.signal.iirdesign without its wp/ws range validation (identical arithmetic).

    The reference clamps only one side of each edge (utils.py:45-47), which for NTSC / PAL-M / PAL-N puts the
    band-pass stop edge marginally below zero; scipy >= 1.6 refuses that in iirdesign although buttord and
    iirfilter handle it.  See SURVEY.md §8c."""
    wp = numpy.atleast_1d(wp)
    ws = numpy.atleast_1d(ws)
    order_fn = _fd.filter_dict[ftype][1]
    band = 2 * (len(wp) - 1) + 1 + (1 if wp[0] >= ws[0] else 0)
    btype = {1: 'lowpass', 2: 'highpass', 3: 'bandstop', 4: 'bandpass'}[band]
    order, wn = order_fn(wp, ws, gpass, gstop)
    return scipy.signal.iirfilter(order, wn, rp=gpass, rs=gstop, btype=btype, ftype=ftype, output='ba')


def iirfilter(N, Wn, rp=None, rs=None, btype='band', ftype='butter', shift=True):
    b, a = scipy.signal.iirfilter(N, Wn, rp, rs, btype, ftype=ftype)
    return FilterFunction(b, a, Wn, btype, shift)


def iirdesign(wp, ws, gpass, gstop, ftype='butter', shift=True):
    tiny = numpy.nextafter(0.0, 1.0)
    below_one = numpy.nextafter(1.0, 0.0)
    b, a = _scipy_iirdesign(numpy.maximum(wp, tiny), numpy.minimum(ws, below_one), gpass, gstop, ftype)
    two_edges = len(numpy.atleast_1d(wp)) > 1 and len(numpy.atleast_1d(ws)) > 1
    btype = 'bandstop' if (two_edges and ws[0] > wp[0]) else 'band'
    return FilterFunction(b, a, wp, btype, shift)


def iirdesign_wc(wc, wp, ws, gpass, gstop, ftype='butter', shift=True):
    return iirdesign([wc - wp, wc + wp], [wc - ws, wc + ws], gpass, gstop, ftype, shift)


def iirsplitter(wc, wp, ws, gpass, gstop, ftype='butter', shift=True):
    def complement_db(db):
        return -(20.0 * numpy.log10(1.0 - 10.0 ** (-db / 20.0)))

    return (iirdesign_wc(wc, wp, ws, gpass, gstop, ftype, shift),
            iirdesign_wc(wc, ws, wp, complement_db(gstop), complement_db(gpass), ftype, shift))


def resampler_taps(up, down):
    """Taps of scipy.signal.resample_poly's default filter for (up, down): (taps, half, up', down')."""
    g = int(numpy.gcd(up, down))
    up, down = up // g, down // g
    biggest = max(up, down)
    half = 10 * biggest
    taps = scipy.signal.firwin(2 * half + 1, 1.0 / biggest, window=('kaiser', 5.0)) * up
    return taps, half, up, down


def turns_fixed(turns):
    """Fraction of a turn -> 0.64 fixed point (wraps)."""
    frac = float(turns) % 1.0
    return int(frac * 18446744073709551616.0) & 0xFFFFFFFFFFFFFFFF


def radians_fixed(rad):
    return turns_fixed(float(rad) / TWO_PI)


class ConstantFrequencyCarrier(object):
    """Closed-form subcarrier phase at the start of a line (reference utils.py:67-88)."""

    @property
    def line_shift(self):
        std = self.line_config.line_standard
        return TWO_PI * ((self.config.fsc / (std.frame_rate * std.total_lines)) % 1.0)

    @property
    def frame_shift(self):
        return TWO_PI * ((self.config.fsc / self.line_config.line_standard.frame_rate) % 1.0)

    @property
    def frame_cycle(self):
        ratio = self.config.fsc / self.line_config.line_standard.frame_rate
        return fractions.Fraction(ratio).limit_denominator().denominator

    def start_phase(self, frame, line):
        std = self.line_config.line_standard
        reference_line = min(std.odd_field_first_active_line, std.even_field_first_active_line)
        frame_part = ((frame % self.frame_cycle) * self.frame_shift) % TWO_PI
        line_part = ((self.line_config.analog_line(line) - reference_line) * self.line_shift) % TWO_PI
        return (frame_part + line_part) % TWO_PI
