"""QAM core design record: drop-in for ``color_modem.qam`` (qam.py:10-72) on the host side."""
import collections

import numpy

from . import _native as N
from . import _slots as S
from . import utils
from .modem import GpuModem

QamConfig = collections.namedtuple('QamConfig', ['fsc', 'bandwidth3db', 'bandwidth20db'])


class QamColorModem(object):
    """The quadrature modem core (qam.py:14-58): filter designs and carrier step — the record the NTSC / PAL handles are
    built from — plus the reference's own line-level calls with an explicit start phase, run on the GPU through a
    band-split handle whose frame / line phase shifts are zero (cm_window.phase_offset carries ``start_phase``)."""

    def __init__(self, wc, wp, ws, gpass, gstop, precision='fp64'):
        self.wc = wc
        self.precision = precision
        self.carrier_phase_step = 0.5 * numpy.pi * wc
        self._chroma_precorrect_lowpass = utils.iirdesign(wp, ws, gpass, gstop)
        self._extract_chroma2x, self._remove_chroma2x = utils.iirsplitter(0.5 * wc, 0.5 * wp, 0.5 * ws, gpass, gstop)
        self._demod_lowpass = utils.iirfilter(6, wc - 0.5 * ws, rs=48.0, btype='lowpass', ftype='cheby2')
        self._line_modems = {}

    @property
    def extract_chroma_phase_shift(self):
        return self._extract_chroma2x.phase_shift

    def _line_modem(self, n):
        m = self._line_modems.get(n)
        if m is None:
            m = self._line_modems[n] = _QamLine(self, n, self.precision)
        return m

    def _modulate_chroma(self, start_phase, u, v):
        """qam.py:20-26"""
        if len(u) != len(v):
            raise AssertionError('u and v differ in length')
        return self.modulate(start_phase, numpy.zeros(len(u)), u, v)

    def modulate(self, start_phase, y, u, v):
        """qam.py:28-32"""
        if not (len(y) == len(u) == len(v)):
            raise AssertionError('y, u and v differ in length')
        rows = numpy.stack([numpy.asarray(y, dtype=numpy.float64), numpy.asarray(u, dtype=numpy.float64),
                            numpy.asarray(v, dtype=numpy.float64)], axis=-1)[None]
        return self._line_modem(len(y))._run_window(True, 0, 0, rows, 0, phase=start_phase)

    def extract_chroma(self, composite):
        """qam.py:34-37"""
        rows = numpy.asarray(composite, dtype=numpy.float64)[None]
        return self._line_modem(rows.shape[1])._run_window(False, 0, 0, rows, 0, mode=N.MODE_EXTRACT_CHROMA)[:, 0]

    def demodulate(self, start_phase, composite, strip_chroma=True):
        """qam.py:43-58"""
        rows = numpy.asarray(composite, dtype=numpy.float64)[None]
        yuv = self._line_modem(rows.shape[1])._run_window(
            False, 0, 0, rows, 0, mode=N.MODE_DEFAULT if strip_chroma else N.MODE_BANDSPLIT_NOSTRIP, phase=start_phase)
        return yuv[:, 0], yuv[:, 1], yuv[:, 2]


def put_filter(desc, slot, ff, n, rate):
    f = desc.filters[slot]
    sos = ff.sos
    if sos.shape[0] > N.MAX_SECTIONS:
        raise ValueError('filter order %d exceeds the CUDA cascade limit' % (2 * sos.shape[0]))
    f.nsec, f.shift, f.n, f.rate = sos.shape[0], ff.shift, int(n), int(rate)
    for s in range(sos.shape[0]):
        for k in range(5):
            f.sos[s][k] = float(sos[s, k])
    desc.nfilters = max(desc.nfilters, slot + 1)


def put_resampler(desc, slot, up, down):
    taps, half, up, down = utils.resampler_taps(up, down)
    r = desc.resamplers[slot]
    r.up, r.down, r.half, r.ntaps = up, down, half, len(taps)
    if len(taps) > N.MAX_TAPS:
        raise ValueError('resampler %d/%d needs %d taps (> %d)' % (up, down, len(taps), N.MAX_TAPS))
    for i, t in enumerate(taps):
        r.taps[i] = float(t)
    desc.nresamplers = max(desc.nresamplers, slot + 1)


def fill_qam_filters(d, q, W):
    """The filter / resampler / carrier-step slots of the QAM family from a QamColorModem design."""
    put_filter(d, S.QF_PRE_LP, q._chroma_precorrect_lowpass, W, 1)
    put_filter(d, S.QF_BP2X, q._extract_chroma2x, 2 * W, 2)
    put_filter(d, S.QF_BS2X, q._remove_chroma2x, 2 * W, 2)
    put_filter(d, S.QF_DEMOD_LP, q._demod_lowpass, 2 * W, 2)
    put_resampler(d, S.QR_UP2, 2, 1)
    put_resampler(d, S.QR_DOWN2, 1, 2)
    d.phases[S.QP_STEP1X] = utils.turns_fixed(q.wc / 2.0)
    d.phases[S.QP_STEP2X] = utils.turns_fixed(q.wc / 4.0)
    d.phases[S.QP_BP_SHIFT] = utils.radians_fixed(q.extract_chroma_phase_shift)


class _QamLine(GpuModem):
    """One-row band-split handle of a bare QamColorModem: identity matrices, no frame / line phase shift."""

    def __init__(self, qam, n, precision):
        class _Raster(object):
            size = (n, 1)
            _line_shift = 0

            class line_standard(object):
                odd_field_first_active_line = even_field_first_active_line = 0
        if n % 4:
            raise NotImplementedError('line lengths must be multiples of 4 samples')
        GpuModem.__init__(self, _Raster, precision)
        self._qam = qam

    def _fill_desc(self, d):
        d.kind, d.flags = N.KIND_QAM_BANDSPLIT, 0
        for i in range(9):
            d.enc_matrix[i] = d.dec_matrix[i] = 1.0 if i % 4 == 0 else 0.0
        fill_qam_filters(d, self._qam, self.width)


class AbstractQamColorModem(GpuModem, utils.ConstantFrequencyCarrier):
    """NTSC/PAL-style modem on the GPU (qam.py:61-72).  Subclasses set ENC / DEC matrices, kind and flags."""
    ENC = None
    DEC = None
    kind = N.KIND_QAM_BANDSPLIT
    flags = 0
    has_demodulate_components = True
    _notch_q = 0.0        # comb decoders: Q of the luma notch (comb.py:18-20); 0 = no notch
    _minavg = False       # 3-line decoders: avg=comb.minavg (comb.py:13-15)

    def __init__(self, line_config, config, precision='fp32'):
        GpuModem.__init__(self, line_config, precision)
        self.config = config
        fs = line_config.fs
        self.qam = QamColorModem(2.0 * config.fsc / fs, 2.0 * config.bandwidth3db / fs,
                                 2.0 * config.bandwidth20db / fs, 3.0, 20.0, precision=precision)

    # host-side helpers of the reference protocol (tiny 3x3 matrix products; not on the hot path)
    @classmethod
    def encode_components(cls, r, g, b):
        m = numpy.asarray(cls.ENC, dtype=numpy.float64).reshape(3, 3)
        r, g, b = (numpy.asarray(x, dtype=numpy.float64) for x in (r, g, b))
        return tuple(m[i, 0] * r + m[i, 1] * g + m[i, 2] * b for i in range(3))

    @classmethod
    def decode_components(cls, y, u, v):
        m = numpy.asarray(cls.DEC, dtype=numpy.float64).reshape(3, 3)
        y, u, v = (numpy.asarray(x, dtype=numpy.float64) for x in (y, u, v))
        return tuple(m[i, 0] * y + m[i, 1] * u + m[i, 2] * v for i in range(3))

    def _fill_desc(self, d):
        W = self.width
        std = self.line_config.line_standard
        fsc = self.config.fsc
        d.kind, d.flags = self.kind, self._flags()
        d.frame_cycle = self.frame_cycle
        d.frame_shift_turns = utils.turns_fixed((fsc / std.frame_rate) % 1.0)
        d.line_shift_turns = utils.turns_fixed((fsc / (std.frame_rate * std.total_lines)) % 1.0)
        for i in range(9):
            d.enc_matrix[i] = self.ENC[i]
            d.dec_matrix[i] = self.DEC[i]
        fill_qam_filters(d, self.qam, W)
        d.phases[S.QP_HALF_LS] = utils.radians_fixed(0.5 * self.line_shift)
        if self._notch_q:
            import scipy.signal
            b, a = scipy.signal.iirnotch(2.0 * fsc / self.line_config.fs, self._notch_q)        # comb.py:18-20
            put_filter(d, S.QF_NOTCH, utils.FilterFunction(b, a, wp=0.0, btype='bandstop', shift=True), W, 1)
            d.flags |= N.FLAG_NOTCH
        if self._minavg:
            d.flags |= N.FLAG_MINAVG

    def _flags(self):
        return self.flags
