"""SECAM presets and modem — drop-in for ``color_modem.color.secam`` (secam.py:10-304), computed on the GPU."""
import collections

import numpy
import scipy.signal

from .. import _native as N
from .. import _slots as S
from .. import utils
from ..modem import GpuModem
from ..qam import put_filter, put_resampler

SecamVariant = collections.namedtuple('SecamVariant',
                                      ['fsc_dr', 'fsc_db', 'fdev_dr', 'fdev_db', 'flimit_minbell', 'flimit_maxbell',
                                       'm0', 'bell_f0', 'bell_kn', 'bell_kd', 'lf_precorrect_f1', 'lf_precorrect_k'])

_NTSC_FSC = 227.5 * 15750.0 * 1000.0 / 1001.0

# possible SECAM I / II after Boetcher & Matzel; SECAM III as proposed; SECAM = IIIb, the broadcast system;
# A / E / M / N: test and regional variants (secam.py:14-124)
SecamVariant.SECAM_I = SecamVariant(4437500.0, 4437500.0, 250000.0, 250000.0, -250000.0, 250000.0,
                                    0.2, 4437500.0, 1.0, 1.0, 0.0, 1.0)
SecamVariant.SECAM_II = SecamVariant(4437500.0, 4437500.0, 250000.0, 250000.0, -250000.0, 250000.0,
                                     0.1, 4437500.0, 16.0, 1.26, 0.0, 1.0)
SecamVariant.SECAM_III = SecamVariant(4437500.0, 4437500.0, 230000.0, 230000.0, -450000.0, 350000.0,
                                      0.1, 4437500.0, 16.0, 1.26, 70000.0, 5.6)
SecamVariant.SECAM = SecamVariant(4406250.0, 4250000.0, 280000.0, 230000.0, -386000.0, 470250.0,
                                  0.115, 4286000.0, 16.0, 1.26, 85000.0, 3.0)
SecamVariant.SECAM_A = SecamVariant(2660000.0, 2660000.0, 250000.0, 250000.0, -250000.0, 250000.0,
                                    0.2, 2660000.0, 1.0, 1.0, 0.0, 1.0)
SecamVariant.SECAM_E = SecamVariant(8370000.0, 8370000.0, 250000.0, 250000.0, -250000.0, 250000.0,
                                    0.2, 8370000.0, 1.0, 1.0, 0.0, 1.0)
SecamVariant.SECAM_M = SecamVariant(_NTSC_FSC, _NTSC_FSC, 230000.0, 230000.0, -500000.0, 500000.0,
                                    0.1, _NTSC_FSC, 16.0, 1.26, 70000.0, 5.6)
SecamVariant.SECAM_N = SecamVariant(3578125.0, 3578125.0, 230000.0, 230000.0, -500000.0, 500000.0,
                                    0.1, 3578125.0, 16.0, 1.26, 70000.0, 5.6)


class FmDecoder(object):
    """Design record of the quadrature FM discriminator (secam.py:127-132)."""

    def __init__(self, fc, dev, resample_rate=2):
        if resample_rate != 2:
            raise NotImplementedError('the CUDA discriminator runs at 2x')
        self._fc = fc
        self._resample_rate = resample_rate
        self._lowpass = utils.iirfilter(6, (2.0 * fc - dev) / resample_rate, rs=48.0, btype='lowpass', ftype='cheby2')


class SecamModem(GpuModem):
    kind = N.KIND_SECAM
    decoder_rows = 2
    ENC = (0.299, 0.587, 0.114,
           -1.333302, 1.116474, 0.216828,
           -0.449995, -0.883435, 1.33343)
    DEC = (1.0, -0.5257623554153522, 0.0,
           1.0, 0.2678074007993021, -0.1290417517983779,
           1.0, 0.0, 0.6644518272425249)

    def __init__(self, line_config, variant=SecamVariant.SECAM, alternate_phases=False, precision='fp32'):
        super(SecamModem, self).__init__(line_config, precision)
        self._line_config = line_config
        self._variant = variant
        fs = line_config.fs
        self._fsc_dr = 2.0 * variant.fsc_dr / fs
        self._fsc_db = 2.0 * variant.fsc_db / fs
        self._fdev_dr = 2.0 * variant.fdev_dr / fs
        self._fdev_db = 2.0 * variant.fdev_db / fs
        self._flimit_min = 2.0 * (variant.bell_f0 + variant.flimit_minbell) / fs
        self._flimit_max = 2.0 * (variant.bell_f0 + variant.flimit_maxbell) / fs
        self._bell_f0 = 2.0 * variant.bell_f0 / fs
        self._start_phase_inversions = ([False, False, False, True, True, True] if alternate_phases
                                        else [False, False, True, False, False, True])
        self._chroma_demod_bell = None
        if variant.bell_kn != variant.bell_kd:
            self._chroma_demod_bell = self._chroma_demod_bell_design(self._bell_f0, self._flimit_max,
                                                                     variant.bell_kn, variant.bell_kd)
        self._chroma_precorrect_lowpass = utils.iirdesign(wp=2.0 * 1300000.0 / fs, ws=2.0 * 3500000.0 / fs,
                                                          gpass=3.0, gstop=30.0)
        self._chroma_precorrect = self._reverse_chroma_precorrect = None
        if variant.lf_precorrect_k != 1.0:
            self._chroma_precorrect, self._reverse_chroma_precorrect = self._chroma_precorrect_design(
                2.0 * variant.lf_precorrect_f1 / fs, variant.lf_precorrect_k)
        center = 0.5 * (self._flimit_min + self._flimit_max)
        dev = 0.5 * (self._flimit_max - self._flimit_min)
        self._chroma_demod_chroma_filter = utils.iirfilter(3, [center - dev, center + dev], rp=0.1,
                                                           btype='bandpass', ftype='cheby1')
        self._chroma_demod_luma_filter = utils.iirfilter(3, [center - dev * numpy.e, center + dev * numpy.e],
                                                         btype='bandstop', ftype='bessel')
        self._chroma_demod = FmDecoder(center, dev)

    @classmethod
    def encode_components(cls, r, g, b):
        m = numpy.asarray(cls.ENC).reshape(3, 3)
        r, g, b = (numpy.asarray(x, dtype=numpy.float64) for x in (r, g, b))
        return tuple(m[i, 0] * r + m[i, 1] * g + m[i, 2] * b for i in range(3))

    @classmethod
    def decode_components(cls, luma, dr, db):
        m = numpy.asarray(cls.DEC).reshape(3, 3)
        luma, dr, db = (numpy.asarray(x, dtype=numpy.float64) for x in (luma, dr, db))
        return tuple(m[i, 0] * luma + m[i, 1] * dr + m[i, 2] * db for i in range(3))

    @staticmethod
    def _chroma_precorrect_design(wc, k):
        """First-order LF pre-emphasis and its exact inverse (secam.py:210-221)."""
        if k == 1.0:
            raise AssertionError('k == 1 means no pre-emphasis')
        fwd_b, fwd_a = scipy.signal.iirfilter(1, k * wc, btype='highpass', ftype='butter')
        fwd_b[0] = (k - 1.0) * fwd_b[0] + 1.0
        fwd_b[1] = (k - 1.0) * fwd_b[1] + fwd_a[1]
        inv_b = numpy.array([1.0, fwd_a[1]]) / fwd_b[0]
        inv_a = numpy.array([1.0, fwd_b[1] / fwd_b[0]])
        return (utils.FilterFunction(fwd_b, fwd_a, k * wc, btype='highpass', shift=False),
                utils.FilterFunction(inv_b, inv_a, k * wc, btype='lowpass', shift=False))

    @staticmethod
    def _chroma_demod_bell_design(f0, f_max, kn, kd):
        """Band-pass approximating the inverse of the encoder's bell curve (secam.py:223-238)."""
        def gain_db(f):
            f2, f02 = f * f, f0 * f0
            num = kd * kd * f02 * f02 + (1 - 2 * kd * kd) * f2 * f02 + kd * kd * f2 * f2
            den = kn * kn * f02 * f02 + (1 - 2 * kn * kn) * f2 * f02 + kn * kn * f2 * f2
            return 10.0 * numpy.log10(numpy.sqrt(num / den))

        wp2 = f0 + 1 / 256.0
        wp1 = f0 * f0 / wp2
        ws2 = f_max
        ws1 = f0 * f0 / ws2
        return utils.iirdesign([wp1, wp2], [ws1, ws2], -gain_db(wp2), -gain_db(ws2), shift=False)

    def _flags(self):
        return ((N.FLAG_SECAM_BELL if self._chroma_demod_bell is not None else 0) |
                (N.FLAG_SECAM_LF if self._chroma_precorrect is not None else 0))

    def _fill_desc(self, d):
        W = self.width
        d.kind, d.flags = self.kind, self._flags()
        for i in range(9):
            d.enc_matrix[i] = self.ENC[i]
            d.dec_matrix[i] = self.DEC[i]
        ncc = W + W // 40 - 1
        put_filter(d, S.SF_PRE_LP, self._chroma_precorrect_lowpass, W, 1)
        put_filter(d, S.SF_LUMA_BS, self._chroma_demod_luma_filter, W, 1)
        put_filter(d, S.SF_CHROMA_BP, self._chroma_demod_chroma_filter, ncc, 1)
        put_filter(d, S.SF_FM_LP, self._chroma_demod._lowpass, 2 * ncc, 2)
        if self._chroma_precorrect is not None:
            put_filter(d, S.SF_PRE_EMPH, self._chroma_precorrect, W, 1)
            put_filter(d, S.SF_DE_EMPH, self._reverse_chroma_precorrect, W, 1)
        if self._chroma_demod_bell is not None:
            put_filter(d, S.SF_ANTI_BELL, self._chroma_demod_bell, ncc, 1)
        put_resampler(d, S.SR_UP2, 2, 1)
        put_resampler(d, S.SR_DOWN2, 1, 2)
        d.phases[S.SP_FM_STEP2X] = utils.turns_fixed(self._chroma_demod._fc / 4.0)
        d.phases[S.SP_FSC_DR_HALF] = utils.turns_fixed(self._fsc_dr / 2.0)
        d.phases[S.SP_FSC_DB_HALF] = utils.turns_fixed(self._fsc_db / 2.0)
        d.phases[S.SP_INVERSIONS] = sum(1 << i for i, inv in enumerate(self._start_phase_inversions) if inv)
        v = self._variant
        for slot, val in ((S.SS_FSC_DR, self._fsc_dr), (S.SS_FSC_DB, self._fsc_db), (S.SS_FDEV_DR, self._fdev_dr),
                          (S.SS_FDEV_DB, self._fdev_db), (S.SS_F_LO, self._flimit_min), (S.SS_F_HI, self._flimit_max),
                          (S.SS_BELL_F0, self._bell_f0), (S.SS_M0, v.m0), (S.SS_KN, v.bell_kn), (S.SS_KD, v.bell_kd),
                          (S.SS_FM_FC, self._chroma_demod._fc)):
            d.scalars[slot] = val
