"""NIIR / SECAM-IV modems — drop-in for ``color_modem.color.niir`` (niir.py:10-202), computed on the GPU."""
import numpy

from .. import _native as N
from .. import _slots as S
from .. import utils
from ..modem import GpuModem
from ..qam import put_filter, put_resampler
from .pal import PalVariant


class NiirModem(GpuModem, utils.ConstantFrequencyCarrier):
    kind = N.KIND_NIIR
    decoder_rows = 2
    has_demodulate_components = True
    # rows (luma, db, dr) over (r, g, b)  — niir.py:35-37
    ENC = (0.299, 0.587, 0.114,
           0.1472906403940887, 0.2891625615763547, -0.4364532019704434,
           0.6149122807017545, -0.5149122807017544, -0.1)
    # rows (r, g, b) over (luma, db, dr)  — niir.py:56-58
    DEC = (1.0, 0.0, 1.14,
           1.0, 0.3942419080068143, -0.5806814310051107,
           1.0, -2.03, 0.0)

    def __init__(self, line_config, config=PalVariant.PAL, noise_level=0.0, precision='fp32'):
        if noise_level != 0.0:
            raise NotImplementedError('noise_level dithering uses an unseeded RNG in the reference and is not built')
        GpuModem.__init__(self, line_config, precision)
        self.config = config
        fs = line_config.fs
        self._carrier_phase_step = 2.0 * numpy.pi * config.fsc / fs
        self._noise_level = 0.0
        self._demodulate_resample_factor = 3
        wc, wp, ws = 2.0 * config.fsc / fs, 2.0 * config.bandwidth3db / fs, 2.0 * config.bandwidth20db / fs
        self._chroma_precorrect_lowpass = utils.iirdesign(wp, ws, 3.0, 20.0)
        self._demodulate_upsampled_baseband_filter = utils.iirdesign(wp / 3, ws / 3, 3.0, 20.0)
        self._demodulate_upsampled_filter = utils.iirdesign_wc(wc / 3, wp / 3, ws / 3, 3.0, 20.0)

    @classmethod
    def encode_components(cls, r, g, b):
        m = numpy.asarray(cls.ENC).reshape(3, 3)
        r, g, b = (numpy.asarray(x, dtype=numpy.float64) for x in (r, g, b))
        return tuple(m[i, 0] * r + m[i, 1] * g + m[i, 2] * b for i in range(3))

    @classmethod
    def decode_components(cls, luma, db, dr):
        m = numpy.asarray(cls.DEC).reshape(3, 3)
        luma, db, dr = (numpy.asarray(x, dtype=numpy.float64) for x in (luma, db, dr))
        return tuple(m[i, 0] * luma + m[i, 1] * db + m[i, 2] * dr for i in range(3))

    def _flags(self):
        return 0

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
        put_filter(d, S.NF_PRE_LP, self._chroma_precorrect_lowpass, W, 1)
        put_filter(d, S.NF_BASE_LP, self._demodulate_upsampled_baseband_filter, 3 * W, 3)
        put_filter(d, S.NF_UP_BP, self._demodulate_upsampled_filter, 3 * W, 3)
        put_resampler(d, S.NR_UP3, 3, 1)
        put_resampler(d, S.NR_DOWN3, 1, 3)
        d.phases[S.NP_STEP1X] = utils.turns_fixed(fsc / self.line_config.fs)
        d.phases[S.NP_LINE_SHIFT] = utils.radians_fixed(self.line_shift)
        d.phases[S.NP_LUMA_ROT] = utils.radians_fixed(numpy.pi - self._demodulate_upsampled_filter.phase_shift)
        d.scalars[S.NS_INV_STEP3] = self._demodulate_resample_factor / self._carrier_phase_step


class HueCorrectingNiirModem(NiirModem):
    """Encoder averages the hue of each line with the next line of the field (niir.py:166-202)."""
    modulation_delay = 1
    encoder_lookahead = True

    def _flags(self):
        return N.FLAG_HUE_CORRECT
