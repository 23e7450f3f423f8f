"""819-line AM proto-SECAM — drop-in for ``color_modem.color.protosecam`` (protosecam.py:9-112), on the GPU."""
import numpy

from .. import _native as N
from .. import _slots as S
from .. import qam, utils
from ..modem import GpuModem
from ..qam import put_filter, put_resampler


class ProtoSecamVariant(qam.QamConfig):
    pass


# 819 * 10237.5 Hz; see the discussion of the 8.37 MHz figure in the reference (protosecam.py:13-25)
ProtoSecamVariant.SECAM_1957 = ProtoSecamVariant(fsc=8384512.5, bandwidth3db=800000.0, bandwidth20db=2000000.0)


class ProtoSecamModem(GpuModem, utils.ConstantFrequencyCarrier):
    kind = N.KIND_PROTOSECAM
    decoder_rows = 2
    ENC = (0.3, 0.59, 0.11,
           1.001, -0.8437, -0.1573,
           -0.336, -0.6608, 0.9968)
    DEC = (1.0, 0.6993006993006993, 0.0,
           1.0, -0.3555766267630674, -0.1664648910411622,
           1.0, 0.0, 0.8928571428571429)

    def __init__(self, line_config, variant=ProtoSecamVariant.SECAM_1957, premod_luma_filter=True, precision='fp32'):
        GpuModem.__init__(self, line_config, precision)
        self.config = variant
        self._premod_luma_filter = bool(premod_luma_filter)
        fs = line_config.fs
        self._carrier_phase_step = numpy.pi * variant.fsc / fs
        self._demodulate_resample_factor = 3
        self._chroma_precorrect_lowpass = utils.iirdesign(2.0 * variant.bandwidth3db / fs,
                                                          2.0 * variant.bandwidth20db / fs, 3.0, 20.0)
        self._extract_chroma_up, self._remove_chroma_up = utils.iirsplitter(
            2.0 * variant.fsc / (3 * fs), 2.0 * variant.bandwidth3db / (3 * fs),
            2.0 * variant.bandwidth20db / (3 * fs), 3.0, 20.0)
        post = variant.bandwidth3db if variant.fsc < variant.bandwidth20db else variant.bandwidth20db
        self._chroma_up_post_demod_filter = utils.iirdesign(
            2.0 * min(post, variant.fsc - post) / (3 * fs), 2.0 * max(post, variant.fsc - post) / (3 * fs), 3.0, 20.0)

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

    def _flags(self):
        return N.FLAG_PROTO_LUMA if self._premod_luma_filter else 0

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
        put_filter(d, S.PF_PRE_LP, self._chroma_precorrect_lowpass, W, 1)
        put_filter(d, S.PF_BP_UP, self._extract_chroma_up, 3 * W, 3)
        put_filter(d, S.PF_BS_UP, self._remove_chroma_up, 3 * W, 3)
        put_filter(d, S.PF_POST_LP, self._chroma_up_post_demod_filter, 3 * W, 3)
        put_resampler(d, S.PR_UP3, 3, 1)
        put_resampler(d, S.PR_DOWN3, 1, 3)
        d.phases[S.PP_STEP1X] = utils.turns_fixed(fsc / self.line_config.fs)
