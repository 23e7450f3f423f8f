"""D2-MAC style time-compressed multiplex — drop-in for ``color_modem.color.mac`` (mac.py:9-125), on the GPU."""
import collections
import fractions

import numpy

from .. import _native as N
from .. import _slots as S
from ..modem import GpuModem
from ..qam import put_resampler

MacVariant = collections.namedtuple('MacVariant', ['width'])

MacVariant.D2MAC_12MHZ = MacVariant(1080)
MacVariant.D2MAC_7MHZ = MacVariant(720)


class MacModem(GpuModem):
    kind = N.KIND_MAC
    decoder_rows = 2
    ENC = (0.299, 0.587, 0.114,
           0.649827, -0.544149, -0.105678,
           -0.219167, -0.430271, 0.649438)
    DEC = (1.0, 1.0787486515641855, 0.0,
           1.0, -0.5494818514781797, -0.2649492993950324,
           1.0, 0.0, 1.364256480218281)

    def __init__(self, line_config, variant_or_width=MacVariant.D2MAC_12MHZ, precision='fp32'):
        GpuModem.__init__(self, line_config, precision)
        try:
            self._width = int(variant_or_width.width)
        except AttributeError:
            self._width = int(variant_or_width)
        if self._width % 4 or self._width > 1080 or self._width <= 0:
            raise NotImplementedError('MAC composite widths must be multiples of 4 and at most 1080 samples')

    @property
    def composite_width(self):
        return self._width

    @property
    def output_width(self):
        return 720              # mac.py:79-80: the decoder always produces 720 samples per line

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
        return 0

    def _fill_desc(self, d):
        d.kind, d.flags = self.kind, self._flags()
        for i in range(9):
            d.enc_matrix[i] = self.ENC[i]
            d.dec_matrix[i] = self.DEC[i]
        for slot, (num, den) in ((S.MR_LUMA_IN, (720, self.width)), (S.MR_CHROMA_IN, (360, self.width)),
                                 (S.MR_OUT, (self._width, 1080)), (S.MR_COMP_IN, (1080, self._width))):
            fr = fractions.Fraction(num, den)
            if fr.numerator != fr.denominator:          # mac.py:49-54, 71-73, 81-83: identity ratios are skipped
                put_resampler(d, slot, fr.numerator, fr.denominator)
        put_resampler(d, S.MR_UP2, 2, 1)
        d.nresamplers = 5
