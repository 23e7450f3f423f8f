"""NTSC presets and modems — drop-in for ``color_modem.color.ntsc`` (ntsc.py:8-82), computed on the GPU."""
import numpy

from .. import _native as N
from .. import _slots as S
from .. import qam


class NtscVariant(qam.QamConfig):
    def __new__(cls, fsc, bandwidth3db=1300000.0, bandwidth20db=3600000.0):
        return super(NtscVariant, cls).__new__(cls, fsc, bandwidth3db, bandwidth20db)


NtscVariant.NTSC = NtscVariant(fsc=227.5 * 15750.0 * 1000.0 / 1001.0)
NtscVariant.NTSC_A = NtscVariant(fsc=2657812.5, bandwidth3db=1000000.0, bandwidth20db=2500000.0)
NtscVariant.NTSC_I = NtscVariant(fsc=4429687.5)
NtscVariant.NTSC443 = NtscVariant(fsc=4433618.75)
NtscVariant.NTSC_N = NtscVariant(fsc=3585937.5)
NtscVariant.NTSC361 = NtscVariant(fsc=229.5 * 15750.0 * 1000.0 / 1001.0)


class NtscModem(qam.AbstractQamColorModem):
    """Band-splitting NTSC modem (ntsc.py:23-49)."""
    ENC = (0.3, 0.59, 0.11,
           -0.1476019510016258, -0.2893575108184752, 0.436959461820101,
           0.6183717846575098, -0.5185533057776567, -0.099818478879853)
    # rows (r, g, b) over columns (y, u, v); the reference lists the v term first (ntsc.py:38-40)
    DEC = (0.9999999999999998, 0.007249535771601484, 1.133735501874552,
           1.0, -0.3834753199055935, -0.5766784873222262,
           1.0, 2.037050709207452, 0.001087790524980047)

    def __init__(self, line_config, variant=NtscVariant.NTSC, precision='fp32'):
        super(NtscModem, self).__init__(line_config, variant, precision)


class NtscCombModem(NtscModem):
    """2-line comb decoder (ntsc.py:52-82 over comb.py:9-68)."""
    kind = N.KIND_NTSC_COMB
    decoder_rows = 2

    def __init__(self, line_config, variant=NtscVariant.NTSC, notch=0.0, precision='fp32'):
        super(NtscCombModem, self).__init__(line_config, variant, precision)
        self._notch_q = float(notch)
        self.backend = self
        sine = numpy.sin(self.line_shift * 0.5)
        self._factor = 0.5 / sine if abs(sine) > 0.05 else float('inf')

    def _flags(self):
        return self.flags | (0 if numpy.isfinite(self._factor) else N.FLAG_NTSC_NO_COMB)

    def _fill_desc(self, d):
        super(NtscCombModem, self)._fill_desc(d)
        d.scalars[S.QS_NTSC_FACTOR] = self._factor if numpy.isfinite(self._factor) else 0.0
