"""PAL presets and modems — drop-in for ``color_modem.color.pal`` (pal.py:9-234), computed on the GPU."""
import numpy

from .. import _native as N
from .. import _slots as S
from .. import qam, utils


class PalVariant(qam.QamConfig):
    def __new__(cls, fsc, bandwidth3db=1300000.0, bandwidth20db=4000000.0):
        return super(PalVariant, cls).__new__(cls, fsc, bandwidth3db, bandwidth20db)


PalVariant.PAL = PalVariant(fsc=4433618.75)
PalVariant.PAL_M = PalVariant(fsc=227.25 * 15750.0 * 1000.0 / 1001.0, bandwidth20db=3600000.0)
PalVariant.PAL_N = PalVariant(fsc=3582056.25, bandwidth20db=3600000.0)
# PalVariant.PAL_A (pal.py:23) is deliberately absent: its band-pass design is unstable in the reference itself
# ("OUT OF SCOPE").


class PalSModem(qam.AbstractQamColorModem):
    """Simple PAL: band-split decoder with V switch (pal.py:28-59)."""
    flags = N.FLAG_PAL_VSWITCH
    ENC = (0.299, 0.587, 0.114,
           -0.147407, -0.289391, 0.436798,
           0.614777, -0.514799, -0.099978)
    DEC = (1.0, 0.0, 1.140250855188141,
           1.0, -0.3939307027516405, -0.5808092090310976,
           1.0, 2.028397565922921, 0.0)

    def __init__(self, line_config, variant=PalVariant.PAL, precision='fp32'):
        super(PalSModem, self).__init__(line_config, variant, precision)


class PalDModem(PalSModem):
    """PAL-D: one-line delay, sum/difference decoder (pal.py:62-127 over comb.py:9-68)."""
    kind = N.KIND_PAL_D
    decoder_rows = 2

    def __init__(self, line_config, variant=PalVariant.PAL, notch=0.0, precision='fp32'):
        super(PalDModem, self).__init__(line_config, variant, precision)
        self._notch_q = float(notch)
        self.backend = self
        self._sin_factor = numpy.sin(0.5 * self.line_shift)
        self._cos_factor = numpy.cos(0.5 * self.line_shift)
        self._filter = utils.iirfilter(6, (1.0 - 1300000.0 / variant.fsc) * self.qam.carrier_phase_step / numpy.pi,
                                       rs=48.0, btype='lowpass', ftype='cheby2')

    def _fill_desc(self, d):
        super(PalDModem, self)._fill_desc(d)
        qam.put_filter(d, S.QF_PALD_LP, self._filter, 2 * self.width, 2)
        d.scalars[S.QS_PALD_SIN] = self._sin_factor
        d.scalars[S.QS_PALD_COS] = self._cos_factor


class Pal3DModem(PalDModem):
    """3-line PAL comb decoder with a one-line output delay (pal.py:130-234)."""
    kind = N.KIND_PAL_3D

    def __init__(self, line_config, variant=PalVariant.PAL, notch=0.0, use_sin=True, use_cos=True, avg=None,
                 precision='fp32'):
        super(Pal3DModem, self).__init__(line_config, variant, notch, precision)
        from .. import comb
        self._minavg = comb._avg_mode(avg)
        lssin = numpy.sin(self.line_shift)
        lscos = numpy.cos(self.line_shift)
        if abs(lssin) < 0.1:
            use_sin = False
        if abs(lscos) > 0.9:
            use_cos = False
        self._use_sin, self._use_cos = bool(use_sin), bool(use_cos)
        if not (self._use_sin and self._use_cos):
            self._minavg = False                 # pal.py:213-218: a single estimate is not combined
        self.demodulation_delay = 1 if (use_sin or use_cos) else 0
        self.decoder_rows = 3 if self.demodulation_delay else 2
        self._sin_sum_factor = 0.5 / lssin if use_sin else 0.0
        self._cos_u_factor = -0.5 / (1.0 - lscos) if use_cos else 0.0
        self._cos_v_factor = -0.5 / (1.0 + lscos) if use_cos else 0.0

    def _flags(self):
        return self.flags | (N.FLAG_PAL3D_SIN if self._use_sin else 0) | (N.FLAG_PAL3D_COS if self._use_cos else 0)

    def _fill_desc(self, d):
        super(Pal3DModem, self)._fill_desc(d)
        # comb.avg of the two estimates (pal.py:210-218) is folded into the factors; comb.minavg needs them separate
        both = 0.5 if (self._use_sin and self._use_cos and not self._minavg) else 1.0
        d.scalars[S.QS_P3D_SINSUM] = both * self._sin_sum_factor
        d.scalars[S.QS_P3D_COSU] = both * self._cos_u_factor
        d.scalars[S.QS_P3D_COSV] = both * self._cos_v_factor
