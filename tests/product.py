"""Build the product (color_modem_b200) modem composition for a test Case — mirrors tests/refload.make_modem."""
from color_modem_b200.line import LineConfig, LineStandard

BUILT_KINDS = {'ntsc', 'pal_s', 'pal_d'}


def make_modem(c, precision='fp32'):
    from color_modem_b200.color import ntsc, pal
    std = getattr(LineStandard, c.standard) if c.standard else None
    lc = LineConfig((c.width, c.height), std)
    k, v = c.kind, c.variant
    if k == 'ntsc':
        m = ntsc.NtscModem(lc, getattr(ntsc.NtscVariant, v), precision=precision)
    elif k == 'pal_s':
        m = pal.PalSModem(lc, getattr(pal.PalVariant, v), precision=precision)
    elif k == 'pal_d':
        m = pal.PalDModem(lc, getattr(pal.PalVariant, v), precision=precision)
    else:
        raise NotImplementedError(k)
    return m
