"""Build the product (color_modem_b200) modem composition for a test Case — mirrors tests/refload.make_modem."""
from color_modem_b200.line import LineConfig, LineStandard

BUILT_KINDS = {'ntsc', 'ntsc_comb', 'ntsc_3d', 'pal_s', 'pal_d', 'pal_3d', 'secam', 'niir', 'niir_hue', 'protosecam',
               'mac', 'scomb+niir_hue', 'scomb+niir', 'scomb3+niir', 'scomb+pal_s', 'scomb3+pal_s', 'scomb+ntsc',
               'scomb3+ntsc', 'scomb+ntsc_comb'}


def make_modem(c, precision='fp32'):
    from color_modem_b200.color import ntsc, pal, secam, niir, protosecam, mac
    from color_modem_b200 import comb
    if not hasattr(ntsc.NtscVariant, 'NTSC_NOCOMB'):      # see oracle/presets.py
        ntsc.NtscVariant.NTSC_NOCOMB = ntsc.NtscVariant(fsc=227.0 * 15750.0 * 1000.0 / 1001.0)
    std = getattr(LineStandard, c.standard) if c.standard else None
    lc = LineConfig((c.width, c.height), std)
    k, v = c.kind, c.variant
    notch = getattr(c, 'notch', 0.0)
    opt = getattr(c, 'opt', '')
    if k.startswith('scomb'):
        head, inner = k.split('+', 1)
        backend = make_modem(c._replace(kind=inner, notch=0.0, opt='', chroma_avg=False), precision)
        m = comb.SimpleCombModem(backend, notch, comb.minavg if opt == 'minavg' else None, head == 'scomb3')
    elif k == 'ntsc':
        m = ntsc.NtscModem(lc, getattr(ntsc.NtscVariant, v), precision=precision)
    elif k == 'ntsc_comb':
        m = ntsc.NtscCombModem(lc, getattr(ntsc.NtscVariant, v), notch, precision=precision)
    elif k == 'ntsc_3d':
        m = comb.Simple3DCombModem(ntsc.NtscCombModem(lc, getattr(ntsc.NtscVariant, v), precision=precision), notch,
                                   comb.minavg if opt == 'minavg' else None)
    elif k == 'pal_3d':
        m = pal.Pal3DModem(lc, getattr(pal.PalVariant, v), notch, use_sin=(opt != 'nosin'), use_cos=(opt != 'nocos'),
                           avg=comb.minavg if opt == 'minavg' else None, precision=precision)
    elif k == 'secam':
        m = secam.SecamModem(lc, getattr(secam.SecamVariant, v), alternate_phases=(opt == 'altph'), precision=precision)
    elif k == 'niir':
        m = niir.NiirModem(lc, getattr(pal.PalVariant, v), precision=precision)
    elif k == 'niir_hue':
        m = niir.HueCorrectingNiirModem(lc, getattr(pal.PalVariant, v), precision=precision)
    elif k == 'protosecam':
        m = protosecam.ProtoSecamModem(lc, getattr(protosecam.ProtoSecamVariant, v), premod_luma_filter=(opt != 'noluma'), precision=precision)
    elif k == 'mac':
        m = mac.MacModem(lc, getattr(mac.MacVariant, v), precision=precision)
    elif k == 'pal_s':
        m = pal.PalSModem(lc, getattr(pal.PalVariant, v), precision=precision)
    elif k == 'pal_d':
        m = pal.PalDModem(lc, getattr(pal.PalVariant, v), notch, precision=precision)
    else:
        raise NotImplementedError(k)
    if c.chroma_avg:
        m = comb.ColorAveragingModem(m)
    return m
