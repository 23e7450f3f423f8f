"""GPU: the reference's per-line protocol (modem.modulate / modem.demodulate driven exactly like
ImageModem drives it, image.py:47-55 / 75-83) gives the same rows as the whole-frame kernels and the oracle."""
import numpy as np
import pytest

import oracle
from oracle import frame as oframe
from cases import GOLDEN_CASES, case_id
from product import BUILT_KINDS, make_modem
from color_modem_b200.synth import synth_frames_u8

pytestmark = pytest.mark.gpu

# one case per composition is enough here (the state machines are per class, not per preset)
_seen = set()
CASES = []
for _c in GOLDEN_CASES:
    key = (_c.kind, _c.chroma_avg)
    if _c.kind in BUILT_KINDS and key not in _seen and _c.width == 720:
        _seen.add(key)
        CASES.append(_c)


def drive_modulate(modem, rgb01, frame):
    """The loop of reference image.py:47-55, on float rows."""
    h = rgb01.shape[0]
    d = getattr(modem, 'modulation_delay', 0)
    out = [None] * h
    for field in range(2):
        for y in range(field, 2 * d, 2):
            modem.modulate(frame, y, rgb01[y, :, 0], rgb01[y, :, 1], rgb01[y, :, 2])
        for y in range(field, h, 2):
            iy = y + 2 * d
            while iy >= h:
                iy -= 2
            out[y] = modem.modulate(frame, y + 2 * d, rgb01[iy, :, 0], rgb01[iy, :, 1], rgb01[iy, :, 2])
    return np.stack(out)


def drive_demodulate(modem, comp, frame):
    """The loop of reference image.py:75-83, on float rows."""
    h = comp.shape[0]
    d = getattr(modem, 'demodulation_delay', 0)
    out = [None] * h
    for field in range(2):
        for y in range(field, 2 * d, 2):
            modem.demodulate(frame, y, comp[y])
        for y in range(field, h, 2):
            iy = y + 2 * d
            while iy >= h:
                iy -= 2
            out[y] = np.stack(modem.demodulate(frame, y + 2 * d, comp[iy]), axis=-1)
    return np.stack(out)


@pytest.mark.parametrize('c', CASES, ids=case_id)
def test_per_line_protocol(c, cuda_required):
    h = 12
    c = c._replace(height=h)
    rgb = synth_frames_u8(1, h, c.width, first_frame=c.frame, seed=c.seed, kind=c.content)[0]
    rgb01 = rgb / 255.0
    om = oracle.build(oracle.ModemSpec(c.kind, c.variant, c.width, h, c.standard, c.chroma_avg, c.notch, c.opt))
    m = make_modem(c, 'fp64')
    comp_ref = om.encode(c.frame, rgb01)
    comp = drive_modulate(m, rgb01, c.frame)
    assert np.abs(comp - comp_ref).max() <= 1e-9
    comp_in = oframe.composite_unlevel(oframe.to_u8(oframe.composite_level(comp_ref)) / 255.0)
    out_ref = om.decode(c.frame, comp_in)
    out = drive_demodulate(m, comp_in, c.frame)
    assert np.abs(out - out_ref).max() <= 1e-9
    # running the same frame again through the same (stateful) object gives the same answer
    out2 = drive_demodulate(m, comp_in, c.frame)
    assert np.array_equal(out, out2)


# ------------------------------------------------------------------------------------------------------------
# The component-level half of the protocol (qam.py:68-72, ntsc.py:43-49, pal.py:48-59, comb.py:40-59, niir.py:80,98):
# modulate_components / demodulate_components(..., strip_chroma) driven line by line like ImageModem drives
# modulate / demodulate, against the oracle's planes.
# ------------------------------------------------------------------------------------------------------------
def drive_components(modem, comp, frame, strip):
    h = comp.shape[0]
    d = getattr(modem, 'demodulation_delay', 0)
    out = [None] * h
    for field in range(2):
        for y in range(field, 2 * d, 2):
            modem.demodulate_components(frame, y, comp[y], strip_chroma=strip)
        for y in range(field, h, 2):
            iy = y + 2 * d
            while iy >= h:
                iy -= 2
            out[y] = np.stack(modem.demodulate_components(frame, y + 2 * d, comp[iy], strip_chroma=strip), axis=-1)
    return np.stack(out)


COMPONENT_CASES = [c for c in CASES if c.kind in ('ntsc', 'ntsc_comb', 'ntsc_3d', 'pal_s', 'pal_d', 'pal_3d', 'niir',
                                                  'niir_hue', 'scomb+pal_s', 'scomb3+ntsc')]


@pytest.mark.parametrize('c', COMPONENT_CASES, ids=case_id)
def test_component_protocol(c, cuda_required):
    h = 12
    c = c._replace(height=h)
    rgb01 = synth_frames_u8(1, h, c.width, first_frame=c.frame, seed=c.seed, kind=c.content)[0] / 255.0
    om = oracle.build(oracle.ModemSpec(c.kind, c.variant, c.width, h, c.standard, c.chroma_avg, c.notch, c.opt))
    m = make_modem(c, 'fp64')
    comp_ref = om.encode(c.frame, rgb01)

    class ViaComponents(object):                      # modulate = modulate_components o encode_components (qam.py:68-69)
        modulation_delay = getattr(m, 'modulation_delay', 0)

        @staticmethod
        def modulate(frame, line, r, g, b):
            return m.modulate_components(frame, line, *m.encode_components(r, g, b))

    comp = drive_modulate(ViaComponents, rgb01, c.frame)
    assert np.abs(comp - comp_ref).max() <= 1e-9
    comp_in = oframe.composite_unlevel(oframe.to_u8(oframe.composite_level(comp_ref)) / 255.0)
    planes_ref = np.stack(om.demodulate_planes(c.frame, comp_in), axis=-1)
    planes = drive_components(m, comp_in, c.frame, True)
    assert np.abs(planes - planes_ref).max() <= 1e-9
    rgb = np.stack(m.decode_components(planes[..., 0], planes[..., 1], planes[..., 2]), axis=-1)
    assert np.abs(rgb - om.decode(c.frame, comp_in)).max() <= 1e-9
    # strip_chroma=False: the same chroma, the composite itself as luma
    raw = drive_components(m, comp_in, c.frame, False)
    assert np.abs(raw[..., 1:] - planes_ref[..., 1:]).max() <= 1e-9
    assert np.array_equal(raw[..., 0], comp_in)


def test_not_comb_wrappable_like_the_reference(cuda_required):
    """SecamModem, ProtoSecamModem and MacModem have no demodulate_components in the reference."""
    from color_modem_b200 import comb
    from color_modem_b200.color import secam
    from color_modem_b200.line import LineConfig
    m = secam.SecamModem(LineConfig((720, 576)))
    with pytest.raises(AttributeError):
        m.demodulate_components(0, 0, np.zeros(720))
    with pytest.raises(AttributeError):
        comb.SimpleCombModem(m)
