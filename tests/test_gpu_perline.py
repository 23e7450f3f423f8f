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
