"""CPU: host-side logic of the product — raster geometry, carrier phase, filter design records and descriptors —
against the oracle (which is pinned to the reference)."""
import numpy as np
import pytest

import oracle
from oracle import dsp
from cases import GOLDEN_CASES, case_id
from product import make_modem
from color_modem_b200 import utils
from color_modem_b200.line import LineConfig, LineStandard
from color_modem_b200.shard import frame_range


def test_line_standard_detect_matches_reference_rule():
    assert LineStandard.detect(480) is LineStandard.NTSC_525
    assert LineStandard.detect(576) is LineStandard.GERBER_625
    assert LineStandard.detect(376) is LineStandard.BAIRD_405
    assert LineStandard.detect(738) is LineStandard.FRENCH_819
    assert LineStandard.detect(760) is LineStandard.BELGIAN_819
    with pytest.raises(IndexError):
        LineStandard.detect(1080)           # reference line.py:28-39 raises for anything above 760 lines


@pytest.mark.parametrize('size,std', [((720, 576), None), ((720, 480), None), ((1920, 1080), 'GERBER_625'),
                                      ((720, 24), 'FRENCH_819'), ((720, 31), 'NTSC_525')])
def test_raster_matches_oracle(size, std):
    lc = LineConfig(size, getattr(LineStandard, std) if std else None)
    r = oracle.Raster(size[0], size[1], std)
    assert lc.fs == r.fs
    for y in range(-4, size[1] + 4):
        assert lc.analog_line(y) == int(r.analog_line(y))
        for frame in (0, 1, 7):
            assert lc.is_alternate_line(frame, y) == bool(r.is_alternate(frame, y))


def test_fixed_point_phase_matches_reference_formula():
    """The kernels' integer phase ((frame % cycle) * FS + (analog - ref) * LS mod 2^64) equals utils.py:82-88."""
    from color_modem_b200.color.pal import PalSModem
    lc = LineConfig((720, 576))
    m = PalSModem(lc)
    d = m.describe()
    for frame in (0, 1, 2, 3, 5, 1001):
        for line in (-2, -1, 0, 1, 2, 287, 574, 575, 577):
            adj = line + d.digital_shift
            analog = (d.even_first if adj % 2 == 0 else d.odd_first) + adj // 2
            fix = ((frame % d.frame_cycle) * d.frame_shift_turns + (analog - d.ref_line) * d.line_shift_turns) % 2 ** 64
            ref = m.start_phase(frame, line)
            diff = (fix / 2.0 ** 64 * 2 * np.pi - ref + np.pi) % (2 * np.pi) - np.pi
            assert abs(diff) < 1e-9


@pytest.mark.parametrize('c', GOLDEN_CASES, ids=case_id)
def test_descriptor_builds_and_filters_match_oracle(c):
    """Every composition yields a descriptor on a CPU box, and its SOS cascades reproduce the oracle's (b, a)."""
    import scipy.signal
    m = make_modem(c)
    om = oracle.build(oracle.ModemSpec(c.kind, c.variant, c.width, c.height, c.standard, c.chroma_avg, c.notch, c.opt))
    if c.kind.startswith('scomb'):          # composed wrapper: the kernels belong to the backend
        om = om.inner
    d = m.describe()
    assert d.width == c.width and d.height == c.height
    filts = [v for obj in (om, getattr(om, 'qam', None), getattr(om, 'fm', None)) if obj is not None
             for v in vars(obj).values() if isinstance(v, dsp.Filt)]
    x = np.random.default_rng(0).standard_normal(400)
    used = [d.filters[i] for i in range(d.nfilters) if d.filters[i].nsec > 0]
    assert len(used) == len(filts) or c.kind.split('+')[-1] in ('pal_s', 'pal_3d', 'ntsc', 'ntsc_comb', 'ntsc_3d', 'mac')
    for f in used:
        sos = np.array([[f.sos[s][0], f.sos[s][1], f.sos[s][2], 1.0, f.sos[s][3], f.sos[s][4]] for s in range(f.nsec)])
        y = scipy.signal.sosfilt(sos, x)
        best = min(np.abs(scipy.signal.lfilter(o.b, o.a, x) - y).max() for o in filts if o.shift == f.shift)
        assert best < 1e-9


def test_resampler_taps_match_scipy():
    for up, down in ((2, 1), (1, 2), (3, 1), (1, 3), (3, 8), (3, 16)):
        taps, half, u, dn = utils.resampler_taps(up, down)
        ref, rhalf, ru, rd = dsp.resample_taps(up, down)
        assert (half, u, dn) == (rhalf, ru, rd) and np.array_equal(taps, ref)


def test_frame_ranges_partition():
    for total in (1, 7, 600, 1000):
        for world in (1, 2, 3, 4, 8):
            got = [frame_range(total, r, world) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == total
            assert all(got[i][1] == got[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in got]
            assert max(sizes) - min(sizes) <= 1
