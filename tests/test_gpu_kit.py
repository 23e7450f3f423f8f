"""GPU: the reference's L0 kit called directly — FilterFunction.__call__ (utils.py:28-36) and QamColorModem.modulate /
demodulate / extract_chroma with an explicit start phase (qam.py:20-58) — against the float64 oracle."""
import numpy as np
import pytest

from oracle import dsp
from oracle.modems import QamCore
from color_modem_b200 import qam, utils

pytestmark = pytest.mark.gpu

FS = 13.5e6
WC, WP, WS = 2.0 * 4433618.75 / FS, 2.0 * 1.3e6 / FS, 2.0 * 4.0e6 / FS


def _signal(n, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(n)
    return 0.4 + 0.3 * np.sin(t / 17.0 + seed) + 0.05 * rng.standard_normal(n)


@pytest.mark.parametrize('n', [720, 724, 1920, 37, 5])
@pytest.mark.parametrize('precision,tol', [('fp64', 1e-10), ('fp32', 2e-5)])
def test_filter_function_call(n, precision, tol, cuda_required):
    x = np.stack([_signal(n, s) for s in range(3)])
    for ours, ref in ((utils.iirdesign(WP, WS, 3.0, 20.0), dsp.design_iirdesign(WP, WS, 3.0, 20.0)),
                      (utils.iirfilter(6, WC - 0.5 * WS, rs=48.0, btype='lowpass', ftype='cheby2'),
                       dsp.design_iirfilter(6, WC - 0.5 * WS, rs=48.0, btype='lowpass', ftype='cheby2')),
                      (utils.iirdesign_wc(WC, WP, WS, 3.0, 20.0), dsp.design_band(WC, WP, WS, 3.0, 20.0))):
        assert ours.shift == ref.shift
        assert abs(ours.phase_shift - ref.phase_shift) < 1e-12
        got = ours(x, precision=precision)
        assert got.shape == x.shape
        assert np.abs(got - ref(x)).max() <= tol
        assert np.abs(ours(x[1], precision=precision) - ref(x[1])).max() <= tol


@pytest.mark.parametrize('n', [720, 1920])
def test_qam_color_modem_calls(n, cuda_required):
    fs = FS * n / 720.0
    wc, wp, ws = 2.0 * 4433618.75 / fs, 2.0 * 1.3e6 / fs, 2.0 * 4.0e6 / fs
    q = qam.QamColorModem(wc, wp, ws, 3.0, 20.0)
    core = QamCore(wc, wp, ws)
    y, u, v = _signal(n, 1), 0.2 * (_signal(n, 2) - 0.4), 0.2 * (_signal(n, 3) - 0.4)
    for start in (0.0, 1.2345, 5.9, -0.7):
        comp_ref = y + core.chroma(np.array([start]), u[None], v[None])[0]
        comp = q.modulate(start, y, u, v)
        assert np.abs(comp - comp_ref).max() <= 1e-10
        assert np.abs(q._modulate_chroma(start, u, v) - (comp_ref - y)).max() <= 1e-10
        for strip in (True, False):
            ref = core.demod(np.array([start]), comp_ref[None], strip)
            got = q.demodulate(start, comp_ref, strip_chroma=strip)
            for a, b in zip(got, ref):
                assert np.abs(a - b[0]).max() <= 1e-10
    assert np.abs(q.extract_chroma(comp_ref) - core.extract(comp_ref[None])[0]).max() <= 1e-10
    assert abs(q.extract_chroma_phase_shift - core.bp2x.phase_shift) < 1e-12
