"""CPU check of the algebra behind the two-pass comb decoders (csrc/cm_qam.cuh: k_qam_rows + k_qam_combine).

The reference combines neighbouring lines *before* demodulating them (pal.py:116-119, ntsc.py:78-79, pal.py:198-218).
The kernels demodulate every row once, at the row's own carrier phase psi_k, into the quadrature pair

    a_k = D[sin(psi_k) X_k],   b_k = D[cos(psi_k) X_k]          D = down2 . low-pass (linear)

and obtain every comb output as a fixed linear combination of the (a, b) of rows k-1, k, k+1, using only that D is
linear and that psi_{k+1} = psi_k + LS:

    D[sin(psi_k + t) X_k] = cos(t) a_k + sin(t) b_k,     D[cos(psi_k + t) X_k] = cos(t) b_k - sin(t) a_k.

This test restates ``pair_uv`` in numpy on top of the oracle's own building blocks and compares it with the oracle's
line-combining decoders (which are pinned to the reference by tests/test_oracle_golden.py) at float64 accuracy.
"""
import numpy as np
import pytest

import oracle
from oracle import dsp
from oracle.modems import _rows_prev, _col
from color_modem_b200.synth import synth_frames_u8

TOL = 1e-11


def _comp(om, frame, h, w, seed):
    rgb = synth_frames_u8(1, h, w, first_frame=frame, seed=seed)[0] / 255.0
    return om.encode(frame, rgb)


def _row_pairs_std(om, frame, comp):
    """(a, b) of every row from B = BP(up2 c) through _demod_lowpass at the row's own phase, without the factor 2."""
    _, u, v = om.qam.demod(om.start_phase(frame, om.rows), comp, False)     # u = D[2 sin(psi) B], v = D[2 cos(psi) B]
    return 0.5 * u, 0.5 * v


def _row_pairs_pald(om, frame, comp):
    """(a, b) of every row through the PAL-D chain (G = up2(extract(c)), PalDModem._filter) at psi - LS/2."""
    ph = (om.start_phase(frame, om.rows) + om.qam.bp2x.phase_shift - 0.5 * om.line_shift) % (2.0 * np.pi)
    e = om.qam.extract(comp)
    return om._am(e, ph), om._am(e, (ph + 0.5 * np.pi) % (2.0 * np.pi))


def test_pal_d_sum_difference_from_row_pairs():
    h, w, frame = 24, 720, 3
    om = oracle.build(oracle.ModemSpec('pal_d', 'PAL', w, h, 'GERBER_625'))
    comp = _comp(om, frame, h, w, 5)
    a, b = _row_pairs_pald(om, frame, comp)
    ap, bp = _rows_prev(a), _rows_prev(b)
    cl, sl = np.cos(om.line_shift), np.sin(om.line_shift)
    s = a + (cl * ap + sl * bp)                               # D[sin(psi_k) (G_k + G_{k-1})]
    d = b - (cl * bp - sl * ap)                               # D[cos(psi_k) (G_k - G_{k-1})]
    u = d * om.sin_f + s * om.cos_f                           # pal.py:121-122
    v = d * om.cos_f - s * om.sin_f
    v = np.where(_col(om.raster.is_alternate(frame, om.rows)), -v, v)
    u_ref, v_ref = om.combed_uv(frame, om.rows, _rows_prev(comp), comp)
    assert np.abs(u - u_ref)[2:].max() < TOL                  # rows 0, 1 are field tops (no predecessor)
    assert np.abs(v - v_ref)[2:].max() < TOL


@pytest.mark.parametrize('variant,std', [('NTSC', 'NTSC_525'), ('NTSC443', 'NTSC_525')])
def test_ntsc_two_and_three_line_comb_from_row_pairs(variant, std):
    h, w, frame = 24, 720, 2
    om = oracle.build(oracle.ModemSpec('ntsc_3d', variant, w, h, std))
    comp = _comp(om, frame, h, w, 6)
    a, b = _row_pairs_std(om, frame, comp)
    ap, bp = _rows_prev(a), _rows_prev(b)
    an, bn = a[om.next_row], b[om.next_row]
    ch, sh = np.cos(0.5 * om.line_shift), np.sin(0.5 * om.line_shift)
    f2 = 2.0 * om.factor

    def two_line(ac, bc, al, bl):
        # phase psi_k - LS/2 on the current row equals psi_{k-1} + LS/2 on the previous one; u from cos, v from -sin
        u = f2 * ((ch * bc + sh * ac) - (ch * bl - sh * al))
        v = -f2 * ((ch * ac - sh * bc) - (ch * al + sh * bl))
        return u, v

    u0, v0 = two_line(a, b, ap, bp)
    u_ref, v_ref = om.combed_uv(frame, om.rows, _rows_prev(comp), comp)
    assert np.abs(u0 - u_ref)[2:].max() < TOL
    assert np.abs(v0 - v_ref)[2:].max() < TOL
    # 3-line: average with the next row's 2-line chroma (field tops take the band-split chroma 2a, 2b; at the field
    # bottom the driver re-feeds the row, so the second term vanishes there)
    top = np.arange(h) < 2
    bottom = om.next_row == np.arange(h)
    u0 = np.where(top[:, None], 2.0 * a, u0)
    v0 = np.where(top[:, None], 2.0 * b, v0)
    u1, v1 = two_line(an, bn, a, b)
    u1 = np.where(bottom[:, None], 0.0, u1)
    v1 = np.where(bottom[:, None], 0.0, v1)
    _, u_ref3, v_ref3 = om.demodulate_planes(frame, comp)
    assert np.abs(0.5 * (u0 + u1) - u_ref3).max() < TOL
    assert np.abs(0.5 * (v0 + v1) - v_ref3).max() < TOL


def test_pal_three_line_comb_from_row_pairs():
    h, w, frame = 24, 720, 5
    om = oracle.build(oracle.ModemSpec('pal_3d', 'PAL', w, h, 'GERBER_625'))
    comp = _comp(om, frame, h, w, 7)
    a, b = _row_pairs_std(om, frame, comp)
    ap, bp = _rows_prev(a), _rows_prev(b)
    bottom = (om.next_row == np.arange(h))[:, None]
    an, bn = a[om.next_row], b[om.next_row]
    cl, sl = np.cos(om.line_shift), np.sin(om.line_shift)
    sin_n = np.where(bottom, a, cl * an - sl * bn)            # D[sin(psi_k) B_{k+1}]  (B_{k+1} := B_k at the bottom)
    cos_n = np.where(bottom, b, cl * bn + sl * an)
    sin_p, cos_p = cl * ap + sl * bp, cl * bp - sl * ap       # D[sin / cos(psi_k) B_{k-1}]
    a_ss, a_cu, a_cv = om.sin_sum, om.cos_u, om.cos_v         # x 0.5 (comb.avg) x 2 (qam.demodulate's factor)
    u = a_ss * (cos_n - cos_p) + a_cu * (sin_n - 2.0 * a + sin_p)
    v = a_ss * (sin_n - sin_p) + a_cv * (cos_n - 2.0 * b + cos_p)
    v = np.where(_col(om.raster.is_alternate(frame, om.rows)), -v, v)
    _, u_ref, v_ref = om.demodulate_planes(frame, comp)
    assert np.abs(u - u_ref)[2:].max() < TOL
    assert np.abs(v - v_ref)[2:].max() < TOL


def test_encoder_lowpass_commutes_with_the_combination():
    """comb.py:53 low-passes (u, v) before re-modulating; pass 1 low-passes a and b instead (alpha, beta)."""
    h, w, frame = 24, 720, 1
    om = oracle.build(oracle.ModemSpec('pal_d', 'PAL', w, h, 'GERBER_625'))
    comp = _comp(om, frame, h, w, 8)
    a, b = _row_pairs_pald(om, frame, comp)
    cl, sl = np.cos(om.line_shift), np.sin(om.line_shift)

    def combine(a, b):
        ap, bp = _rows_prev(a), _rows_prev(b)
        s = a + (cl * ap + sl * bp)
        d = b - (cl * bp - sl * ap)
        return d * om.sin_f + s * om.cos_f, d * om.cos_f - s * om.sin_f

    u, v = combine(a, b)
    ul, vl = combine(om.qam.pre_lp(a), om.qam.pre_lp(b))
    assert np.abs(ul - om.qam.pre_lp(u)).max() < TOL
    assert np.abs(vl - om.qam.pre_lp(v)).max() < TOL
