"""Stateless, frame-level float64 restatement of the reference modems (oracle; test infrastructure only).

The reference processes one scan line per call and keeps one-line memories inside stateful objects
(comb.py, pal.py:130-234, secam.py:278-304, ...), driven by ImageModem in field order with delay
priming (image.py:47-55, 75-83).  Because every memory is reset whenever ``frame`` changes or
``line != last_line + 2``, the output row ``y`` of a frame is a pure function of ``frame``, ``y`` and
the input rows ``y-2, y, y+2`` of the same field.  This module writes that function down directly,
vectorised over all rows of a frame:

    modem = build(ModemSpec(kind='pal_d', variant='PAL', width=720, height=576))
    comp  = modem.encode(frame, rgb01)      # [H, W, 3] float64 in [0,1] -> [H, Wc] composite
    rgb   = modem.decode(frame, comp)       # [H, Wc] -> [H, Wo, 3]

``kind`` names the reference composition:
    ntsc        NtscModem                      (ntsc.py:23-49)   band-split decoder
    ntsc_comb   NtscCombModem                  (ntsc.py:52-82)   2-line comb
    ntsc_3d     Simple3DCombModem(NtscCombModem)  (comb.py:71-127)  3-line comb
    pal_s       PalSModem                      (pal.py:28-59)
    pal_d       PalDModem                      (pal.py:62-127)
    pal_3d      Pal3DModem                     (pal.py:130-234)
    secam       SecamModem                     (secam.py:152-304)
    niir        NiirModem                      (niir.py:10-163)
    niir_hue    HueCorrectingNiirModem         (niir.py:166-202)
    protosecam  ProtoSecamModem                (protosecam.py:28-112)
    mac         MacModem                       (mac.py:15-125)
    scomb+K     SimpleCombModem(K)             (comb.py:71-122)  K in ntsc, pal_s, ntsc_comb, niir, niir_hue
    scomb3+K    Simple3DCombModem(K)           (comb.py:125-127)
``chroma_avg=True`` wraps the encoder in ColorAveragingModem (comb.py:130-167); ``notch=Q`` gives the comb decoders
the luma notch of comb.py:18-20.
"""
import collections
import fractions

import numpy as np

from . import dsp, presets
from .raster import Raster

ModemSpec = collections.namedtuple('ModemSpec', 'kind variant width height standard chroma_avg notch opt')
ModemSpec.__new__.__defaults__ = (None, None, False, 0.0, '')

TWO_PI = 2.0 * np.pi


def _rows_prev(x):
    """x[y-2] for y >= 2 (rows 0,1 get their own row; callers mask them as field tops)."""
    idx = np.maximum(np.arange(x.shape[0]) - 2, 0)
    idx[:2] = np.arange(min(2, x.shape[0]))
    return x[idx]


def _col(v):
    return np.asarray(v)[:, None]


def _avg(a, b):                                                                   # comb.py:9-10
    return 0.5 * (a + b)


def _minavg(a, b):                                                                # comb.py:13-15
    sign = (1.0 - np.signbit(a)) - np.signbit(b)
    return sign * np.minimum(np.abs(a), np.abs(b))


# =================================================================================================
# QAM core, qam.py:14-58
# =================================================================================================
class QamCore(object):
    def __init__(self, wc, wp, ws, gpass=3.0, gstop=20.0):
        self.step = 0.5 * np.pi * wc                                           # qam.py:15
        self.pre_lp = dsp.design_iirdesign(wp, ws, gpass, gstop)               # qam.py:16
        self.bp2x, self.bs2x = dsp.design_splitter(0.5 * wc, 0.5 * wp, 0.5 * ws, gpass, gstop)  # :17
        self.demod_lp = dsp.design_iirfilter(6, wc - 0.5 * ws, rs=48.0, btype='lowpass', ftype='cheby2')  # :18

    def chroma(self, start, u, v):
        """qam.py:20-26"""
        u = self.pre_lp(u)
        v = self.pre_lp(v)
        ph = dsp.carrier_ramp(start, 2.0 * self.step, u.shape[-1])
        return np.sin(ph) * u + np.cos(ph) * v

    def extract(self, comp):
        """qam.py:34-37"""
        return dsp.resample(self.bp2x(dsp.resample(comp, 2, 1)), 1, 2)

    def demod(self, start, comp, strip):
        """qam.py:43-58"""
        start = np.asarray(start) + self.bp2x.phase_shift
        c2 = dsp.resample(comp, 2, 1)
        ch2 = self.bp2x(c2)
        ph = dsp.carrier_ramp(start, self.step, ch2.shape[-1])
        u2 = self.demod_lp(2.0 * np.sin(ph) * ch2)
        v2 = self.demod_lp(2.0 * np.cos(ph) * ch2)
        u = dsp.resample(u2, 1, 2)
        v = dsp.resample(v2, 1, 2)
        y = comp
        if strip:
            y = dsp.resample(self.bs2x(c2), 1, 2)
        return y, u, v


class _Base(object):
    """Common plumbing: raster, row helpers, the ColorAveragingModem encoder wrapper."""
    comp_width = None      # None -> same as input width
    out_width = None

    def __init__(self, spec):
        self.spec = spec
        self.raster = Raster(spec.width, spec.height, spec.standard)
        self.rows = self.raster.rows()
        self.next_row = self.raster.next_in_field()
        self.top = self.raster.is_field_top()

    def _avg_with_next(self, a):
        """comb.py:141-152: chroma of row y averaged with the next row of the field (itself at the bottom)."""
        return 0.5 * (a[self.next_row] + a)

    def encode(self, frame, rgb):
        r, g, b = rgb[..., 0], rgb[..., 1], rgb[..., 2]
        y, c1, c2 = self.encode_components(r, g, b)
        if self.spec.chroma_avg:
            c1 = self._avg_with_next(c1)
            c2 = self._avg_with_next(c2)
        return self.modulate_components(frame, self.rows, y, c1, c2)

    def decode(self, frame, comp):
        return np.stack(self.decode_components(*self.demodulate_planes(frame, comp)), axis=-1)


# =================================================================================================
# NTSC / PAL family
# =================================================================================================
class _QamModem(_Base):
    pal = False

    def __init__(self, spec, cfg):
        super(_QamModem, self).__init__(spec)
        self.cfg = cfg
        fs = self.raster.fs
        self.qam = QamCore(2.0 * cfg.fsc / fs, 2.0 * cfg.bw3 / fs, 2.0 * cfg.bw20 / fs)   # qam.py:65-66
        self.line_shift = self.raster.line_shift(cfg.fsc)
        self.notch = None
        if getattr(spec, 'notch', 0.0):                                               # comb.py:18-20
            import scipy.signal
            b, a = scipy.signal.iirnotch(2.0 * cfg.fsc / fs, spec.notch)
            self.notch = dsp.Filt(b, a, 0.0, 'bandstop', True)

    def start_phase(self, frame, lines):
        return self.raster.start_phase(self.cfg.fsc, frame, lines)

    def _vswitch(self, frame, lines, v):
        if not self.pal:
            return v
        return np.where(_col(self.raster.is_alternate(frame, lines)), -v, v)      # pal.py:48-59

    def modulate_components(self, frame, lines, y, u, v):
        """ntsc.py:43-45 / pal.py:48-52"""
        return y + self.qam.chroma(self.start_phase(frame, lines), u, self._vswitch(frame, lines, v))

    def bandsplit(self, frame, lines, comp, strip):
        """ntsc.py:47-49 / pal.py:54-59"""
        y, u, v = self.qam.demod(self.start_phase(frame, lines), comp, strip)
        return y, u, self._vswitch(frame, lines, v)

    def remod(self, frame, lines, u, v):
        return self.modulate_components(frame, lines, np.zeros_like(u), u, v)


class Ntsc(_QamModem):
    @staticmethod
    def encode_components(r, g, b):                                               # ntsc.py:27-33
        return (0.3 * r + 0.59 * g + 0.11 * b,
                -0.1476019510016258 * r - 0.2893575108184752 * g + 0.436959461820101 * b,
                0.6183717846575098 * r - 0.5185533057776567 * g - 0.099818478879853 * b)

    @staticmethod
    def decode_components(y, u, v):                                               # ntsc.py:36-41
        return (0.9999999999999998 * y + 1.133735501874552 * v + 0.007249535771601484 * u,
                y - 0.5766784873222262 * v - 0.3834753199055935 * u,
                y + 0.001087790524980047 * v + 2.037050709207452 * u)

    def demodulate_planes(self, frame, comp):
        return self.bandsplit(frame, self.rows, comp, True)


class NtscComb(Ntsc):
    def __init__(self, spec, cfg):
        super(NtscComb, self).__init__(spec, cfg)
        sine = np.sin(self.line_shift * 0.5)                                      # ntsc.py:55-59
        self.factor = 0.5 / sine if abs(sine) > 0.05 else np.inf

    def combed_uv(self, frame, lines, last, curr):
        """ntsc.py:61-82 (u/v come back swapped from the quadrature demodulator of the line difference)."""
        if not np.isfinite(self.factor):
            _, u, v = self.bandsplit(frame, lines, curr, False)
            return u, v
        ph = self.start_phase(frame, lines) - 0.5 * self.line_shift
        ph = np.where(ph < 0.0, ph + TWO_PI, ph)
        _, v, u = self.qam.demod(ph, curr - last, False)
        return u * self.factor, v * (-self.factor)

    def comb_stage_uv(self, frame, comp):
        """(u, v) as returned by NtscCombModem.demodulate_components(strip_chroma=False), comb.py:47-59."""
        u, v = self.combed_uv(frame, self.rows, _rows_prev(comp), comp)
        _, ut, vt = self.bandsplit(frame, self.rows[:2], comp[:2], False)
        u[:2], v[:2] = ut, vt
        return u, v

    def demodulate_planes(self, frame, comp):
        u, v = self.comb_stage_uv(frame, comp)
        y = comp - self.remod(frame, self.rows, u, v)                             # comb.py:52-53
        if self.notch:
            y = self.notch(y)                                                     # comb.py:54-55 (combed rows only)
        yt, _, _ = self.bandsplit(frame, self.rows[:2], comp[:2], True)           # field top: comb.py:48-49
        y[:2] = yt
        return y, u, v


class Ntsc3D(NtscComb):
    """Simple3DCombModem(NtscCombModem): comb.py:96-113 with own_delay=1."""

    def demodulate_planes(self, frame, comp):
        u0, v0 = self.comb_stage_uv(frame, comp)
        # the delayed call: line number y+2, previous = row y, current = next row of the field (row y at the bottom)
        u1, v1 = self.combed_uv(frame, self.rows + 2, comp, comp[self.next_row])
        avg = _minavg if getattr(self.spec, 'opt', '') == 'minavg' else _avg       # comb.py:81-84, 101-102
        u = avg(u0, u1)
        v = avg(v0, v1)
        y = comp - self.remod(frame, self.rows, u, v)
        if self.notch:
            y = self.notch(y)                                                     # comb.py:108-109: every kept row
        return y, u, v


class PalS(_QamModem):
    pal = True

    @staticmethod
    def encode_components(r, g, b):                                               # pal.py:32-38
        return (0.299 * r + 0.587 * g + 0.114 * b,
                -0.147407 * r - 0.289391 * g + 0.436798 * b,
                0.614777 * r - 0.514799 * g - 0.099978 * b)

    @staticmethod
    def decode_components(y, u, v):                                               # pal.py:40-46
        return (y + 1.140250855188141 * v,
                y - 0.5808092090310976 * v - 0.3939307027516405 * u,
                y + 2.028397565922921 * u)

    def demodulate_planes(self, frame, comp):
        return self.bandsplit(frame, self.rows, comp, True)


class PalD(PalS):
    def __init__(self, spec, cfg):
        super(PalD, self).__init__(spec, cfg)
        self.sin_f = np.sin(0.5 * self.line_shift)                                # pal.py:65-66
        self.cos_f = np.cos(0.5 * self.line_shift)
        self.am_lp = dsp.design_iirfilter(6, (1.0 - 1300000.0 / cfg.fsc) * self.qam.step / np.pi,
                                          rs=48.0, btype='lowpass', ftype='cheby2')   # pal.py:67-69

    def _am(self, data, start):
        """pal.py:71-77"""
        d2 = dsp.resample(data, 2, 1)
        ph = dsp.carrier_ramp(start, self.qam.step, d2.shape[-1])
        return dsp.resample(self.am_lp(d2 * np.sin(ph)), 1, 2)

    def combed_uv(self, frame, lines, last, curr):
        """pal.py:79-127"""
        ph = (self.start_phase(frame, lines) + self.qam.bp2x.phase_shift - 0.5 * self.line_shift) % TWO_PI
        s = self._am(self.qam.extract(curr + last), ph)
        d = self._am(self.qam.extract(curr - last), (ph + 0.5 * np.pi) % TWO_PI)
        u = d * self.sin_f + s * self.cos_f
        v = d * self.cos_f - s * self.sin_f
        v = np.where(_col(self.raster.is_alternate(frame, lines)), v * -1.0, v)
        return u, v

    def _pald_planes(self, frame, comp):
        u, v = self.combed_uv(frame, self.rows, _rows_prev(comp), comp)
        y = comp - self.remod(frame, self.rows, u, v)
        if self.notch:
            y = self.notch(y)                                                     # comb.py:54-55 (combed rows only)
        yt, ut, vt = self.bandsplit(frame, self.rows[:2], comp[:2], True)
        y[:2], u[:2], v[:2] = yt, ut, vt
        return y, u, v

    def demodulate_planes(self, frame, comp):
        return self._pald_planes(frame, comp)


class Pal3D(PalD):
    def __init__(self, spec, cfg, use_sin=True, use_cos=True):
        super(Pal3D, self).__init__(spec, cfg)
        lssin = np.sin(self.line_shift)                                           # pal.py:155-161
        lscos = np.cos(self.line_shift)
        if abs(lssin) < 0.1:
            use_sin = False
        if abs(lscos) > 0.9:
            use_cos = False
        self.use_sin, self.use_cos = use_sin, use_cos
        if use_sin:
            self.sin_sum = 0.5 / lssin                                            # pal.py:168-169
        if use_cos:
            self.cos_u = -0.5 / (1.0 - lscos)                                     # pal.py:171-173
            self.cos_v = -0.5 / (1.0 + lscos)

    def demodulate_planes(self, frame, comp):
        if not (self.use_sin or self.use_cos):
            return self._pald_planes(frame, comp)                                 # pal.py:181-182
        prev = _rows_prev(comp)
        nxt = comp[self.next_row]
        curr_diff = nxt - comp                                                    # pal.py:198
        last_diff = comp - prev
        ssig = curr_diff + last_diff                                              # pal.py:203-204
        dsig = curr_diff - last_diff
        start = self.start_phase(frame, self.rows)                                # pal.py:206 (line - 2 == row)
        _, su, sv = self.qam.demod(start, ssig, False)
        _, du, dv = self.qam.demod(start, dsig, False)
        if self.use_sin and self.use_cos:                                         # pal.py:210-218
            avg = _minavg if getattr(self.spec, 'opt', '') == 'minavg' else _avg   # pal.py:146-149
            u = avg(sv * self.sin_sum, du * self.cos_u)
            v = avg(su * self.sin_sum, dv * self.cos_v)
        elif self.use_sin:
            u, v = self.sin_sum * sv, self.sin_sum * su
        else:
            u, v = self.cos_u * du, self.cos_v * dv
        v = np.where(_col(self.raster.is_alternate(frame, self.rows)), v * -1.0, v)
        # first row of each field: band-split chroma of the row itself, luma un-stripped (pal.py:191-202)
        _, ut, vt = self.bandsplit(frame, self.rows[:2], comp[:2], False)
        u[:2], v[:2] = ut, vt
        y = comp - self.remod(frame, self.rows, u, v)                             # pal.py:225-226
        if self.notch:
            y = self.notch(y)                                                     # pal.py:227-228: every kept row
        return y, u, v


# =================================================================================================
# SECAM, secam.py:127-304
# =================================================================================================
class FmDiscriminator(object):
    def __init__(self, fc, dev, rate=2):
        self.fc, self.rate = fc, rate
        self.lp = dsp.design_iirfilter(6, (2.0 * fc - dev) / rate, rs=48.0, btype='lowpass', ftype='cheby2')

    def __call__(self, data):
        """secam.py:134-149"""
        up = dsp.resample(data, self.rate, 1)
        n = up.shape[-1]
        ph = np.linspace(0.0, (n * np.pi * self.fc) / self.rate, num=n, endpoint=False)
        i = self.lp(up * np.cos(ph))
        q = self.lp(up * np.sin(ph))
        ang = np.unwrap(np.angle(i - 1.0j * q), axis=-1)
        dphi = np.diff(np.concatenate((ang[..., 0:1], ang), axis=-1), axis=-1)
        return dsp.resample(self.fc + self.rate * dphi / np.pi, 1, self.rate)


class Secam(_Base):
    def __init__(self, spec, v, alternate_phases=False):
        super(Secam, self).__init__(spec)
        fs = self.raster.fs
        self.v = v
        self.fsc_dr, self.fsc_db = 2.0 * v.fsc_dr / fs, 2.0 * v.fsc_db / fs        # secam.py:156-162
        self.fdev_dr, self.fdev_db = 2.0 * v.fdev_dr / fs, 2.0 * v.fdev_db / fs
        self.f_lo = 2.0 * (v.bell_f0 + v.flim_lo) / fs
        self.f_hi = 2.0 * (v.bell_f0 + v.flim_hi) / fs
        self.bell_f0 = 2.0 * v.bell_f0 / fs
        self.inversions = ([False, False, False, True, True, True] if alternate_phases
                           else [False, False, True, False, False, True])         # secam.py:163-166
        self.anti_bell = None
        if v.bell_kn != v.bell_kd:
            self.anti_bell = self._design_anti_bell(self.bell_f0, self.f_hi, v.bell_kn, v.bell_kd)
        self.pre_lp = dsp.design_iirdesign(wp=2.0 * 1300000.0 / fs, ws=2.0 * 3500000.0 / fs, gpass=3.0, gstop=30.0)
        self.pre_emph = self.de_emph = None
        if v.lf_k != 1.0:
            self.pre_emph, self.de_emph = self._design_lf(2.0 * v.lf_f1 / fs, v.lf_k)
        centre = 0.5 * (self.f_lo + self.f_hi)                                    # secam.py:179-187
        dev = 0.5 * (self.f_hi - self.f_lo)
        self.chroma_bp = dsp.design_iirfilter(3, [centre - dev, centre + dev], rp=0.1, btype='bandpass',
                                              ftype='cheby1')
        self.luma_bs = dsp.design_iirfilter(3, [centre - dev * np.e, centre + dev * np.e], btype='bandstop',
                                            ftype='bessel')
        self.fm = FmDiscriminator(centre, dev)

    @staticmethod
    def encode_components(r, g, b):                                               # secam.py:193-200
        return (0.299 * r + 0.587 * g + 0.114 * b,
                -1.333302 * r + 1.116474 * g + 0.216828 * b,
                -0.449995 * r - 0.883435 * g + 1.33343 * b)

    @staticmethod
    def decode_components(luma, dr, db):                                          # secam.py:203-208
        return (luma - 0.5257623554153522 * dr,
                luma + 0.2678074007993021 * dr - 0.1290417517983779 * db,
                luma + 0.6644518272425249 * db)

    @staticmethod
    def _design_lf(wc, k):
        """secam.py:210-221: first-order LF pre-emphasis and its exact inverse."""
        import scipy.signal
        fb, fa = scipy.signal.iirfilter(1, k * wc, btype='highpass', ftype='butter')
        fb[0] = (k - 1.0) * fb[0] + 1.0
        fb[1] = (k - 1.0) * fb[1] + fa[1]
        bb = np.array([1.0, fa[1]]) / fb[0]
        ba = np.array([1.0, fb[1] / fb[0]])
        return (dsp.Filt(fb, fa, k * wc, 'highpass', False), dsp.Filt(bb, ba, k * wc, 'lowpass', False))

    @staticmethod
    def _design_anti_bell(f0, f_max, kn, kd):
        """secam.py:223-238"""
        def gain_db(f):
            return 10.0 * np.log10(np.sqrt(
                (kd * kd * f0 * f0 * f0 * f0 + (1 - 2 * kd * kd) * f * f * f0 * f0 + kd * kd * f * f * f * f) / (
                    kn * kn * f0 * f0 * f0 * f0 + (1 - 2 * kn * kn) * f * f * f0 * f0 + kn * kn * f * f * f * f)))
        wp2 = f0 + 1 / 256.0
        wp1 = f0 * f0 / wp2
        ws2 = f_max
        ws1 = f0 * f0 / ws2
        return dsp.design_iirdesign([wp1, wp2], [ws1, ws2], -gain_db(wp2), -gain_db(ws2), shift=False)

    def _phase_inverted(self, frame, lines):
        """secam.py:248-256 (625-line numbers hard-coded regardless of the standard)."""
        fr = frame % 6
        lines = np.asarray(lines)
        lif = np.where(lines % 2 == 0, 23 + lines // 2, 336 + lines // 2)
        seq = (fr * 625 + lif) % 6
        return np.asarray(self.inversions)[seq] ^ (fr % 2 == 1)

    def _fm_chroma(self, start, freq):
        """secam.py:240-246"""
        big_f = freq / self.bell_f0 - self.bell_f0 / freq
        g = self.v.m0 * (1.0 + 1.0j * self.v.bell_kn * big_f) / (1.0 + 1.0j * self.v.bell_kd * big_f)
        ps = np.pi * freq
        ph = (_col(start) - ps[:, 0:1] - np.angle(g[:, 0:1]) + np.cumsum(ps, axis=-1)) % TWO_PI
        return np.real(g) * np.cos(ph) - np.imag(g) * np.sin(ph)

    def modulate_components(self, frame, lines, luma, dr, db):
        """secam.py:261-276"""
        alt = _col(self.raster.is_alternate(frame, lines))
        c = self.pre_lp(np.where(alt, db, dr))
        if self.pre_emph is not None:
            c = self.pre_emph(c)
        freq = np.where(alt, self.fsc_db + self.fdev_db * c, self.fsc_dr + self.fdev_dr * c)
        freq = np.minimum(np.maximum(freq, self.f_lo), self.f_hi)
        start = np.where(self._phase_inverted(frame, lines), np.pi, 0.0)
        return luma + self._fm_chroma(start, freq)

    def demodulate_planes(self, frame, comp):
        """secam.py:278-304"""
        n = comp.shape[-1]
        luma = self.luma_bs(comp)
        warm = np.flip(comp[:, 1:n // 40], axis=-1)
        ch = self.chroma_bp(np.concatenate((warm, comp), axis=-1))
        if self.anti_bell is not None:
            ch = self.anti_bell(ch)
        freq = self.fm(ch)[:, -n:]
        freq = np.minimum(np.maximum(freq, self.f_lo), self.f_hi)
        alt = _col(self.raster.is_alternate(frame, self.rows))
        x = np.where(alt, (freq - self.fsc_db) / self.fdev_db, (freq - self.fsc_dr) / self.fdev_dr)
        if self.de_emph is not None:
            x = self.de_emph(x)
        xp = _rows_prev(x)
        xp[:2] = 0.0                                                              # secam.py:279-280
        return luma, np.where(alt, xp, x), np.where(alt, x, xp)


# =================================================================================================
# NIIR / SECAM-IV, niir.py
# =================================================================================================
class Niir(_Base):
    hue_correcting = False

    def __init__(self, spec, cfg):
        super(Niir, self).__init__(spec)
        fs = self.raster.fs
        self.cfg = cfg
        self.step = TWO_PI * cfg.fsc / fs                                         # niir.py:12
        self.pre_lp = dsp.design_iirdesign(2.0 * cfg.bw3 / fs, 2.0 * cfg.bw20 / fs, 3.0, 20.0)
        self.rate = 3
        wc, wp, ws = 2.0 * cfg.fsc / fs, 2.0 * cfg.bw3 / fs, 2.0 * cfg.bw20 / fs
        self.base_lp = dsp.design_iirdesign(wp / 3, ws / 3, 3.0, 20.0)            # niir.py:90-93
        self.up_bp = dsp.design_band(wc / 3, wp / 3, ws / 3, 3.0, 20.0)
        self.line_shift = self.raster.line_shift(cfg.fsc)

    @staticmethod
    def encode_components(r, g, b):                                               # niir.py:29-38
        return (0.299 * r + 0.587 * g + 0.114 * b,
                0.1472906403940887 * r + 0.2891625615763547 * g - 0.4364532019704434 * b,
                0.6149122807017545 * r - 0.5149122807017544 * g - 0.1 * b)

    @staticmethod
    def decode_components(luma, db, dr):                                          # niir.py:50-59
        return (luma + 1.14 * dr,
                luma + 0.3942419080068143 * db - 0.5806814310051107 * dr,
                luma - 2.03 * db)

    @staticmethod
    def _add_offset(db, dr):                                                      # niir.py:40-48 (noise_level 0)
        sat = np.sqrt(db * db + dr * dr) + 0.1
        hue = np.arctan2(db, dr)
        return sat * np.sin(hue), sat * np.cos(hue)

    @staticmethod
    def _remove_offset(db, dr):                                                   # niir.py:61-65
        sat = np.maximum(np.sqrt(db * db + dr * dr) - 0.1, 0.0)
        hue = np.arctan2(db, dr)
        return sat * np.sin(hue), sat * np.cos(hue)

    def _carrier_mod(self, frame, lines, db, dr):
        """niir.py:67-74"""
        start = self.raster.start_phase(self.cfg.fsc, frame, lines)
        ph = dsp.carrier_ramp(start, self.step, db.shape[-1])
        alt = _col(self.raster.is_alternate(frame, lines))
        return np.where(alt, -np.sqrt(db * db + dr * dr) * np.sin(ph), db * np.sin(ph) + dr * np.cos(ph))

    def _mod_offset(self, frame, lines, luma, db, dr):
        """niir.py:83-88"""
        return luma + self._carrier_mod(frame, lines, self.pre_lp(db), self.pre_lp(dr))

    def modulate_components(self, frame, lines, luma, db, dr):
        return self._mod_offset(frame, lines, luma, *self._add_offset(db, dr))    # niir.py:80-81

    def demodulate_planes(self, frame, comp):
        """niir.py:98-163"""
        return self.planes_of_calls(frame, self.rows, comp, np.maximum(self.rows - 2, 0), self.rows < 2, True)

    def planes_of_calls(self, frame, lines, comp, prev_idx, top, strip):
        """niir.py:98-163 for a list of calls: call i has line number lines[i] and composite comp[i]; its predecessor
        (line - 2, same frame) is call prev_idx[i] unless top[i] (reset branch, niir.py:103-106)."""
        lines = np.asarray(lines)
        n = comp.shape[-1]
        rate = self.rate
        up = dsp.resample(comp, rate, 1)
        mod_up = self.up_bp(up)
        sat_up = self.base_lp(0.5 * np.pi * np.abs(mod_up))
        pm_up = mod_up / sat_up
        # previous row's phase carrier; at field top a synthetic reference carrier (niir.py:103-106)
        ones = np.ones((comp.shape[0], n))
        synth = self.up_bp(dsp.resample(self._carrier_mod(frame, lines - 2, ones, 0.0 * ones), rate, 1))
        last_pm = np.where(_col(top), synth, pm_up[prev_idx])
        alt = _col(self.raster.is_alternate(frame, lines))
        carrier = np.where(alt, pm_up, last_pm)                                   # niir.py:114-121
        huemod = np.where(alt, last_pm, pm_up)
        shift = np.where(alt, -self.line_shift, self.line_shift)
        mid = 0.5 * (carrier[:, 0:-1] + carrier[:, 1:])
        z = np.zeros((comp.shape[0], 1))
        alt_carrier = np.concatenate((z, np.diff(mid, axis=-1), z), axis=-1) * rate / self.step
        sinphi = dsp.resample(huemod * carrier, 1, rate)
        cosphi = dsp.resample(huemod * alt_carrier, 1, rate)
        norm = np.sqrt(cosphi * cosphi + sinphi * sinphi)
        cosphi = cosphi / norm
        sinphi = sinphi / norm
        sinphi, cosphi = (-cosphi * np.sin(shift) - sinphi * np.cos(shift),
                          sinphi * np.sin(shift) - cosphi * np.cos(shift))       # niir.py:136-137
        sat = dsp.resample(sat_up, 1, rate)
        db = sat * sinphi
        dr = sat * cosphi
        luma = comp
        if strip:
            sincar = dsp.resample(carrier, 1, rate)                               # niir.py:143-158
            coscar = dsp.resample(alt_carrier, 1, rate)
            psh = np.where(alt, 0.0, self.line_shift) + (np.pi - self.up_bp.phase_shift)
            us = np.where(alt, -np.sqrt(db * db + dr * dr), db)
            vs = np.where(alt, 0.0, dr)
            us, vs = us * np.cos(psh) - vs * np.sin(psh), us * np.sin(psh) + vs * np.cos(psh)
            luma = comp - (us * sincar + vs * coscar)
        db, dr = self._remove_offset(db, dr)                                      # niir.py:98-100
        return luma, db, dr


class NiirHue(Niir):
    """HueCorrectingNiirModem encoder, niir.py:166-202 (decoder inherited)."""

    def encode(self, frame, rgb):
        luma, db, dr = self.encode_components(rgb[..., 0], rgb[..., 1], rgb[..., 2])
        if self.spec.chroma_avg:
            raise NotImplementedError('ColorAveragingModem(HueCorrectingNiirModem) is not a reference composition')
        ndb, ndr = db[self.next_row], dr[self.next_row]
        sat_y = np.sqrt(db * db + dr * dr)
        sat_n = np.sqrt(ndb * ndb + ndr * ndr)
        div = sat_y + sat_n
        div[np.equal(div, 0.0)] = 1.0
        adb = (db * sat_y + ndb * sat_n) / div
        adr = (dr * sat_y + ndr * sat_n) / div
        mag = sat_y + 0.1
        hue = np.arctan2(adb, adr)
        return self._mod_offset(frame, self.rows, luma, mag * np.sin(hue), mag * np.cos(hue))


# =================================================================================================
# 819-line AM proto-SECAM, protosecam.py
# =================================================================================================
class ProtoSecam(_Base):
    def __init__(self, spec, cfg, premod_luma_filter=True):
        super(ProtoSecam, self).__init__(spec)
        fs = self.raster.fs
        self.cfg = cfg
        self.premod = premod_luma_filter
        self.step = np.pi * cfg.fsc / fs                                          # protosecam.py:33
        self.pre_lp = dsp.design_iirdesign(2.0 * cfg.bw3 / fs, 2.0 * cfg.bw20 / fs, 3.0, 20.0)
        self.rate = 3
        self.bp_up, self.bs_up = dsp.design_splitter(2.0 * cfg.fsc / (3 * fs), 2.0 * cfg.bw3 / (3 * fs),
                                                     2.0 * cfg.bw20 / (3 * fs), 3.0, 20.0)
        post = cfg.bw3 if cfg.fsc < cfg.bw20 else cfg.bw20                        # protosecam.py:42-50
        self.post_lp = dsp.design_iirdesign(2.0 * min(post, cfg.fsc - post) / (3 * fs),
                                            2.0 * max(post, cfg.fsc - post) / (3 * fs), 3.0, 20.0)

    @staticmethod
    def encode_components(r, g, b):                                               # protosecam.py:56-61
        return (0.3 * r + 0.59 * g + 0.11 * b,
                1.001 * r - 0.8437 * g - 0.1573 * b,
                -0.336 * r - 0.6608 * g + 0.9968 * b)

    @staticmethod
    def decode_components(luma, dr, db):                                          # protosecam.py:64-69
        return (luma + 0.6993006993006993 * dr,
                luma - 0.3555766267630674 * dr - 0.1664648910411622 * db,
                luma + 0.8928571428571429 * db)

    def modulate_components(self, frame, lines, luma, dr, db):
        """protosecam.py:74-90"""
        alt = _col(self.raster.is_alternate(frame, lines))
        c = 0.125 * (1.0 + self.pre_lp(np.where(alt, db, dr)))
        if self.premod:
            luma = dsp.resample(self.bs_up(dsp.resample(luma, self.rate, 1)), 1, self.rate)
        start = self.raster.start_phase(self.cfg.fsc, frame, lines)
        ph = dsp.carrier_ramp(start, 2.0 * self.step, c.shape[-1])
        return luma + np.cos(ph) * c

    def demodulate_planes(self, frame, comp):
        """protosecam.py:92-112"""
        up = dsp.resample(comp, self.rate, 1)
        c_up = self.post_lp(0.5 * np.pi * np.abs(self.bp_up(up)))
        luma = dsp.resample(self.bs_up(up), 1, self.rate)
        x = 8.0 * dsp.resample(c_up, 1, self.rate) - 1.0
        alt = _col(self.raster.is_alternate(frame, self.rows))
        xp = _rows_prev(x)
        xp[:2] = 0.0
        return luma, np.where(alt, xp, x), np.where(alt, x, xp)


# =================================================================================================
# D2-MAC, mac.py
# =================================================================================================
class Mac(_Base):
    def __init__(self, spec, width):
        super(Mac, self).__init__(spec)
        self.comp_width = int(width)
        self.out_width = 720

    @staticmethod
    def encode_components(r, g, b):                                               # mac.py:26-31
        return (0.299 * r + 0.587 * g + 0.114 * b,
                0.649827 * r - 0.544149 * g - 0.105678 * b,
                -0.219167 * r - 0.430271 * g + 0.649438 * b)

    @staticmethod
    def decode_components(luma, dr, db):                                          # mac.py:34-39
        return (luma + 1.0787486515641855 * dr,
                luma - 0.5494818514781797 * dr - 0.2649492993950324 * db,
                luma + 1.364256480218281 * db)

    @staticmethod
    def _fit(x, target):
        fr = fractions.Fraction(target, x.shape[-1])
        if fr.numerator != fr.denominator:
            return dsp.resample(x, fr.numerator, fr.denominator)
        return np.array(x)

    def modulate_components(self, frame, lines, luma, dr, db):
        """mac.py:41-74"""
        alt = _col(self.raster.is_alternate(frame, lines))
        luma = self._fit(luma, 720)
        ch = self._fit(np.where(alt, db, dr), 360) + 0.5
        out = 0.5 * np.ones((luma.shape[0], 1080))
        out[:, 15] = 0.4375 + 0.125 * ch[:, 2]
        out[:, 16] = 0.25 + 0.5 * ch[:, 3]
        out[:, 17] = 0.0625 + 0.875 * ch[:, 4]
        out[:, 18:369] = ch[:, 5:356]
        out[:, 369] = 0.875 * ch[:, 356] + 0.125 * luma[:, 8]
        out[:, 370] = 0.5 * ch[:, 357] + 0.5 * luma[:, 9]
        out[:, 371] = 0.125 * ch[:, 358] + 0.875 * luma[:, 10]
        out[:, 372:1071] = luma[:, 11:710]
        out[:, 1071] = 0.0625 + 0.875 * luma[:, 710]
        out[:, 1072] = 0.25 + 0.5 * luma[:, 711]
        out[:, 1073] = 0.4375 + 0.125 * luma[:, 712]
        return self._fit(out, self.comp_width)

    def demodulate_planes(self, frame, comp):
        """mac.py:77-122"""
        c = self._fit(comp, 1080)
        nrow = c.shape[0]
        luma = 0.5 * np.ones((nrow, 720))
        ch = 0.5 * np.ones((nrow, 360))
        luma[:, 11:710] = c[:, 372:1071]
        luma[:, 710] = (c[:, 1071] - 0.0625) / 0.875
        luma[:, 711] = 2.0 * c[:, 1072] - 0.5
        luma[:, 712] = 8.0 * c[:, 1073] - 3.5
        ch[:, 5:356] = c[:, 18:369]
        ch[:, 2] = 8.0 * c[:, 15] - 3.5
        ch[:, 3] = 2.0 * c[:, 16] - 0.5
        ch[:, 4] = (c[:, 17] - 0.0625) / 0.875
        luma[:, 8] = 8.0 * c[:, 369] - 7.0 * ch[:, 355]
        luma[:, 9] = 2.0 * c[:, 370] - ch[:, 355]
        luma[:, 10] = (c[:, 371] - 0.125 * ch[:, 355]) / 0.875
        luma[:, 0:8] = luma[:, 8:9]
        luma[:, 713:] = luma[:, 712:713]
        ch[:, 0] = ch[:, 2]                       # mac.py:105 writes element 0 only; element 1 stays 0.5
        ch[:, 356] = (c[:, 369] - 0.125 * luma[:, 11]) / 0.875
        ch[:, 357] = 2.0 * c[:, 370] - luma[:, 11]
        ch[:, 358] = 8.0 * c[:, 371] - 7.0 * luma[:, 11]
        ch[:, 359] = ch[:, 358]
        x = dsp.resample(ch, 2, 1) - 0.5
        alt = _col(self.raster.is_alternate(frame, self.rows))
        xp = _rows_prev(x)
        xp[:2] = 0.0
        return luma, np.where(alt, xp, x), np.where(alt, x, xp)


# =================================================================================================
# SimpleCombModem / Simple3DCombModem over any backend with demodulate_components, comb.py:71-127
# =================================================================================================
class SimpleComb(_Base):
    """comb.py:71-122, written down for the call sequence of ImageModem (image.py:75-83): per field the backend is called
    for rows y = field, field+2, ... with line number y (delay=False), or - one call ahead - for rows y+2 with line
    number y+2, the last row of the field being fed a second time with the line number after it (delay=True)."""

    def __init__(self, spec, inner, delay):
        super(SimpleComb, self).__init__(spec)
        self.inner = inner
        self.delay = bool(delay)
        self.avg = _minavg if 'minavg' in (getattr(spec, 'opt', '') or '') else _avg
        self.notch = None
        if getattr(spec, 'notch', 0.0):                                               # comb.py:18-20, 86-88
            import scipy.signal
            b, a = scipy.signal.iirnotch(2.0 * inner.cfg.fsc / self.raster.fs, spec.notch)
            self.notch = dsp.Filt(b, a, 0.0, 'bandstop', True)
        self.encode_components = inner.encode_components
        self.decode_components = inner.decode_components

    def encode(self, frame, rgb):
        return self.inner.encode(frame, rgb)                                          # comb.py:89-93

    def _calls(self, comp):
        """(lines, rows fed, predecessor call, reset flag) of the backend calls of one frame."""
        H = self.raster.height
        lines, src, prev, top = list(range(H)), list(range(H)), [max(r - 2, 0) for r in range(H)], [r < 2 for r in range(H)]
        if self.delay:
            for f in range(min(2, H)):
                last = H - 1 - ((H - 1 - f) % 2)          # last row of field f
                lines.append(last + 2)
                src.append(last)
                prev.append(last)
                top.append(False)
        return np.array(lines), np.array(src), np.array(prev), np.array(top)

    def _backend_uv(self, frame, comp):
        """(u, v) of backend.demodulate_components(..., strip_chroma=False) for every call (comb.py:98,100)."""
        lines, src, prev, top = self._calls(comp)
        inner, c = self.inner, comp[src]
        if isinstance(inner, NtscComb):
            u, v = inner.combed_uv(frame, lines, c[prev], c)
            _, ut, vt = inner.bandsplit(frame, lines[top], c[top], False)
            u[top], v[top] = ut, vt
            return u, v
        if isinstance(inner, Niir):
            _, u, v = inner.planes_of_calls(frame, lines, c, prev, top, False)
            return u, v
        if type(inner) in (Ntsc, PalS):
            _, u, v = inner.bandsplit(frame, lines, c, False)
            return u, v
        raise NotImplementedError('oracle: SimpleCombModem over %s' % type(inner).__name__)

    def _remod(self, frame, rows, u, v):
        """backend.modulate_components(frame, line - 2 * (own_delay - modulation_delay), 0, u, v) for the rows that take
        the averaging branch, in call order per field (comb.py:104-107)."""
        inner = self.inner
        if not isinstance(inner, NiirHue):
            return inner.modulate_components(frame, rows, np.zeros_like(u), u, v)
        # HueCorrectingNiirModem.modulate_components is stateful (niir.py:179-202): every call answers for the previous
        # call's line, with the magnitude of the previous call's chroma and the hue of the saturation-weighted mean of
        # the two; the first call of a field is paired with itself.  The line number passed is row + 2 and the encoder
        # emits for line - 2 = row.
        out = np.empty_like(u)
        for f in range(2):
            idx = np.nonzero(rows % 2 == f)[0]
            if not len(idx):
                continue
            pidx = np.concatenate((idx[:1], idx[:-1]))
            ldb, ldr, db, dr = u[pidx], v[pidx], u[idx], v[idx]
            last_sat = np.sqrt(ldb * ldb + ldr * ldr)
            sat = np.sqrt(db * db + dr * dr)
            div = last_sat + sat
            div[np.equal(div, 0.0)] = 1.0
            adb = (ldb * last_sat + db * sat) / div
            adr = (ldr * last_sat + dr * sat) / div
            mag = last_sat + 0.1
            hue = np.arctan2(adb, adr)
            out[idx] = inner._mod_offset(frame, rows[idx], np.zeros_like(db), mag * np.sin(hue), mag * np.cos(hue))
        return out

    def demodulate_planes(self, frame, comp):
        H = self.raster.height
        cu, cv = self._backend_uv(frame, comp)
        rows = self.rows
        if self.delay:
            # output row y: last = call of row y, curr = the call after it in the field (the re-fed row at the bottom)
            nxt = np.where(rows + 2 < H, rows + 2, H + (rows % 2))
            u, v = self.avg(cu[rows], cu[nxt]), self.avg(cv[rows], cv[nxt])
            y = comp - self._remod(frame, rows, u, v)
            if self.notch:
                y = self.notch(y)
            return y, u, v
        prev = np.maximum(rows - 2, 0)
        u, v = self.avg(cu[prev], cu[rows]), self.avg(cv[prev], cv[rows])
        y = np.array(comp)
        body = rows >= 2
        if np.any(body):
            yb = comp[body] - self._remod(frame, rows[body], u[body], v[body])
            if self.notch:
                yb = self.notch(yb)
            y[body] = yb
        u[:2], v[:2] = cu[:2], cv[:2]                     # comb.py:97-99: the first row of a field, luma un-stripped
        return y, u, v


# =================================================================================================
def build(spec):
    kind = spec.kind
    opt = getattr(spec, 'opt', '') or ''      # non-default constructor knobs, see tests/cases.py
    if kind.startswith('scomb+') or kind.startswith('scomb3+'):     # SimpleCombModem / Simple3DCombModem over a backend
        head, inner_kind = kind.split('+', 1)
        inner = build(spec._replace(kind=inner_kind, notch=0.0, opt=''))
        return SimpleComb(spec, inner, head == 'scomb3')
    if kind in ('ntsc', 'ntsc_comb', 'ntsc_3d'):
        cls = {'ntsc': Ntsc, 'ntsc_comb': NtscComb, 'ntsc_3d': Ntsc3D}[kind]
        return cls(spec, presets.NTSC[spec.variant or 'NTSC'])
    if kind in ('pal_s', 'pal_d', 'pal_3d'):
        cls = {'pal_s': PalS, 'pal_d': PalD, 'pal_3d': Pal3D}[kind]
        if kind == 'pal_3d':
            return Pal3D(spec, presets.PAL[spec.variant or 'PAL'], use_sin=(opt != 'nosin'), use_cos=(opt != 'nocos'))
        return cls(spec, presets.PAL[spec.variant or 'PAL'])
    if kind == 'secam':
        return Secam(spec, presets.SECAM[spec.variant or 'SECAM'], alternate_phases=(opt == 'altph'))
    if kind in ('niir', 'niir_hue'):
        return (Niir if kind == 'niir' else NiirHue)(spec, presets.PAL[spec.variant or 'PAL'])
    if kind == 'protosecam':
        return ProtoSecam(spec, presets.PROTOSECAM[spec.variant or 'SECAM_1957'], premod_luma_filter=(opt != 'noluma'))
    if kind == 'mac':
        v = spec.variant or 'D2MAC_12MHZ'
        return Mac(spec, presets.MAC[v] if isinstance(v, str) else int(v))
    raise ValueError('unknown modem kind %r' % (kind,))
