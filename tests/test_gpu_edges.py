"""Edge cases of the CUDA path (run with -m gpu): empty batches, odd and tiny heights (one row per field, fields of
unequal length, row counts that do not divide the rows-per-CTA / rows-per-segment constants of the kernels), large
absolute frame numbers (carrier phase wraps), unsupported widths.  Results are held to the same +-1 LSB bound against
the float64 oracle as the golden cases."""
import numpy as np
import pytest

import oracle
from oracle import frame as oframe
from color_modem_b200 import comb
from color_modem_b200.color import ntsc, pal, secam, niir, protosecam, mac
from color_modem_b200.line import LineConfig, LineStandard as LS
from color_modem_b200.synth import synth_frames_u8

pytestmark = pytest.mark.gpu


def _lsb(a, b):
    return int(np.abs(a.astype(np.int32) - b.astype(np.int32)).max())


MAKERS = {
    'ntsc': ('NTSC', 'NTSC_525', lambda lc: ntsc.NtscModem(lc)),
    'ntsc_comb': ('NTSC', 'NTSC_525', lambda lc: ntsc.NtscCombModem(lc)),
    'ntsc_3d': ('NTSC', 'NTSC_525', lambda lc: comb.Simple3DCombModem(ntsc.NtscCombModem(lc))),
    'pal_s': ('PAL', 'GERBER_625', lambda lc: pal.PalSModem(lc)),
    'pal_d': ('PAL', 'GERBER_625', lambda lc: pal.PalDModem(lc)),
    'pal_3d': ('PAL', 'GERBER_625', lambda lc: pal.Pal3DModem(lc)),
    'secam': ('SECAM', 'GERBER_625', lambda lc: secam.SecamModem(lc)),
    'niir': ('PAL', 'GERBER_625', lambda lc: niir.NiirModem(lc)),
    'niir_hue': ('PAL', 'GERBER_625', lambda lc: niir.HueCorrectingNiirModem(lc)),
    'protosecam': ('SECAM_1957', 'FRENCH_819', lambda lc: protosecam.ProtoSecamModem(lc)),
    'mac': ('D2MAC_12MHZ', 'GERBER_625', lambda lc: mac.MacModem(lc)),
}


def _roundtrip_vs_oracle(kind, h, w=720, frames=(0,), n=None):
    import torch
    variant, std, make = MAKERS[kind]
    m = make(LineConfig((w, h), getattr(LS, std)))
    om = oracle.build(oracle.ModemSpec(kind, variant, w, h, std))
    first = frames[0]
    n = n or len(frames)
    rgb = synth_frames_u8(n, h, w, first_frame=first, seed=11)
    comp = m.encode_frames(torch.from_numpy(rgb).cuda(), first_frame=first)
    out = m.decode_frames(comp, first_frame=first)
    comp, out = comp.cpu().numpy(), out.cpu().numpy()
    for i in range(n):
        assert _lsb(comp[i], oframe.encode_frame_u8(om, first + i, rgb[i])) <= 1, (kind, h, i)
        assert _lsb(out[i], oframe.decode_frame_u8(om, first + i, comp[i])) <= 1, (kind, h, i)


@pytest.mark.parametrize('h', [2, 3, 5, 7, 17, 34, 35])
@pytest.mark.parametrize('kind', sorted(MAKERS))
def test_odd_and_tiny_heights(kind, h, cuda_required):
    _roundtrip_vs_oracle(kind, h, n=3)


@pytest.mark.parametrize('kind', ['ntsc_3d', 'pal_d', 'pal_3d', 'secam', 'niir'])
def test_large_absolute_frame_numbers(kind, cuda_required):
    # 2^31 + 12345: beyond int32, far beyond any frame cycle of the carrier (4 frames NTSC, 8 frames PAL)
    _roundtrip_vs_oracle(kind, 24, frames=(2147483648 + 12345,), n=2)


def test_empty_batch(cuda_required):
    import torch
    m = pal.PalDModem(LineConfig((720, 24), LS.GERBER_625))
    comp = m.encode_frames(torch.empty((0, 24, 720, 3), dtype=torch.uint8, device='cuda'))
    assert tuple(comp.shape) == (0, 24, 720)
    out = m.decode_frames(comp)
    assert tuple(out.shape) == (0, 24, 720, 3)
    assert m.encode_frames_host(np.empty((0, 24, 720, 3), np.uint8)).shape == (0, 24, 720)


def test_batch_not_multiple_of_chunks(cuda_required):
    """Ragged batches: 17 frames are less than one 32-frame chunk of the host entry points, 33 and 129 one more than one / four; the
    pass-1 / pass-2 chunking of the comb decoders (frames per 2 GiB of scratch) is exercised with CM_CHUNK in
    tests/test_gpu_parity.py::test_comb_decoder_paths and below."""
    import torch
    m = comb.Simple3DCombModem(ntsc.NtscCombModem(LineConfig((720, 24), LS.NTSC_525)))
    rgb = synth_frames_u8(129, 24, 720, first_frame=0, seed=4)
    x = torch.from_numpy(rgb).cuda()
    comp = m.encode_frames(x)
    out = m.decode_frames(comp).cpu().numpy()
    comp = comp.cpu().numpy()
    for n in (17, 33, 129):
        c = m.encode_frames_host(rgb[:n], 0)
        assert np.array_equal(c, comp[:n])
        assert np.array_equal(m.decode_frames_host(c, 0), out[:n])
    one = m.decode_frames(torch.from_numpy(comp[128:129]).cuda(), first_frame=128).cpu().numpy()
    assert np.array_equal(one[0], out[128])


def test_decode_chunking_is_invisible(cuda_required, monkeypatch):
    """the same 129 frames decoded with 64-, 50- and 1-frame pass-1 / pass-2 chunks"""
    import torch
    m = comb.Simple3DCombModem(ntsc.NtscCombModem(LineConfig((720, 24), LS.NTSC_525)))
    rgb = synth_frames_u8(129, 24, 720, first_frame=0, seed=4)
    comp = m.encode_frames(torch.from_numpy(rgb).cuda())
    ref = m.decode_frames(comp).cpu().numpy()
    for c in ('64', '50', '1'):
        monkeypatch.setenv('CM_CHUNK', c)          # tuning knobs are read once per handle: a fresh modem per setting
        m2 = comb.Simple3DCombModem(ntsc.NtscCombModem(LineConfig((720, 24), LS.NTSC_525)))
        assert np.array_equal(m2.decode_frames(comp).cpu().numpy(), ref)
    # line-sequential decoders cut their pairing scratch the same way
    s1 = comb.ColorAveragingModem(secam.SecamModem(LineConfig((720, 24), LS.GERBER_625)))
    monkeypatch.delenv('CM_CHUNK')
    sref = comb.ColorAveragingModem(secam.SecamModem(LineConfig((720, 24), LS.GERBER_625)))
    c24 = sref.encode_frames(torch.from_numpy(rgb[:23]).cuda())
    want = sref.decode_frames(c24).cpu().numpy()
    monkeypatch.setenv('CM_CHUNK', '5')
    s1 = comb.ColorAveragingModem(secam.SecamModem(LineConfig((720, 24), LS.GERBER_625)))
    assert np.array_equal(s1.decode_frames(c24).cpu().numpy(), want)


@pytest.mark.parametrize('kind', ['niir', 'niir_hue'])
def test_niir_encoder_strip_length_is_invisible(kind, cuda_required, monkeypatch):
    """k_niir_encode2 walks strips of consecutive field rows and keeps the field neighbour's planes for the next iteration:
    the composite must not depend on where the strips are cut (1, 3 or 16 rows, against the launcher's own choice), and
    must stay within 1 LSB of the oracle; the averaging front end (ColorAveragingModem) takes the same path."""
    import torch
    h, w = 70, 720
    lc = LineConfig((w, h), LS.GERBER_625)
    make = {'niir': lambda: comb.ColorAveragingModem(niir.NiirModem(lc)), 'niir_hue': lambda: niir.HueCorrectingNiirModem(lc)}[kind]
    rgb = synth_frames_u8(3, h, w, first_frame=6, seed=13)
    x = torch.from_numpy(rgb).cuda()
    ref = make().encode_frames(x, first_frame=6).cpu().numpy()
    for r in ('1', '3', '16'):
        monkeypatch.setenv('CM_ROWS_MAX', r)          # read once per handle
        assert np.array_equal(make().encode_frames(x, first_frame=6).cpu().numpy(), ref), r
    monkeypatch.delenv('CM_ROWS_MAX')
    if kind == 'niir_hue':
        om = oracle.build(oracle.ModemSpec('niir_hue', 'PAL', w, h, 'GERBER_625'))
        for i in range(3):
            assert _lsb(ref[i], oframe.encode_frame_u8(om, 6 + i, rgb[i])) <= 1


def test_unsupported_width_is_a_clean_error(cuda_required):
    """Line widths must be multiples of 4 samples (the reference accepts any): refused when the modem is constructed,
    and by cm_create for callers of the C ABI."""
    import ctypes as C
    from color_modem_b200 import _native as N
    with pytest.raises(NotImplementedError):
        pal.PalDModem(LineConfig((722, 24), LS.GERBER_625))
    d = pal.PalDModem(LineConfig((720, 24), LS.GERBER_625)).describe()
    d.width = d.comp_width = d.out_width = 722
    ptr = C.c_void_p()
    assert N.load().cm_create(C.byref(d), N.FP32, C.byref(ptr)) != 0
    assert b'multiples of 4' in N.load().cm_last_error()


@pytest.mark.parametrize('kind', ['niir', 'niir_hue'])
def test_niir_grey_pixels_follow_the_reference_rounding(kind, cuda_required):
    """Grey pixels have zero chroma in exact arithmetic; the NIIR encoders nevertheless add their 0.1 saturation offset
    along the direction of the float64 rounding residue of the reference's matrix product (niir.py:35-48).  The kernel
    reproduces that residue (csrc/cm_niir.cuh: niir_chroma_f64), so flat grey areas, grey ramps and grey/colour edges
    encode to the reference's bytes."""
    import torch
    from color_modem_b200.color import niir as pniir
    h, w = 24, 720
    rgb = synth_frames_u8(2, h, w, first_frame=3, seed=21)
    rgb[:, :, 100:400, :] = rgb[:, :, 100:400, :1]                 # grey picture content (r = g = b)
    rgb[:, 4:12, 400:600, :] = 0                                   # black
    rgb[:, 12:20, 400:600, :] = 255                                # white
    rgb[0, :, 600:700, :] = (np.arange(100, dtype=np.uint8) * 2)[None, :, None]   # grey ramp
    lc = LineConfig((w, h), LS.GERBER_625)
    m = pniir.NiirModem(lc) if kind == 'niir' else pniir.HueCorrectingNiirModem(lc)
    om = oracle.build(oracle.ModemSpec(kind, 'PAL', w, h, 'GERBER_625'))
    comp = m.encode_frames(torch.from_numpy(rgb).cuda(), first_frame=3).cpu().numpy()
    for i in range(2):
        assert _lsb(comp[i], oframe.encode_frame_u8(om, 3 + i, rgb[i])) <= 1


def _special_frames(h, w):
    """black, white, flat colour, 1-pixel impulses on grey, full-swing vertical stripes, hard colour edges"""
    f = np.zeros((6, h, w, 3), np.uint8)
    f[1] = 255
    f[2] = (200, 40, 90)
    f[3] = 128
    f[3, h // 2, w // 3] = (255, 0, 0)
    f[3, h // 3, w // 2] = (0, 0, 255)
    f[4, :, ::2] = 255
    f[5, :, : w // 2] = (255, 255, 0)
    f[5, :, w // 2:] = (0, 0, 255)
    f[5, h // 2:, :] = f[5, h // 2:, ::-1]
    return f


@pytest.mark.parametrize('kind', sorted(MAKERS))
def test_flat_and_extreme_pictures(kind, cuda_required):
    """Pictures that exercise exact zeros, clipping and full-swing transients (the level map clips to [0, 1] on both
    sides of the path, image.py:7-8)."""
    import torch
    h, w = 24, 720
    variant, std, make = MAKERS[kind]
    m = make(LineConfig((w, h), getattr(LS, std)))
    om = oracle.build(oracle.ModemSpec(kind, variant, w, h, std))
    rgb = _special_frames(h, w)
    comp = m.encode_frames(torch.from_numpy(rgb).cuda(), first_frame=2)
    out = m.decode_frames(comp, first_frame=2)
    comp, out = comp.cpu().numpy(), out.cpu().numpy()
    for i in range(rgb.shape[0]):
        assert _lsb(comp[i], oframe.encode_frame_u8(om, 2 + i, rgb[i])) <= 1, (kind, 'encode', i)
        assert _lsb(out[i], oframe.decode_frame_u8(om, 2 + i, comp[i])) <= 1, (kind, 'decode', i)
