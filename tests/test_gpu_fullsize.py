"""GPU parity at BASELINE.json's full sizes (run with -m gpu on the B200).

The oracle needs 0.2-3 s per full frame, so whole sequences are checked through size-independent properties and a
few sampled frames are compared with the oracle directly:

  * sharding invariance (configs[2]): a 600-frame NTSC 3-line-comb sequence encoded and decoded in one call equals,
    byte for byte, the same sequence processed as 1/2/4/8 contiguous frame ranges that are each given their absolute
    first frame;
  * batch invariance (configs[3]): 1000 frames of ColorAveraging(SECAM) / HueCorrectingNiir through the chunked host
    entry points equal the device-resident batch;
  * sampled frames (first, shard boundaries, last) against the float64 oracle: +-1 LSB;
  * encode -> decode round trip of smooth content stays close to the source picture (a property of the modem chain
    that does not depend on the frame count);
  * configs[4]: three 1920x1080 presets, 96-frame batches (the 500-frame batch of BASELINE.json is 3 GB of input per
    preset; the property tested — independence of batch size and position — is the same), one frame each vs the oracle.
"""
import numpy as np
import pytest

import oracle
from oracle import frame as oframe
from color_modem_b200 import comb, shard
from color_modem_b200.color import ntsc, pal, secam, niir
from color_modem_b200.line import LineConfig, LineStandard as LS
from color_modem_b200.synth import synth_frames_u8

pytestmark = pytest.mark.gpu


def _lsb(a, b):
    return int(np.abs(a.astype(np.int32) - b.astype(np.int32)).max())


def _frames(n, h, w, first=0, seed=5):
    """n frames that depend on the absolute frame number: 24 distinct pictures, cycled (content generation is host-side
    numpy and would dominate the test otherwise)."""
    base = synth_frames_u8(24, h, w, first_frame=0, seed=seed)
    idx = (np.arange(n) + first) % 24
    return base[idx]


def _run_sharded(m, rgb, world):
    import torch
    n = rgb.shape[0]
    comps, outs = [], []
    for r in range(world):
        lo, hi = shard.frame_range(n, r, world)
        x = torch.from_numpy(rgb[lo:hi]).cuda()
        c = m.encode_frames(x, first_frame=lo)
        comps.append(c.cpu().numpy())
        outs.append(m.decode_frames(c, first_frame=lo).cpu().numpy())
    return np.concatenate(comps), np.concatenate(outs)


def test_ntsc3d_600_frames_sharding_invariance(cuda_required):
    n, h, w = 600, 480, 720
    rgb = _frames(n, h, w)
    m = comb.Simple3DCombModem(ntsc.NtscCombModem(LineConfig((w, h))))
    comp1, out1 = _run_sharded(m, rgb, 1)
    for world in (2, 4, 8):
        comp_w, out_w = _run_sharded(m, rgb, world)
        assert np.array_equal(comp_w, comp1)
        assert np.array_equal(out_w, out1)
    om = oracle.build(oracle.ModemSpec('ntsc_3d', 'NTSC', w, h))
    for f in (0, 74, 75, 299, 300, 599):                       # first, shard boundaries +-1 (8 and 2 shards), last
        assert _lsb(comp1[f], oframe.encode_frame_u8(om, f, rgb[f])) <= 1
        assert _lsb(out1[f], oframe.decode_frame_u8(om, f, comp1[f])) <= 1
    # round trip: smooth content survives NTSC encode -> 3-line comb decode (away from the frame border)
    err = np.abs(out1[:8, 8:-8, 16:-16].astype(np.int32) - rgb[:8, 8:-8, 16:-16].astype(np.int32))
    assert err.mean() < 6.0


@pytest.mark.parametrize('which', ['secam_avg', 'niir_hue'])
def test_1000_frames_batch_invariance(which, cuda_required):
    import torch
    n, h, w = 1000, 576, 720
    lc = LineConfig((w, h))
    if which == 'secam_avg':
        m = comb.ColorAveragingModem(secam.SecamModem(lc))
        spec = oracle.ModemSpec('secam', 'SECAM', w, h, None, True)
    else:
        m = niir.HueCorrectingNiirModem(lc)
        spec = oracle.ModemSpec('niir_hue', 'PAL', w, h)
    rgb = _frames(n, h, w, seed=9)
    x = torch.from_numpy(rgb).cuda()
    comp = m.encode_frames(x, first_frame=0)
    out = m.decode_frames(comp, first_frame=0)
    comp, out = comp.cpu().numpy(), out.cpu().numpy()
    del x
    # host entry points (32-frame chunks over three streams) on a window that does not start at frame 0
    lo, hi = 333, 333 + 100
    assert np.array_equal(m.encode_frames_host(rgb[lo:hi], lo), comp[lo:hi])
    assert np.array_equal(m.decode_frames_host(comp[lo:hi], lo), out[lo:hi])
    om = oracle.build(spec)
    for f in (0, 1, 500, 999):
        assert _lsb(comp[f], oframe.encode_frame_u8(om, f, rgb[f])) <= 1
        assert _lsb(out[f], oframe.decode_frame_u8(om, f, comp[f])) <= 1


HD = [('ntsc_3d', 'NTSC443', 'NTSC_525',
       lambda lc: comb.Simple3DCombModem(ntsc.NtscCombModem(lc, ntsc.NtscVariant.NTSC443))),
      ('pal_d', 'PAL_N', 'GERBER_625', lambda lc: pal.PalDModem(lc, pal.PalVariant.PAL_N)),
      ('secam', 'SECAM_E', 'FRENCH_819', lambda lc: comb.ColorAveragingModem(secam.SecamModem(lc, secam.SecamVariant.SECAM_E)))]


@pytest.mark.parametrize('kind,variant,std,make', HD, ids=[h[0] + '-' + h[1] for h in HD])
def test_1080p_batches(kind, variant, std, make, cuda_required):
    import torch
    n, h, w = 96, 1080, 1920
    m = make(LineConfig((w, h), getattr(LS, std)))
    base = synth_frames_u8(4, h, w, first_frame=0, seed=2)
    rgb = base[np.arange(n) % 4]
    x = torch.from_numpy(rgb).cuda()
    comp = m.encode_frames(x, first_frame=0)
    out = m.decode_frames(comp, first_frame=0)
    # the same frames as a later, shorter batch at their absolute position
    comp_b = m.encode_frames(x[40:56], first_frame=40)
    out_b = m.decode_frames(comp_b, first_frame=40)
    assert torch.equal(comp_b, comp[40:56])
    assert torch.equal(out_b, out[40:56])
    om = oracle.build(oracle.ModemSpec(kind, variant, w, h, std, kind == 'secam'))
    f = 41
    c_ref = oframe.encode_frame_u8(om, f, rgb[f])
    assert _lsb(comp[f].cpu().numpy(), c_ref) <= 1
    assert _lsb(out[f].cpu().numpy(), oframe.decode_frame_u8(om, f, comp[f].cpu().numpy())) <= 1
