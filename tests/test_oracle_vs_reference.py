"""Builder-container only: re-check the oracle against the live reference at full frame size.
Skipped wherever /root/reference is absent (e.g. the GPU box)."""
import warnings

import numpy as np
import pytest

import oracle
from oracle import frame as oframe
from cases import Case
from refload import reference_available, load_reference
from color_modem_b200.synth import synth_frames_u8

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not reference_available(), reason='/root/reference not present')]

FULL = [
    Case('ntsc', 'NTSC', 720, 480, None, False, 0, 0, 'smooth'),          # BASELINE config 1
    Case('pal_d', 'PAL', 720, 576, None, False, 0, 0, 'smooth'),          # BASELINE config 2
    Case('ntsc_3d', 'NTSC', 720, 480, None, False, 599, 3, 'smooth'),     # BASELINE config 3 (last frame)
    Case('secam', 'SECAM', 720, 576, None, True, 7, 4, 'noise'),          # BASELINE config 4a
    Case('niir_hue', 'PAL', 720, 576, None, False, 999, 5, 'smooth'),     # BASELINE config 4b
    Case('mac', 'D2MAC_7MHZ', 1920, 1080, 'GERBER_625', False, 2, 6, 'smooth'),   # BASELINE config 5 sample
]


@pytest.mark.parametrize('c', FULL, ids=lambda c: '%s-%s-%dx%d' % (c.kind, c.variant, c.width, c.height))
def test_full_frame_u8_identical(c):
    warnings.filterwarnings('ignore')
    from PIL import Image
    ref = load_reference()
    rgb = synth_frames_u8(1, c.height, c.width, first_frame=c.frame, seed=c.seed, kind=c.content)[0]
    driver = ref.image.ImageModem(ref.make_modem(c))
    comp_img = driver.modulate(Image.fromarray(rgb, 'RGB'), c.frame)
    comp_ref = np.asarray(comp_img)
    out_ref = np.asarray(driver.demodulate(comp_img, c.frame))
    modem = oracle.build(oracle.ModemSpec(c.kind, c.variant, c.width, c.height, c.standard, c.chroma_avg, c.notch, c.opt))
    assert np.array_equal(oframe.encode_frame_u8(modem, c.frame, rgb), comp_ref)
    assert np.array_equal(oframe.decode_frame_u8(modem, c.frame, comp_ref), out_ref)
