"""Line lengths between the two rasters of BASELINE (720 and 1920 samples): the row kernels pick their team geometry by
line length (csrc/cm_api.cu: plan_*_kernel), and 1024 / 1440 samples land on the geometries that neither 720 nor 1920
exercises (4 warps with the short chunks for the QAM family, the 4- and 8-warp geometries at a shorter line for the others);
MAC runs at 1008 samples (ratios 5/7 and 5/14: up > 4, the general resampler instead of the aligned polyphase form) and at 1200
(3/5 and 3/10: polyphase with a window alignment that changes from output group to output group).  u8 frames within +-1 LSB of the float64 oracle, like every golden case."""
import numpy as np
import pytest

import oracle
from oracle import frame as oframe
from test_gpu_edges import MAKERS, _lsb
from color_modem_b200.line import LineConfig, LineStandard as LS
from color_modem_b200.synth import synth_frames_u8

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('w', [1024, 1440])
@pytest.mark.parametrize('kind', sorted(MAKERS))
def test_intermediate_line_lengths(kind, w, cuda_required):
    import torch
    variant, std, make = MAKERS[kind]
    if kind == 'mac':
        w = {1024: 1008, 1440: 1200}[w]
    h, first, n = 12, 3, 2
    try:
        om = oracle.build(oracle.ModemSpec(kind, variant, w, h, std))
    except Exception as e:                                   # noqa: BLE001  (a preset whose design fails at this rate in scipy)
        pytest.skip('reference design does not exist at this sampling rate: %s' % e)
    m = make(LineConfig((w, h), getattr(LS, std)))
    rgb = synth_frames_u8(n, h, w, first_frame=first, seed=17)
    comp = m.encode_frames(torch.from_numpy(rgb).cuda(), first_frame=first)
    out = m.decode_frames(comp, first_frame=first)
    comp, out = comp.cpu().numpy(), out.cpu().numpy()
    for i in range(n):
        assert _lsb(comp[i], oframe.encode_frame_u8(om, first + i, rgb[i])) <= 1, (kind, w, i)
        assert _lsb(out[i], oframe.decode_frame_u8(om, first + i, comp[i])) <= 1, (kind, w, i)
