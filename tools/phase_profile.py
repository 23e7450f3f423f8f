"""Tuning aid: average cycles per barrier-delimited phase of the PAL-D decode kernel (thread 0 of every CTA)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from color_modem_b200 import _native as N
from color_modem_b200.line import LineConfig
from color_modem_b200.color.pal import PalDModem
from color_modem_b200.synth import synth_frames_u8

F = 64
m = PalDModem(LineConfig((720, 576)))
rgb = torch.from_numpy(synth_frames_u8(16, 576, 720)).repeat(4, 1, 1, 1).cuda()
comp = m.encode_frames(rgb)
out = m.decode_frames(comp)
torch.cuda.synchronize()
cnt = torch.zeros(32, dtype=torch.int64, device='cuda')
N.check(N.load().cm_phase_profile(m._handle(), cnt.data_ptr()))
m.decode_frames(comp, out=out)
torch.cuda.synchronize()
N.check(N.load().cm_phase_profile(m._handle(), None))
c = cnt.cpu().numpy().astype(float)
names = ['load', 'up2', 'tail', 'BP', 'down2 E', 'up2 G', 'tail', 'LP S/D', 'down2 S,D', 'rotate', 'tail', 'pre_lp', 'final']
ncta = F * 2 * 72 + F * 2   # combed groups + top-row CTAs of the band-split launch (not instrumented)
tot = c.sum()
for i, v in enumerate(c):
    if v:
        print('%-10s %8.0f cycles/CTA  %5.1f %%' % (names[i] if i < len(names) else i, v / (F * 2 * 72), 100 * v / tot))
print('total %.0f cycles/CTA' % (tot / (F * 2 * 72)))
