"""A/B aid for library variants (tools/variants.sh): device time and output digests of one modem's encode->decode.

    CM_B200_LIB=tools/variants/<name>.so python tools/ab.py [pald|ntsc3d|secam|niir] [frames]

Two builds agree bit for bit when the digests of the composite and of the decoded frames match."""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                                            # noqa: E402
from color_modem_b200.line import LineConfig                            # noqa: E402
from color_modem_b200.color import ntsc, pal, secam, niir               # noqa: E402
from color_modem_b200 import comb                                       # noqa: E402
from color_modem_b200.synth import synth_frames_u8                      # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else 'pald'
F = int(sys.argv[2]) if len(sys.argv) > 2 else 256
lc5, lc6 = LineConfig((720, 480)), LineConfig((720, 576))
m = {'pald': lambda: pal.PalDModem(lc6), 'ntsc3d': lambda: comb.Simple3DCombModem(ntsc.NtscCombModem(lc5)),
     'secam': lambda: comb.ColorAveragingModem(secam.SecamModem(lc6)),
     'niir': lambda: niir.HueCorrectingNiirModem(lc6)}[which]()
rgb = torch.from_numpy(synth_frames_u8(8, m.height, m.width)).repeat(F // 8, 1, 1, 1).contiguous().cuda()
comp = m.encode_frames(rgb)
out = m.decode_frames(comp)
for _ in range(3):
    m.encode_frames(rgb, out=comp)
    m.decode_frames(comp, out=out)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
reps = 10
e[0].record()
for _ in range(reps):
    m.encode_frames(rgb, out=comp)
e[1].record()
for _ in range(reps):
    m.decode_frames(comp, out=out)
e[2].record()
torch.cuda.synchronize()
te, td = e[0].elapsed_time(e[1]) / reps, e[1].elapsed_time(e[2]) / reps
dg = [hashlib.sha256(t.cpu().numpy().tobytes()).hexdigest()[:16] for t in (comp, out)]
print('%s %s x%d: enc %.2f dec %.2f us/frame = %.0f frames/s | composite %s rgb %s' % (
    os.environ.get('CM_B200_LIB', 'default')[-24:], which, F, 1e3 * te / F, 1e3 * td / F, F / (te + td) * 1e3, dg[0], dg[1]))
