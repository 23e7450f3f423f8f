"""Tuning aid: per-kernel device time (CUDA events inside the library) of one modem's encode->decode.

    python tools/kt.py [pald|ntsc3d|ntsc|pals|pal3d|secam|niir|proto|mac|mac7|pald1080|ntsc3d1080|secam1080|proto1080|mac1080] [frames]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                                            # noqa: E402
from color_modem_b200 import _native as N                                # noqa: E402
from color_modem_b200.line import LineConfig, LineStandard as LS        # noqa: E402
from color_modem_b200.color import ntsc, pal, secam, niir, protosecam, mac   # noqa: E402
from color_modem_b200 import comb                                        # noqa: E402
from color_modem_b200.synth import synth_frames_u8                       # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else 'pald'
F = int(sys.argv[2]) if len(sys.argv) > 2 else 256
lc5, lc6 = LineConfig((720, 480)), LineConfig((720, 576))
make = {'pald': lambda: pal.PalDModem(lc6), 'pals': lambda: pal.PalSModem(lc6), 'pal3d': lambda: pal.Pal3DModem(lc6),
        'ntsc': lambda: ntsc.NtscModem(lc5), 'ntsc2d': lambda: ntsc.NtscCombModem(lc5),
        'ntsc3d': lambda: comb.Simple3DCombModem(ntsc.NtscCombModem(lc5)),
        'secam': lambda: comb.ColorAveragingModem(secam.SecamModem(lc6)), 'niir': lambda: niir.HueCorrectingNiirModem(lc6),
        'proto': lambda: comb.ColorAveragingModem(protosecam.ProtoSecamModem(LineConfig((720, 576), LS.FRENCH_819))),
        'mac': lambda: comb.ColorAveragingModem(mac.MacModem(lc6)),
        'mac7': lambda: mac.MacModem(lc6, mac.MacVariant.D2MAC_7MHZ),
        # 1920x1080 (BASELINE configs[4]; explicit line standards)
        'pald1080': lambda: pal.PalDModem(LineConfig((1920, 1080), LS.GERBER_625), pal.PalVariant.PAL_N),
        'ntsc3d1080': lambda: comb.Simple3DCombModem(ntsc.NtscCombModem(LineConfig((1920, 1080), LS.NTSC_525), ntsc.NtscVariant.NTSC443)),
        'secam1080': lambda: comb.ColorAveragingModem(secam.SecamModem(LineConfig((1920, 1080), LS.GERBER_625), secam.SecamVariant.SECAM_III)),
        'proto1080': lambda: comb.ColorAveragingModem(protosecam.ProtoSecamModem(LineConfig((1920, 1080), LS.FRENCH_819))),
        'mac1080': lambda: mac.MacModem(LineConfig((1920, 1080), LS.GERBER_625), mac.MacVariant.D2MAC_7MHZ)}[which]
m = make()
h, w = m.height, m.width
rgb = torch.from_numpy(synth_frames_u8(8, h, w)).repeat(F // 8, 1, 1, 1).contiguous().cuda()
comp = m.encode_frames(rgb)
out = m.decode_frames(comp)
for _ in range(3):
    m.encode_frames(rgb, out=comp)
    m.decode_frames(comp, out=out)
torch.cuda.synchronize()
impl = getattr(m, '_impl', m)
impl.timing(True)
reps = 5
e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
e[0].record()
for _ in range(reps):
    m.encode_frames(rgb, out=comp)
    m.decode_frames(comp, out=out)
e[1].record()
torch.cuda.synchronize()
tot = e[0].elapsed_time(e[1]) / reps
names = {N.K_ENCODE: 'encode', N.K_BANDSPLIT: 'bandsplit', N.K_PALD: 'pald/rows', N.K_COMB: 'comb', N.K_DECODE_OTHER: 'other/pair'}
parts = []
for kid, name in names.items():
    ms, n = impl.timing_read(kid)
    if n:
        parts.append('%s %.2f us/frame (%d launches/step)' % (name, 1e3 * ms / reps / F, n // reps))
print('%s %s: total %.2f us/frame = %.0f frames/s | %s' % (os.environ.get('CM_B200_LIB', 'default')[-24:], which, 1e3 * tot / F, F / tot * 1e3, ' | '.join(parts)))
