"""Throughput sweep over modem compositions / presets (BASELINE.json configs 3-5) on one GPU.

    python tools/sweep.py [--frames N] [--size 720x576|1920x1080|both] [--json out.json]

Device-resident encode->decode frames/s (CUDA events, 3 warm-up + 5 timed passes) for every composition at 720-wide
and for the 20 preset/standard pairs of config 5 at 1920x1080 (explicit line standards).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch                                                            # noqa: E402
from color_modem_b200.line import LineConfig, LineStandard as LS        # noqa: E402
from color_modem_b200.color import ntsc, pal, secam, niir, protosecam, mac   # noqa: E402
from color_modem_b200 import comb                                        # noqa: E402
from color_modem_b200.synth import synth_frames_u8                       # noqa: E402

NV, PV, SV = ntsc.NtscVariant, pal.PalVariant, secam.SecamVariant


def sd_cases():
    lc5, lc6 = LineConfig((720, 480)), LineConfig((720, 576))
    lc8 = LineConfig((720, 576), LS.FRENCH_819)
    return [
        ('NtscModem NTSC 720x480', lambda: ntsc.NtscModem(lc5)),
        ('NtscCombModem NTSC 720x480', lambda: ntsc.NtscCombModem(lc5)),
        ('Simple3DComb(NtscComb) NTSC 720x480', lambda: comb.Simple3DCombModem(ntsc.NtscCombModem(lc5))),
        ('PalSModem PAL 720x576', lambda: pal.PalSModem(lc6)),
        ('PalDModem PAL 720x576', lambda: pal.PalDModem(lc6)),
        ('Pal3DModem PAL 720x576', lambda: pal.Pal3DModem(lc6)),
        ('ColorAveraging(SecamModem) SECAM 720x576', lambda: comb.ColorAveragingModem(secam.SecamModem(lc6))),
        ('SecamModem SECAM 720x576', lambda: secam.SecamModem(lc6)),
        ('HueCorrectingNiirModem 720x576', lambda: niir.HueCorrectingNiirModem(lc6)),
        ('NiirModem 720x576', lambda: niir.NiirModem(lc6)),
        ('ColorAveraging(ProtoSecam) 720x576 FRENCH_819', lambda: comb.ColorAveragingModem(protosecam.ProtoSecamModem(lc8))),
        ('ColorAveraging(MacModem 12 MHz) 720x576', lambda: comb.ColorAveragingModem(mac.MacModem(lc6))),
        ('MacModem 7 MHz 720x576', lambda: mac.MacModem(lc6, mac.MacVariant.D2MAC_7MHZ)),
    ]


def hd_cases():
    def lc(std):
        return LineConfig((1920, 1080), std)
    out = []
    for name, std in (('NTSC_A', LS.BAIRD_405), ('NTSC443', LS.NTSC_525), ('NTSC361', LS.NTSC_525),
                      ('NTSC_I', LS.GERBER_625), ('NTSC_N', LS.GERBER_625)):
        out.append(('Simple3DComb(NtscComb) %s 1920x1080' % name,
                    lambda n=name, s=std: comb.Simple3DCombModem(ntsc.NtscCombModem(lc(s), getattr(NV, n)))))
    for name, std in (('PAL_M', LS.NTSC_525), ('PAL_N', LS.GERBER_625)):
        out.append(('PalDModem %s 1920x1080' % name, lambda n=name, s=std: pal.PalDModem(lc(s), getattr(PV, n))))
    for name, std in (('SECAM_A', LS.BAIRD_405), ('SECAM_M', LS.NTSC_525), ('SECAM_I', LS.GERBER_625),
                      ('SECAM_II', LS.GERBER_625), ('SECAM_III', LS.GERBER_625), ('SECAM_N', LS.GERBER_625),
                      ('SECAM_E', LS.FRENCH_819)):
        out.append(('ColorAveraging(SecamModem) %s 1920x1080' % name,
                    lambda n=name, s=std: comb.ColorAveragingModem(secam.SecamModem(lc(s), getattr(SV, n)))))
    out.append(('ColorAveraging(ProtoSecam) SECAM_1957 1920x1080',
                lambda: comb.ColorAveragingModem(protosecam.ProtoSecamModem(lc(LS.FRENCH_819)))))
    out.append(('MacModem D2MAC_7MHZ 1920x1080', lambda: mac.MacModem(lc(LS.GERBER_625), mac.MacVariant.D2MAC_7MHZ)))
    return out


def measure(name, make, frames):
    m = make()
    h, w = m.height, m.width
    base = synth_frames_u8(min(frames, 4), h, w, seed=1)
    rgb = torch.from_numpy(base).repeat(-(-frames // base.shape[0]), 1, 1, 1)[:frames].contiguous().cuda()
    comp = m.encode_frames(rgb)
    out = m.decode_frames(comp)
    for _ in range(2):
        m.encode_frames(rgb, out=comp)
        m.decode_frames(comp, out=out)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    reps = 5
    e[0].record()
    for _ in range(reps):
        m.encode_frames(rgb, out=comp)
    e[1].record()
    for _ in range(reps):
        m.decode_frames(comp, out=out)
    e[2].record()
    torch.cuda.synchronize()
    t_enc, t_dec = e[0].elapsed_time(e[1]) / reps, e[1].elapsed_time(e[2]) / reps
    alg = 3 * w * h + 2 * m.composite_width * h + 3 * m.output_width * h
    fps = frames / ((t_enc + t_dec) * 1e-3)
    return {'modem': name, 'frames': frames, 'encode_us_per_frame': 1e3 * t_enc / frames,
            'decode_us_per_frame': 1e3 * t_dec / frames, 'frames_per_s': fps,
            'algorithmic_bytes_per_frame': alg, 'GBps': fps * alg / 1e9}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=64)
    ap.add_argument('--size', default='both')
    ap.add_argument('--json', default='')
    a = ap.parse_args()
    cases = []
    if a.size in ('720x576', 'both'):
        cases += [(n, f, a.frames) for n, f in sd_cases()]
    if a.size in ('1920x1080', 'both'):
        cases += [(n, f, max(8, a.frames // 4)) for n, f in hd_cases()]
    rows = []
    for name, make, frames in cases:
        try:
            r = measure(name, make, frames)
            print('%-52s enc %7.2f us  dec %8.2f us  %9.0f frames/s  %7.1f GB/s' % (
                name, r['encode_us_per_frame'], r['decode_us_per_frame'], r['frames_per_s'], r['GBps']), flush=True)
        except Exception as ex:                                          # noqa: BLE001
            r = {'modem': name, 'error': str(ex)}
            print('%-52s ERROR %s' % (name, ex), flush=True)
        rows.append(r)
    if a.json:
        with open(a.json, 'w') as f:
            json.dump(rows, f, indent=1)


if __name__ == '__main__':
    main()
