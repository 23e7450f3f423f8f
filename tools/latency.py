"""Single-frame latency (BASELINE configs[0] and configs[1] are single-frame configurations).

    python tools/latency.py [--reps 200] [--json out.json]

For PalDModem 720x576 and NtscModem / Simple3DCombModem(NtscCombModem) 720x480, one frame at a time:
  device   encode + decode of ONE device-resident frame, CUDA events around the pair, median / p90 over `reps` runs
  host     ImageModem.modulate_batch + demodulate_batch of one frame in pinned host memory, wall clock (copies included)
  graph    the device pair replayed from a CUDA graph (one graph launch instead of 5-6 kernel launches)
with the rows a CTA walks through set to 2 (the throughput setting) and to 1 (CM_RPC=1: twice the CTAs, half the rows each).
"""
import argparse
import json
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch                                                            # noqa: E402
from color_modem_b200.line import LineConfig                             # noqa: E402
from color_modem_b200.color import ntsc, pal                             # noqa: E402
from color_modem_b200 import comb                                        # noqa: E402
from color_modem_b200.image import ImageModem                            # noqa: E402
from color_modem_b200.synth import synth_frames_u8                       # noqa: E402


def pct(v, q):
    v = sorted(v)
    return v[min(len(v) - 1, int(q * len(v)))]


def measure(name, make, reps):
    m = make()
    h, w = m.height, m.width
    rgb_h = torch.from_numpy(synth_frames_u8(1, h, w, seed=1)).pin_memory()
    rgb = rgb_h.cuda()
    comp = m.encode_frames(rgb)
    out = m.decode_frames(comp)
    for _ in range(10):
        m.encode_frames(rgb, out=comp)
        m.decode_frames(comp, out=out)
    torch.cuda.synchronize()
    dev = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        m.encode_frames(rgb, out=comp)
        m.decode_frames(comp, out=out)
        b.record()
        b.synchronize()
        dev.append(a.elapsed_time(b) * 1e3)
    # the same pair from a CUDA graph
    graph_us = None
    try:
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            m.encode_frames(rgb, out=comp)
            m.decode_frames(comp, out=out)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            m.encode_frames(rgb, out=comp)
            m.decode_frames(comp, out=out)
        for _ in range(5):
            g.replay()
        torch.cuda.synchronize()
        gr = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            g.replay()
            b.record()
            b.synchronize()
            gr.append(a.elapsed_time(b) * 1e3)
        graph_us = {'median': statistics.median(gr), 'p90': pct(gr, 0.9)}
    except Exception as e:                                               # noqa: BLE001
        graph_us = {'error': str(e)[:200]}
    img = ImageModem(m)
    np_rgb = rgb_h.numpy()
    hc = torch.empty((1, h, m.composite_width), dtype=torch.uint8).pin_memory().numpy()
    ho = torch.empty((1, h, m.output_width, 3), dtype=torch.uint8).pin_memory().numpy()
    for _ in range(5):
        img.modulate_batch(np_rgb, 0, out=hc)
        img.demodulate_batch(hc, 0, out=ho)
    host = []
    for _ in range(reps):
        t0 = time.perf_counter()
        img.modulate_batch(np_rgb, 0, out=hc)
        img.demodulate_batch(hc, 0, out=ho)
        host.append((time.perf_counter() - t0) * 1e6)
    return {'modem': name, 'rows_per_cta': int(os.environ.get('CM_RPC', '2')),
            'device_us': {'median': statistics.median(dev), 'p90': pct(dev, 0.9), 'min': min(dev)},
            'graph_us': graph_us,
            'host_us': {'median': statistics.median(host), 'p90': pct(host, 0.9), 'min': min(host)}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reps', type=int, default=200)
    ap.add_argument('--json', default='')
    a = ap.parse_args()
    lc5, lc6 = LineConfig((720, 480)), LineConfig((720, 576))
    cases = [('PalDModem PAL 720x576 (BASELINE configs[1])', lambda: pal.PalDModem(lc6)),
             ('NtscModem NTSC 720x480 (BASELINE configs[0])', lambda: ntsc.NtscModem(lc5)),
             ('Simple3DCombModem(NtscCombModem) NTSC 720x480', lambda: comb.Simple3DCombModem(ntsc.NtscCombModem(lc5)))]
    rows = []
    for rpc in ('2', '1'):
        os.environ['CM_RPC'] = rpc                    # read once per handle, in cm_create
        for name, make in cases:
            r = measure(name, make, a.reps)
            rows.append(r)
            print(json.dumps(r), flush=True)
    del os.environ['CM_RPC']
    if a.json:
        with open(a.json, 'w') as f:
            json.dump(rows, f, indent=1)


if __name__ == '__main__':
    main()
