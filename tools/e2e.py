"""End-to-end rate of ImageModem.transcode_batch on pinned host buffers (PAL-D 720x576), for A/B runs of the host pipeline:

    CM_HOST_CHUNK=32 python tools/e2e.py [frames per call] [seconds]

prints frames/s with and without the composite copied out, and the two-call (modulate_batch, demodulate_batch) rate."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch                                                            # noqa: E402
from color_modem_b200.line import LineConfig                             # noqa: E402
from color_modem_b200.color.pal import PalDModem                         # noqa: E402
from color_modem_b200.image import ImageModem                            # noqa: E402
from color_modem_b200.synth import synth_frames_u8                       # noqa: E402


def main():
    F = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    secs = float(sys.argv[2]) if len(sys.argv) > 2 else 1.5
    h, w = 576, 720
    img = ImageModem(PalDModem(LineConfig((w, h))))
    base = synth_frames_u8(16, h, w, seed=0)
    rgb = torch.from_numpy(base).repeat(-(-F // 16), 1, 1, 1)[:F].contiguous().pin_memory().numpy()
    comp = torch.empty((F, h, w), dtype=torch.uint8).pin_memory().numpy()
    out = torch.empty((F, h, w, 3), dtype=torch.uint8).pin_memory().numpy()

    def rate(fn):
        fn()
        fn()
        t0 = time.perf_counter()
        fn()
        n = max(3, int(secs / max(time.perf_counter() - t0, 1e-4)))
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        return F * n / (time.perf_counter() - t0)

    r_rgb = rate(lambda: img.transcode_batch(rgb, 0, out=out, want_composite=False))
    r_all = rate(lambda: img.transcode_batch(rgb, 0, out=out, comp_out=comp))

    def two():
        img.modulate_batch(rgb, 0, out=comp)
        img.demodulate_batch(comp, 0, out=out)
    r_two = rate(two)
    print('host_chunk=%s frames/call=%d: transcode rgb-only %.0f f/s | with composite out %.0f | two calls (one thread) %.0f'
          % (os.environ.get('CM_HOST_CHUNK', 'default'), F, r_rgb, r_all, r_two))


if __name__ == '__main__':
    main()
