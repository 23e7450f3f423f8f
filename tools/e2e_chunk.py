"""Tuning aid: end-to-end frames/s of the host entry points for different CM_HOST_CHUNK values (frames per staged chunk)."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == 'child':
    import time
    import torch
    from concurrent.futures import ThreadPoolExecutor
    from color_modem_b200.line import LineConfig
    from color_modem_b200.color.pal import PalDModem
    from color_modem_b200.image import ImageModem
    from color_modem_b200.synth import synth_frames_u8
    F, H, W = 256, 576, 720
    rgb = torch.from_numpy(synth_frames_u8(16, H, W)).repeat(16, 1, 1, 1).contiguous().pin_memory().numpy()
    comp = [torch.empty((F, H, W), dtype=torch.uint8).pin_memory().numpy() for _ in range(2)]
    out = torch.empty((F, H, W, 3), dtype=torch.uint8).pin_memory().numpy()
    a, b = ImageModem(PalDModem(LineConfig((W, H)))), ImageModem(PalDModem(LineConfig((W, H))))
    a.modulate_batch(rgb, 0, out=comp[0]); b.demodulate_batch(comp[0], 0, out=out)
    t0 = time.perf_counter()
    for _ in range(4):
        a.modulate_batch(rgb, 0, out=comp[0]); a.demodulate_batch(comp[0], 0, out=out)
    seq = 4 * F / (time.perf_counter() - t0)
    pool = ThreadPoolExecutor(2)
    t0 = time.perf_counter()
    for i in range(6):
        x = pool.submit(a.modulate_batch, rgb, 0, comp[(i + 1) & 1]); y = pool.submit(b.demodulate_batch, comp[i & 1], 0, out)
        x.result(); y.result()
    pip = 6 * F / (time.perf_counter() - t0)
    print('CM_HOST_CHUNK=%s sequential %.0f pipelined %.0f frames/s' % (os.environ.get('CM_HOST_CHUNK', 'default'), seq, pip))
else:
    for c in ('4', '8', '16', '32', '64'):
        subprocess.run([sys.executable, __file__, 'child'], env=dict(os.environ, CM_HOST_CHUNK=c))
