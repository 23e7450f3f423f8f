#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) in this container.

    python tests/golden/make_golden.py [--missing]        (--missing: only the cases without a fixture yet)

For every case in tests/cases.py this imports kFYatek/color_modem read-only, installs the scipy
``iirdesign`` validation shim of SURVEY.md §8c (needed by NTSC / PAL-M / PAL-N presets on scipy >= 1.6,
bit-identical to stock scipy on valid input), builds the reference modem composition and calls the
reference's own frame driver ``ImageModem.modulate`` / ``.demodulate`` (image.py:27,58).  Stored per case:

    comp_u8   [H, Wc]     composite frame as produced by ImageModem.modulate
    rgb_u8    [H, Wo, 3]  frame decoded by ImageModem.demodulate from comp_u8
    comp_f64  [R, Wc]     float64 level-mapped composite (argument of image._as_bytes) for rows FLOAT_ROWS
    rgb_f64   [R, Wo, 3]  float64 decoded RGB (arguments of image._as_bytes) for rows FLOAT_ROWS

The inputs are regenerated from color_modem_b200.synth (integer-only, platform independent).
/root/reference does not exist on the GPU box; only the committed .npz files travel.
"""
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.dont_write_bytecode = True
os.environ['PYTHONDONTWRITEBYTECODE'] = '1'

import numpy as np                                             # noqa: E402

from refload import load_reference                             # noqa: E402
from cases import GOLDEN_CASES, FLOAT_ROWS, case_id            # noqa: E402
from color_modem_b200.synth import synth_frames_u8             # noqa: E402


def main():
    warnings.filterwarnings('ignore')
    ref = load_reference()
    from PIL import Image
    only_missing = '--missing' in sys.argv
    for c in GOLDEN_CASES:
        if only_missing and os.path.exists(os.path.join(HERE, case_id(c) + '.npz')):
            continue
        rgb = synth_frames_u8(1, c.height, c.width, first_frame=c.frame, seed=c.seed, kind=c.content)[0]
        modem = ref.make_modem(c)
        driver = ref.image.ImageModem(modem)
        with ref.capture_floats() as cap:
            comp_img = driver.modulate(Image.fromarray(rgb, 'RGB'), c.frame)
        comp = np.asarray(comp_img)
        comp_f = ref.rows_in_raster_order(cap, c.height, 1)
        with ref.capture_floats() as cap:
            out_img = driver.demodulate(comp_img, c.frame)
        out = np.asarray(out_img)
        rgb_f = ref.rows_in_raster_order(cap, c.height, 3)
        rows = [r for r in FLOAT_ROWS if r < c.height]
        path = os.path.join(HERE, case_id(c) + '.npz')
        np.savez_compressed(path, comp_u8=comp, rgb_u8=out,
                            comp_f64=comp_f[rows, 0], rgb_f64=np.moveaxis(rgb_f[rows], 1, -1))
        print('%-60s comp %s rgb %s  %d KB' % (case_id(c), comp.shape, out.shape, os.path.getsize(path) // 1024))


if __name__ == '__main__':
    main()
