"""Command line and raw-video ingest.  CPU: argument handling, the composition registry, raw file
I/O.  GPU: files streamed through the modem equal the batch API on the same frames, whole or cut into frame ranges."""
import os

import numpy as np
import pytest

from color_modem_b200 import ingest
from color_modem_b200.__main__ import build_modem, main, parse, registry
from color_modem_b200.synth import synth_frames_u8


def test_parse_and_registry():
    a = parse(['transcode', '--modem', 'pal-d', '--size', '720x576', '--frames', '2:10', '--devices', '0,3', 'in.rgb',
               'out.rgb', 'comp.gray'])
    assert (a.op, a.size, a.frames, a.devices, a.raw, a.output2) == ('transcode', (720, 576), (2, 10), [0, 3], True, 'comp.gray')
    assert parse(['modulate', '--modem', 'ntsc', 'a.png', 'b.png']).raw is False
    for name in registry():
        m = build_modem(name, (720, 576), None, 'FRENCH_819' if 'proto' in name else None)
        assert m.width == 720 and m.height == 576
    with pytest.raises(SystemExit):
        build_modem('nope', (720, 576))
    with pytest.raises(NotImplementedError):
        build_modem('ntsc', (854, 480), None, 'NTSC_525')          # widths must be multiples of 4


def test_raw_video_io(tmp_path):
    frames = synth_frames_u8(5, 8, 16)
    p = str(tmp_path / 'v.rgb')
    frames.tofile(p)
    v = ingest.RawVideo(p, (8, 16, 3))
    assert v.nframes == 5
    buf = np.empty((3, 8, 16, 3), np.uint8)
    assert v.read_into(2, buf) == 3 and np.array_equal(buf, frames[2:])
    assert v.read_into(4, buf) == 1
    v.close()
    w = ingest.RawVideo(str(tmp_path / 'o.rgb'), (8, 16, 3), 'a')
    w.write_at(1, frames[:2])
    w.close()
    back = np.fromfile(str(tmp_path / 'o.rgb'), np.uint8).reshape(3, 8, 16, 3)
    assert np.array_equal(back[1:], frames[:2])
    assert ingest.frame_batches(3, 10, 4) == [(3, 7), (7, 10)]


@pytest.mark.gpu
def test_files_through_the_modem_equal_the_batch_api(tmp_path, cuda_required):
    w, h, n = 720, 32, 7
    rgb = synth_frames_u8(n, h, w, seed=5)
    src = str(tmp_path / 'in.rgb')
    rgb.tofile(src)
    m = build_modem('pal-d', (w, h), None, 'GERBER_625')
    comp_ref = m.encode_frames_host(rgb, 0)
    out_ref = m.decode_frames_host(comp_ref, 0)
    out, comp = str(tmp_path / 'out.rgb'), str(tmp_path / 'comp.gray')
    assert main(['transcode', '--modem', 'pal-d', '--standard', 'GERBER_625', '--size', '%dx%d' % (w, h), '--batch', '3',
                 src, out, comp]) == 0
    assert np.array_equal(np.fromfile(comp, np.uint8).reshape(n, h, w), comp_ref)
    assert np.array_equal(np.fromfile(out, np.uint8).reshape(n, h, w, 3), out_ref)
    # the same in two frame ranges written into one file (what --devices does, one process per range)
    out2 = str(tmp_path / 'out2.rgb')
    open(out2, 'wb').close()
    for rng in ((0, 3), (3, n)):
        ingest.run_file(m, 'demodulate', comp, out2, frames=rng, batch=2, out_offset=rng[0])
    assert np.array_equal(np.fromfile(out2, np.uint8).reshape(n, h, w, 3), out_ref)
    # --frames: absolute frame numbers are kept, the output starts at its first frame
    part = str(tmp_path / 'part.gray')
    assert main(['modulate', '--modem', 'pal-d', '--standard', 'GERBER_625', '--size', '%dx%d' % (w, h), '--frames', '2:5',
                 src, part]) == 0
    assert np.array_equal(np.fromfile(part, np.uint8).reshape(3, h, w), comp_ref[2:5])


@pytest.mark.gpu
def test_transcode_batch_equals_the_two_calls(cuda_required):
    from color_modem_b200.image import ImageModem
    w, h, n = 720, 32, 5
    rgb = synth_frames_u8(n, h, w, seed=6)
    for name, std in (('pal-d', 'GERBER_625'), ('secam-avg', 'GERBER_625'), ('ntsc-3d', 'NTSC_525')):
        im = ImageModem(build_modem(name, (w, h), None, std))
        comp = im.modulate_batch(rgb, 3)
        out = im.demodulate_batch(comp, 3)
        c2, o2 = im.transcode_batch(rgb, 3)
        assert np.array_equal(c2, comp) and np.array_equal(o2, out)
        c3, o3 = im.transcode_batch(rgb, 3, want_composite=False)
        assert c3 is None and np.array_equal(o3, out)


@pytest.mark.gpu
def test_calls_can_be_captured_in_a_cuda_graph(cuda_required):
    """A single frame is launch-bound (tools/latency.py): the device-resident calls are stream-ordered and allocate
    nothing after their first use, so a user can replay them from a CUDA graph."""
    import torch
    w, h = 720, 64
    m = build_modem('pal-d', (w, h), None, 'GERBER_625')
    rgb = torch.from_numpy(synth_frames_u8(1, h, w, seed=7)).cuda()
    comp = m.encode_frames(rgb)
    out = m.decode_frames(comp)
    ref = out.clone()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        m.encode_frames(rgb, out=comp)
        m.decode_frames(comp, out=out)
    s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        m.encode_frames(rgb, out=comp)
        m.decode_frames(comp, out=out)
    out.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
