"""Batched frame ingest: raw packed video files <-> the batch entry points.

The reference opens and saves one PIL image per frame (cli.py:19,62,64).  Here frame sequences are read as packed raw
video — ``rgb24`` ([N, H, W, 3] bytes, what ``ffmpeg -f rawvideo -pix_fmt rgb24`` writes) and ``gray8`` composite ([N, H, Wc])
— straight into pinned host buffers that the native host entry points copy from: a reader thread fills batch i+1 and a
writer thread drains batch i-1 while batch i is on the GPU (whose own three-stream pipeline overlaps its copies with the
kernels).  No per-frame Python objects, no PIL on this path.
"""
import os
import queue
import threading

import numpy


def _pinned(shape):
    """uint8 array in page-locked memory when torch + CUDA are there (the copies then run at link speed), plain otherwise."""
    try:
        import torch
        if torch.cuda.is_available():
            return torch.empty(shape, dtype=torch.uint8).pin_memory().numpy()
    except Exception:                                                   # noqa: BLE001
        pass
    return numpy.empty(shape, dtype=numpy.uint8)


class RawVideo(object):
    """A file of packed frames of ``frame_shape`` bytes each (``(H, W, 3)`` for rgb24, ``(H, W)`` for gray8)."""

    def __init__(self, path, frame_shape, mode='r'):
        self.path, self.frame_shape = path, tuple(int(v) for v in frame_shape)
        self.frame_bytes = int(numpy.prod(self.frame_shape))
        self._f = open(path, 'rb' if mode == 'r' else ('r+b' if os.path.exists(path) and mode == 'a' else 'wb'))
        self.nframes = os.path.getsize(path) // self.frame_bytes if mode == 'r' else 0

    def read_into(self, first, out):
        """Frames [first, first + len(out)) into ``out`` (a C-contiguous uint8 array); returns the number read."""
        self._f.seek(first * self.frame_bytes)
        got = self._f.readinto(memoryview(out).cast('B'))
        return got // self.frame_bytes

    def write_at(self, first, frames):
        self._f.seek(first * self.frame_bytes)
        self._f.write(memoryview(numpy.ascontiguousarray(frames)).cast('B'))

    def close(self):
        self._f.close()


def frame_batches(first, last, batch):
    return [(b, min(b + batch, last)) for b in range(first, last, batch)]


def run_file(modem, op, src_path, dst_path, dst2_path=None, first_frame=0, frames=None, batch=64, out_offset=None):
    """Stream a raw video file through ``modem``.

    op = 'modulate'   rgb24 -> gray8 composite            (dst)
         'demodulate' gray8 composite -> rgb24            (dst)
         'transcode'  rgb24 -> rgb24 decoded (dst) and, if dst2_path is given, the gray8 composite (dst2)
    Frame i of the file is absolute frame i (carrier phase and line parity depend on it).  ``frames`` = (a, b) restricts
    the run to file frames [a, b); they are written at the same positions of the outputs (``out_offset`` shifts them),
    so several processes — one per GPU, each with its own contiguous range — can fill one output file.
    Returns the number of frames processed."""
    from .image import ImageModem
    img = ImageModem(modem)
    h, w, wc, wo = modem.height, modem.width, modem.composite_width, modem.output_width
    in_shape = (h, wc) if op == 'demodulate' else (h, w, 3)
    out_shape = (h, wc) if op == 'modulate' else (h, wo, 3)
    src = RawVideo(src_path, in_shape)
    a, b = (0, src.nframes) if frames is None else (max(0, frames[0]), min(src.nframes, frames[1]))
    shift = -a if out_offset is None else out_offset - a
    dst = RawVideo(dst_path, out_shape, 'a')
    dst2 = RawVideo(dst2_path, (h, wc), 'a') if (op == 'transcode' and dst2_path) else None
    ring = 3
    bufs_in = [_pinned((batch,) + in_shape) for _ in range(ring)]
    bufs_out = [_pinned((batch,) + out_shape) for _ in range(ring)]
    bufs_mid = [_pinned((batch, h, wc)) for _ in range(ring)] if dst2 else [None] * ring
    free, filled, done = queue.Queue(), queue.Queue(), queue.Queue()
    for i in range(ring):
        free.put(i)
    err = []

    def reader():
        try:
            for lo, hi in frame_batches(a, b, batch):
                i = free.get()
                n = src.read_into(lo, bufs_in[i][:hi - lo])
                filled.put((i, lo, n))
        except Exception as e:                                           # noqa: BLE001
            err.append(e)
        filled.put(None)

    def writer():
        try:
            while True:
                item = done.get()
                if item is None:
                    break
                i, lo, n = item
                dst.write_at(lo + shift, bufs_out[i][:n])
                if dst2:
                    dst2.write_at(lo + shift, bufs_mid[i][:n])
                free.put(i)
        except Exception as e:                                           # noqa: BLE001
            err.append(e)

    tr, tw = threading.Thread(target=reader), threading.Thread(target=writer)
    tr.start()
    tw.start()
    total = 0
    try:
        while True:
            item = filled.get()
            if item is None or err:
                break
            i, lo, n = item
            if n:
                if op == 'modulate':
                    img.modulate_batch(bufs_in[i][:n], first_frame + lo, out=bufs_out[i][:n])
                elif op == 'demodulate':
                    img.demodulate_batch(bufs_in[i][:n], first_frame + lo, out=bufs_out[i][:n])
                else:
                    img.transcode_batch(bufs_in[i][:n], first_frame + lo, out=bufs_out[i][:n],
                                        comp_out=bufs_mid[i][:n] if dst2 else None, want_composite=bool(dst2))
            total += n
            done.put((i, lo, n))
    finally:
        done.put(None)
        tr.join()
        tw.join()
        src.close()
        dst.close()
        if dst2:
            dst2.close()
    if err:
        raise err[0]
    return total
