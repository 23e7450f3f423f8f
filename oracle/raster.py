"""Raster geometry and subcarrier phase (oracle; test infrastructure only).

Follows reference color_modem/line.py:6-65 and color_modem/utils.py:67-88.
"""
import fractions

import numpy as np

# (frame_rate, total_lines, odd_first, odd_last, even_first, even_last, total_width_factor)
# reference color_modem/line.py:42-46
STANDARDS = {
    'BAIRD_405': (25.0, 405, 16, 203, 218, 405, 1.2),
    'NTSC_525': (30000.0 / 1001.0, 525, 21, 263, 283, 525, 858.0 / 720.0),
    'GERBER_625': (25.0, 625, 336, 623, 23, 310, 1.2),
    'FRENCH_819': (25.0, 819, 39, 407, 448, 816, 1.2),
    'BELGIAN_819': (25.0, 819, 437, 816, 27, 406, 1.2),
}


def active_lines(std):
    _, _, of, ol, ef, el, _ = std
    return (ol - of) + (el - ef) + 2            # line.py:24-26


def detect_standard(height):
    """Smallest standard with >= height active lines (line.py:28-39)."""
    best = None
    for name, std in sorted(STANDARDS.items(), key=lambda kv: active_lines(kv[1]), reverse=True):
        if active_lines(std) < height:
            break
        best = name
    if best is None:
        raise IndexError('No supported line standard supports %d lines' % (height,))
    return best


class Raster(object):
    def __init__(self, width, height, standard=None):
        if standard is None:
            standard = detect_standard(height)
        self.name = standard
        self.width, self.height = int(width), int(height)
        (self.frame_rate, self.total_lines, self.odd_first, _ol, self.even_first, _el,
         self.width_factor) = STANDARDS[standard]
        # line.py:53
        self.fs = self.frame_rate * self.total_lines * width * self.width_factor
        # line.py:55
        self.digital_shift = (active_lines(STANDARDS[standard]) - height) // 2

    def analog_line(self, y):
        """line.py:57-62, vectorised; numpy // and % are floor-based like Python's."""
        adj = np.asarray(y, dtype=np.int64) + self.digital_shift
        return np.where(adj % 2 == 0, self.even_first + adj // 2, self.odd_first + adj // 2)

    def is_alternate(self, frame, y):
        """line.py:64-65"""
        return self.analog_line(y) % 2 == frame % 2

    # ---- constant-frequency carrier, utils.py:67-88 -------------------------------------
    def line_shift(self, fsc):
        return 2.0 * np.pi * ((fsc / (self.frame_rate * self.total_lines)) % 1.0)

    def frame_shift(self, fsc):
        return 2.0 * np.pi * ((fsc / self.frame_rate) % 1.0)

    def frame_cycle(self, fsc):
        return fractions.Fraction(fsc / self.frame_rate).limit_denominator().denominator

    def start_phase(self, fsc, frame, y):
        ref = min(self.odd_first, self.even_first)
        fr = frame % self.frame_cycle(fsc)
        fshift = (fr * self.frame_shift(fsc)) % (2.0 * np.pi)
        lshift = ((self.analog_line(y) - ref) * self.line_shift(fsc)) % (2.0 * np.pi)
        return (fshift + lshift) % (2.0 * np.pi)

    # ---- field helpers (frame driver semantics, image.py:47-55 / 75-83) ------------------
    def rows(self):
        return np.arange(self.height)

    def next_in_field(self):
        """Row fed together with y when a wrapper has a one-line delay: y+2, wrapped down by 2 while >= H."""
        y = self.rows() + 2
        while np.any(y >= self.height):
            y = np.where(y >= self.height, y - 2, y)
        return y

    def is_field_top(self):
        return self.rows() < 2
