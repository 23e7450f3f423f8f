"""Raster geometry: drop-in for the reference's ``color_modem.line`` (line.py:6-65).

Same names, fields and behaviour: ``LineStandard`` (a namedtuple with the five presets as class attributes and
``detect``), ``LineConfig(size, line_standard=None)`` with ``fs``, ``line_standard``, ``analog_line`` and
``is_alternate_line``.  Pure host-side integer/float bookkeeping — the kernels receive these numbers in cm_desc.
"""
import collections

_Fields = collections.namedtuple(
    'LineStandard',
    'frame_rate total_lines ' + ' '.join('%s_field_%s_active_line' % (f, e) for f in ('odd', 'even') for e in ('first', 'last'))
    + ' total_width_factor')


class LineStandard(_Fields):
    __slots__ = ()

    def __new__(cls, *args, **kwargs):
        self = super(LineStandard, cls).__new__(cls, *args, **kwargs)
        odd_span = self.odd_field_last_active_line - self.odd_field_first_active_line
        even_span = self.even_field_last_active_line - self.even_field_first_active_line
        if odd_span < 0 or even_span < 0 or odd_span != even_span or self.active_lines > self.total_lines:
            raise AssertionError('inconsistent line standard')
        return self

    @property
    def active_lines(self):
        return (self.odd_field_last_active_line - self.odd_field_first_active_line) + \
               (self.even_field_last_active_line - self.even_field_first_active_line) + 2

    @classmethod
    def presets(cls):
        return [v for v in vars(cls).values() if isinstance(v, cls)]

    @classmethod
    def detect(cls, active_lines):
        """Smallest preset that has at least ``active_lines`` visible lines; IndexError if none (line.py:28-39)."""
        fitting = [s for s in cls.presets() if s.active_lines >= active_lines]
        if not fitting:
            raise IndexError('No supported line standard supports %d lines' % (active_lines,))
        return min(fitting, key=lambda s: s.active_lines)


# name: frames/s, lines per frame, (first, last) active line of the odd field, of the even field, line length / active length
for _name, (_rate, _lines, _odd, _even, _wf) in {
        'BAIRD_405': (25.0, 405, (16, 203), (218, 405), 1.2),
        'NTSC_525': (30000.0 / 1001.0, 525, (21, 263), (283, 525), 858.0 / 720.0),
        'GERBER_625': (25.0, 625, (336, 623), (23, 310), 1.2),
        'FRENCH_819': (25.0, 819, (39, 407), (448, 816), 1.2),
        'BELGIAN_819': (25.0, 819, (437, 816), (27, 406), 1.2)}.items():
    setattr(LineStandard, _name, LineStandard(_rate, _lines, _odd[0], _odd[1], _even[0], _even[1], _wf))


class LineConfig(object):
    def __init__(self, size, line_standard=None):
        if line_standard is None:
            line_standard = LineStandard.detect(size[1])
        self.size = (int(size[0]), int(size[1]))
        self.line_standard = line_standard
        self.fs = line_standard.frame_rate * line_standard.total_lines * size[0] * line_standard.total_width_factor
        self._line_shift = (line_standard.active_lines - size[1]) // 2

    def analog_line(self, digital_line):
        shifted = digital_line + self._line_shift
        first = (self.line_standard.even_field_first_active_line if shifted % 2 == 0
                 else self.line_standard.odd_field_first_active_line)
        return first + shifted // 2

    def is_alternate_line(self, frame, line):
        return self.analog_line(line) % 2 == frame % 2
