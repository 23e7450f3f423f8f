"""Line-memory wrappers — drop-in for ``color_modem.comb`` (comb.py:71-167).

The reference composes decoders at run time out of Python objects that call each other line by line.  On the
GPU every supported composition is one fused kernel family, so the wrappers here *select* the composition:
``Simple3DCombModem(NtscCombModem(lc))`` becomes the 3-line NTSC comb kernel, ``ColorAveragingModem(backend)``
turns on the look-ahead chroma averaging of the backend's encoder.  Compositions that the reference would accept
but that are not built raise NotImplementedError (there is no CPU fallback).
"""
import copy

import numpy

from . import _native as N


def avg(val1, val2):
    """comb.py:9-10"""
    return 0.5 * (val1 + val2)


def minavg(val1, val2):
    """comb.py:13-15: the smaller magnitude where the two estimates agree in sign, zero where they do not"""
    sign = (1.0 - numpy.signbit(val1)) - numpy.signbit(val2)
    return sign * numpy.minimum(numpy.abs(val1), numpy.abs(val2))


def _avg_mode(fn):
    """Which of the reference's two combiners `fn` is: False = mean (None or comb.avg), True = comb.minavg."""
    if fn is None or fn is avg or getattr(fn, '__name__', '') == 'avg':
        return False
    if fn is minavg or getattr(fn, '__name__', '') == 'minavg':
        return True
    raise NotImplementedError('avg= accepts comb.avg or comb.minavg; arbitrary Python callables cannot run in the kernels')


def _clone(backend, **changes):
    m = copy.copy(backend)
    m._handles = {}
    m._enc_mem = m._dec_mem = None
    for k, v in changes.items():
        setattr(m, k, v)
    return m


class _Wrapper(object):
    """Delegates the protocol to a re-configured clone of the backend."""

    def __init__(self, backend, impl):
        self.backend = backend
        self._impl = impl

    config = property(lambda self: self.backend.config)
    line_config = property(lambda self: self.backend.line_config)
    modulation_delay = property(lambda self: self._impl.modulation_delay)
    demodulation_delay = property(lambda self: self._impl.demodulation_delay)

    def __getattr__(self, name):          # encode_frames, decode_frames, modulate, demodulate, describe, ...
        return getattr(self._impl, name)


class Simple3DCombModem(_Wrapper):
    """comb.py:125-127 over comb.py:71-122 (delay=True)."""

    def __init__(self, backend, notch=0.0, avg=None):
        from .color.ntsc import NtscCombModem
        if type(backend) is not NtscCombModem:
            raise NotImplementedError('Simple3DCombModem is built for NtscCombModem backends only')
        # the wrapper's own notch (comb.py:108-109); the backend's is never applied (it is called with strip_chroma=False)
        impl = _clone(backend, kind=N.KIND_NTSC_3D, decoder_rows=3, _notch_q=float(notch), _minavg=_avg_mode(avg),
                      demodulation_delay=getattr(backend, 'demodulation_delay', 0) + 1)
        super(Simple3DCombModem, self).__init__(backend, impl)


class SimpleCombModem(_Wrapper):
    def __init__(self, backend, notch=0.0, avg=None, delay=False):
        raise NotImplementedError('SimpleCombModem (delay=False) is not built; use Simple3DCombModem(NtscCombModem)')


class ColorAveragingModem(_Wrapper):
    """Encoder-side averaging of each line's chroma with the next line of the field (comb.py:130-167)."""

    def __init__(self, backend):
        if getattr(backend, 'encoder_lookahead', False):
            raise NotImplementedError('ColorAveragingModem over a look-ahead encoder is not a reference composition')
        base_flags = backend._flags

        def flags():
            return base_flags() | N.FLAG_CHROMA_AVG

        impl = _clone(backend, _flags=flags, encoder_lookahead=True,
                      modulation_delay=getattr(backend, 'modulation_delay', 0) + 1)
        super(ColorAveragingModem, self).__init__(backend, impl)
