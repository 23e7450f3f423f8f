"""Frame-level u8 driver of the oracle (test infrastructure only).

Follows reference color_modem/image.py:7-84: u8/255 -> modem -> 0.6v+0.2 -> clip -> rint(255v) -> u8 on the
encode side, (5v-1)/3 on the decode side.  The field-ordered loop, delay priming and bottom-row wrap of
image.py:47-55 / 75-83 are already folded into the stateless row functions of oracle.modems.
"""
import numpy as np


def to_u8(x):
    """image.py:7-8 (numpy.rint is round-half-to-even)."""
    return np.uint8(np.rint(255.0 * np.maximum(np.minimum(x, 1.0), 0.0)))


def composite_level(v):
    return 0.6 * v + 0.2                      # image.py:16-21


def composite_unlevel(v):
    return (5.0 * v - 1.0) / 3.0              # image.py:24-25


def encode_frame_float(modem, frame, rgb_u8):
    """[H, W, 3] u8 -> [H, Wc] float64 composite before the level map."""
    return modem.encode(frame, np.asarray(rgb_u8, dtype=np.uint8) / 255.0)


def encode_frame_u8(modem, frame, rgb_u8):
    return to_u8(composite_level(encode_frame_float(modem, frame, rgb_u8)))


def decode_frame_float(modem, frame, comp_u8):
    """[H, Wc] u8 -> [H, Wo, 3] float64 RGB before clipping."""
    return modem.decode(frame, composite_unlevel(np.asarray(comp_u8, dtype=np.uint8) / 255.0))


def decode_frame_u8(modem, frame, comp_u8):
    return to_u8(decode_frame_float(modem, frame, comp_u8))
