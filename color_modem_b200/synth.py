"""Deterministic synthetic RGB frames (integer-only arithmetic, so bit-identical on every platform).

Two distributions, following SURVEY.md §8d:
  'smooth' – band-limited triangle-wave gratings (periods 37–91 px) whose phase advances per frame,
             plus a few LSB of hashed noise (realistic picture content);
  'noise'  – uniform u8 noise (worst case for clipping and for the SECAM limiter).
"""
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _hash64(x):
    """splitmix64 finaliser on uint64 arrays."""
    x = x.astype(np.uint64)
    with np.errstate(over='ignore'):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return x ^ (x >> np.uint64(31))


def _tri(t, period, amp):
    """Integer triangle wave in [-amp, amp] with the given period."""
    ph = np.mod(t, period)
    half = period // 2
    up = (ph * 2 * amp) // half - amp
    down = amp - ((ph - half) * 2 * amp) // (period - half)
    return np.where(ph < half, up, down)


def synth_frames_u8(n_frames, height, width, first_frame=0, seed=0, kind='smooth'):
    """uint8 array [n_frames, height, width, 3]; frame i depends only on (seed, first_frame + i)."""
    f = (np.arange(n_frames, dtype=np.int64) + first_frame)[:, None, None, None]
    y = np.arange(height, dtype=np.int64)[None, :, None, None]
    x = np.arange(width, dtype=np.int64)[None, None, :, None]
    c = np.arange(3, dtype=np.int64)[None, None, None, :]
    key = (((f * 4099 + y) * 8209 + x) * 4 + c) + np.int64(seed) * 1000003
    h = _hash64(key.astype(np.uint64))
    if kind == 'noise':
        return (h >> np.uint64(56)).astype(np.uint8)
    if kind != 'smooth':
        raise ValueError(kind)
    px = np.array([37, 53, 91], dtype=np.int64)[None, None, None, :]
    py = np.array([91, 61, 37], dtype=np.int64)[None, None, None, :]
    t1 = x * py + y * px + f * (7 + 3 * c) * px          # diagonal grating, period px horizontally
    g1 = _tri(t1, px * py, 70)
    t2 = x * 3 - y * 5 + f * 11 + c * 40
    g2 = _tri(t2, 240 + 0 * c, 32)
    noise = ((h & np.uint64(15)).astype(np.int64) + ((h >> np.uint64(8)) & np.uint64(15)).astype(np.int64)
             + ((h >> np.uint64(16)) & np.uint64(15)).astype(np.int64) - 22) // 2
    v = 128 + g1 + g2 + noise
    return np.clip(v, 0, 255).astype(np.uint8)
