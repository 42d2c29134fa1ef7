"""Seeded image pairs shared by make_reference_golden.py (runs the reference's own kernels on a GPU
box) and tests/test_reference_golden.py (checks the oracle against what they produced, on any CPU).
Pure integer arithmetic on uint64 -- no dependence on numpy's random streams."""
import numpy as np


def _hash(n, seed):
    x = np.arange(n, dtype=np.uint64) + np.uint64((seed * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)  # wraps mod 2^64
    for mul, sh in ((0xBF58476D1CE4E5B9, 30), (0x94D049BB133111EB, 27)):
        x ^= x >> np.uint64(sh)
        x = (x * np.uint64(mul)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    x ^= x >> np.uint64(31)
    return x


def noise(h, w, seed, levels=256):
    return (_hash(h * w, seed) % np.uint64(levels)).astype(np.uint8).reshape(h, w)


def ramp(h, w, seed):
    """smooth integer texture + a little noise: correlated images with a moderate score"""
    yy, xx = np.mgrid[0:h, 0:w].astype(np.int64)
    v = (xx * 3 + yy * 5 + (xx * yy) // 97 + seed * 11) % 256
    n = (_hash(h * w, seed + 1000) % np.uint64(9)).astype(np.int64).reshape(h, w) - 4
    return np.clip(v + n, 0, 255).astype(np.uint8)


def pairs():
    """name -> (render top-down, warped), all u8"""
    a = ramp(96, 160, 1)
    out = {
        "noise_160x96": (noise(96, 160, 11), noise(96, 160, 12)),
        "correlated_160x96": (a, np.clip(a.astype(np.int64) // 2 + 60 + (noise(96, 160, 13, 7).astype(np.int64) - 3), 0, 255).astype(np.uint8)),
        "identical_160x96": (a, a.copy()),
        "background_160x96": (np.where(noise(96, 160, 14, 3) == 0, 255, a).astype(np.uint8), np.where(noise(96, 160, 15, 4) == 0, 0, ramp(96, 160, 2)).astype(np.uint8)),
        "ragged_101x37": (noise(37, 101, 16), ramp(37, 101, 3)),
        "four_levels_64x48": ((noise(48, 64, 17, 4) * 80).astype(np.uint8), (noise(48, 64, 18, 2) * 200).astype(np.uint8)),
        "constant_64x48": (np.full((48, 64), 255, np.uint8), np.full((48, 64), 7, np.uint8)),
        "noise_752x480": (noise(480, 752, 19), noise(480, 752, 20)),
        "nearly_independent_752x480": (ramp(480, 752, 4), np.ascontiguousarray(ramp(480, 752, 5)[::-1, ::-1])),
    }
    return out
