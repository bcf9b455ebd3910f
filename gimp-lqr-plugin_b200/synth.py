"""Deterministic synthetic inputs for the parity tests and bench.py.

``smooth_noise``: v_c = clamp(128 + 80*sin(2*pi*x/lambda_c)*cos(2*pi*y/mu_c) + U[-24,24]) with
(lambda, mu) = (97,61), (131,89), (173,113) for R, G, B; alpha 255 or U[128,255].  The noise is a
counter-based hash (splitmix64 of seed + pixel/channel index), so any sub-rectangle or any image of a
batch can be generated independently and identically on every host.
"""
from __future__ import annotations

import numpy as np

SEED = 0x5EA7C0DE
_LAMBDA_MU = [(97.0, 61.0), (131.0, 89.0), (173.0, 113.0)]


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _uniform_int(h: int, w: int, stream: int, seed: int, lo: int, hi: int) -> np.ndarray:
    """iid integers in [lo, hi] keyed by (seed, stream, pixel index)."""
    with np.errstate(over="ignore"):
        idx = np.arange(h * w, dtype=np.uint64).reshape(h, w)
        key = idx * np.uint64(8) + np.uint64(stream) + (np.uint64(seed) << np.uint64(32))
        r = _splitmix64(key)
    return (lo + (r % np.uint64(hi - lo + 1)).astype(np.int64)).astype(np.int64)


def smooth_noise(w: int, h: int, channels: int = 4, seed: int = SEED, alpha: str = "opaque") -> np.ndarray:
    """(h, w, channels) uint8.  channels: 1 GRAY, 2 GRAYA, 3 RGB, 4 RGBA (alpha last)."""
    colour = channels - (1 if channels in (2, 4) else 0)
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    out = np.zeros((h, w, channels), dtype=np.uint8)
    for c in range(colour):
        lam, mu = _LAMBDA_MU[c % 3]
        base = 128.0 + 80.0 * np.sin(2 * np.pi * x / lam) * np.cos(2 * np.pi * y / mu)
        v = np.rint(base) + _uniform_int(h, w, c, seed, -24, 24)
        out[:, :, c] = np.clip(v, 0, 255).astype(np.uint8)
    if channels in (2, 4):
        if alpha == "opaque":
            out[:, :, channels - 1] = 255
        elif alpha == "random":
            out[:, :, channels - 1] = _uniform_int(h, w, 7, seed, 128, 255).astype(np.uint8)
        elif alpha == "holes":  # alpha = 0 regions: energy weighting by alpha (help/en/index.wiki:48)
            a = _uniform_int(h, w, 7, seed, 0, 255)
            a[(x // 17 + y // 13) % 3 == 0] = 0
            out[:, :, channels - 1] = a.astype(np.uint8)
        else:
            raise ValueError(alpha)
    return out


def iid(w: int, h: int, channels: int = 4, seed: int = SEED) -> np.ndarray:
    out = np.zeros((h, w, channels), dtype=np.uint8)
    for c in range(channels):
        out[:, :, c] = _uniform_int(h, w, c, seed ^ 0xA5A5, 0, 255).astype(np.uint8)
    return out


def flat(w: int, h: int, channels: int = 4, value: int = 128) -> np.ndarray:
    out = np.full((h, w, channels), value, dtype=np.uint8)
    if channels in (2, 4):
        out[:, :, channels - 1] = 255
    return out


def ramp(w: int, h: int, channels: int = 4) -> np.ndarray:
    y, x = np.mgrid[0:h, 0:w]
    out = np.zeros((h, w, channels), dtype=np.uint8)
    for c in range(channels):
        out[:, :, c] = ((x * (c + 1) + y * (3 - c % 3)) % 256).astype(np.uint8)
    if channels in (2, 4):
        out[:, :, channels - 1] = 255
    return out


def ellipse_mask(w: int, h: int, fx: float = 0.30, fy: float = 0.40, channels: int = 4) -> np.ndarray:
    """Opaque white ellipse covering the centre fx*W x fy*H, transparent elsewhere (config 3 preservation mask)."""
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    inside = ((x - w / 2) / (fx * w / 2)) ** 2 + ((y - h / 2) / (fy * h / 2)) ** 2 <= 1.0
    out = np.zeros((h, w, channels), dtype=np.uint8)
    out[inside] = 255
    return out


def band_mask(w: int, h: int, x_lo: float = 0.6, x_hi: float = 0.8, channels: int = 4) -> np.ndarray:
    """Opaque white vertical band x in [x_lo*W, x_hi*W) (config 3 rigidity mask)."""
    out = np.zeros((h, w, channels), dtype=np.uint8)
    out[:, int(x_lo * w):int(x_hi * w), :] = 255
    return out
