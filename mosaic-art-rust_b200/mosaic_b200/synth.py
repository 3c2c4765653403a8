"""Counter-based synthetic inputs: identical bytes wherever they are generated.

h(k) = splitmix64(seed + k).  Materials: a base colour uniform over the RGB cube plus +-32 texture
noise; target: smooth sin/cos colour fields plus +-24 noise, so that usage caps, the adjacency
penalty and the fallback/reset path of the greedy matcher are all exercised.
"""
from __future__ import annotations

import numpy as np

SEED_TARGET = 0x5EED0001
SEED_MATERIALS = 0x5EED0002
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    """Vectorised splitmix64 finaliser of (already seed-offset) counters; uint64 in, uint64 out."""
    with np.errstate(over="ignore"):
        z = (x.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def materials(n: int, w: int = 64, h: int = 64, seed: int = SEED_MATERIALS, start: int = 0) -> np.ndarray:
    """(n, h, w, 3) uint8.  Material i = start + row index."""
    out = np.empty((n, h, w, 3), np.uint8)
    px = np.arange(h * w, dtype=np.uint64)
    chunk = 2048
    for b in range(0, n, chunk):
        e = min(n, b + chunk)
        ids = np.arange(start + b, start + e, dtype=np.uint64)
        with np.errstate(over="ignore"):
            base = splitmix64(np.uint64(seed) + ids)  # 3 low bytes = base colour
            noise = splitmix64(np.uint64(seed) + (ids[:, None] * np.uint64(h * w) + px[None, :])
                               + np.uint64(0x100000000))
        for c in range(3):
            bc = ((base >> np.uint64(8 * c)) & np.uint64(255)).astype(np.int32)[:, None]
            nz = ((noise >> np.uint64(8 * c)) & np.uint64(63)).astype(np.int32) - 32
            out[b:e, :, :, c] = np.clip(bc + nz, 0, 255).astype(np.uint8).reshape(e - b, h, w)
    return out


def target(w: int, h: int, seed: int = SEED_TARGET) -> np.ndarray:
    """(h, w, 3) uint8 smooth colour field with texture noise."""
    out = np.empty((h, w, 3), np.uint8)
    params = ((3.0, 2.0, 0.0), (5.0, 3.0, 1.0 / 3.0), (2.0, 7.0, 2.0 / 3.0))
    xs = np.arange(w, dtype=np.float64)
    rows = max(1, (1 << 22) // max(w, 1))
    for y0 in range(0, h, rows):
        y1 = min(h, y0 + rows)
        ys = np.arange(y0, y1, dtype=np.float64)
        with np.errstate(over="ignore"):
            k = (np.arange(y0, y1, dtype=np.uint64)[:, None] * np.uint64(w) + np.arange(w, dtype=np.uint64)[None, :])
            hv = splitmix64(np.uint64(seed) + k)
        for c, (fx, fy, ph) in enumerate(params):
            field = 127.5 + 80.0 * np.sin(2.0 * np.pi * (fx * xs / w + ph))[None, :] * \
                np.cos(2.0 * np.pi * fy * ys / h)[:, None]
            nz = ((hv >> np.uint64(16 * c)) & np.uint64(0xFFFF)).astype(np.float64) * (48.0 / 65536.0) - 24.0
            out[y0:y1, :, c] = np.floor(np.clip(field + nz, 0.0, 255.0)).astype(np.uint8)
    return out


def lab_points(n: int, seed: int = 0x5EED0004) -> np.ndarray:
    """(n, 3) float32 Lab triples, L in [0,100], a,b in [-100,100] (C4 similarity-database input)."""
    with np.errstate(over="ignore"):
        hv = splitmix64(np.uint64(seed) + np.arange(n, dtype=np.uint64))
    u = [((hv >> np.uint64(21 * c)) & np.uint64(0x1FFFFF)).astype(np.float64) / float(1 << 21) for c in range(3)]
    return np.stack([100.0 * u[0], 200.0 * u[1] - 100.0, 200.0 * u[2] - 100.0], axis=1).astype(np.float32)


# BASELINE.json configs: name -> parameters
CONFIGS = {
    "C1": dict(target_w=1920, target_h=1080, n_materials=500, grid_w=50, grid_h=28, max_usage=3,
               adjacency_weight=0.3, color_strength=0.3, sa_iterations=0, seed=42),
    "C2": dict(target_w=1920, target_h=1080, n_materials=1500, grid_w=80, grid_h=60, max_usage=3,
               adjacency_weight=0.25, color_strength=0.4, sa_iterations=2000, seed=42),
    "C3": dict(target_w=7680, target_h=4320, n_materials=10000, grid_w=200, grid_h=120, max_usage=3,
               adjacency_weight=0.3, color_strength=0.3, sa_iterations=1000, seed=42),
    "C5": dict(target_w=30720, target_h=17280, n_materials=100000, grid_w=1000, grid_h=1000, max_usage=3,
               adjacency_weight=0.3, color_strength=0.3, sa_iterations=0, seed=42),
}
