"""How many (cell, material) pairs a slab |dL| <= sqrt(tau) over materials sorted by L would leave to the top-K scan on the
24,000-tile configuration (DESIGN.md, K2 / "What comes next"): CPU only, through the oracle.
   exact tau (the k-th smallest d2 of the cell): 42 % of the materials stay  -> 58 % of the pairs skipped
   seeded tau (largest d2 from the cell to the previous cell's k materials, what the kernel has when it starts a cell):
                                                 52 % stay                    -> 48 % skipped"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mosaic-art-rust_b200")]
from mosaic_b200 import synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402

mats, tgt = synth.materials(10000), synth.target(7680, 4320)
ml, cl = orc.material_lab(mats, threads=8), orc.cell_lab(tgt, 200, 120, threads=8)
idx, d2 = orc.topk_kdtree(ml, cl, 100, threads=8)
L = np.sort(ml[:, 0])


def inside(radius):
    lo = np.searchsorted(L, cl[:, 0] - radius, "left")
    hi = np.searchsorted(L, cl[:, 0] + radius, "right")
    return (hi - lo) / len(L)


exact = inside(np.sqrt(d2[:, -1]))
prev = np.roll(idx, 1, axis=0)
seed = ((ml[prev] - cl[:, None, :]) ** 2).sum(-1).max(1)
seed[::200] = 1e9                                   # the first cell of a grid row starts unseeded
seeded = inside(np.sqrt(seed))
print(f"exact tau : {exact.mean():.3f} of the materials inside the slab (median {np.median(exact):.3f}, p95 {np.percentile(exact, 95):.3f})")
print(f"seeded tau: {seeded.mean():.3f} (median {np.median(seeded):.3f})")
