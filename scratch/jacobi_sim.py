"""Experiment (CPU): how fast does a Jacobi fixed-point iteration of the greedy placement converge?

Every cell recomputes its choice from the choices all EARLIER cells made in the previous sweep; the prefix on which two
consecutive sweeps agree is exactly the serial result (induction over the raster order).  Prints the growth of that
prefix per sweep on a BASELINE config.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "mosaic-art-rust_b200"))
import oracle as orc  # noqa: E402
from mosaic_b200 import synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
cfg = synth.CONFIGS[name]
gw, gh, n = cfg["grid_w"], cfg["grid_h"], cfg["n_materials"]
cap, w = cfg["max_usage"], np.float32(cfg["adjacency_weight"])
cache = f"/tmp/jacobi_{name}.npz"
if os.path.exists(cache):
    z = np.load(cache)
    mlab, clab, idx, d2, truth, fb = z["mlab"], z["clab"], z["idx"], z["d2"], z["truth"], z["fb"]
else:
    mats = synth.materials(n)
    tgt = synth.target(cfg["target_w"], cfg["target_h"])
    mlab = orc.material_lab(mats, threads=8)
    clab = orc.cell_lab(tgt, gw, gh, threads=8)
    idx, d2 = orc.topk_kdtree(mlab, clab, 100, threads=8)
    truth, nfb, fb, _ = orc.place(idx, d2, mlab, gw, gh, cap, float(w))
    np.savez(cache, mlab=mlab, clab=clab, idx=idx, d2=d2, truth=truth, fb=fb)
C = gw * gh
print("fallback cells:", fb)
limit = int(fb[0]) if len(fb) else C  # simulate up to the first reset
idx = idx.astype(np.int64)
xs = np.arange(C) % gw
ys = np.arange(C) // gw
cell = np.arange(C)
INF = 1 << 40


def sweep(prev, lo, hi, usage_base=None):
    """choices of cells [lo, hi) given prev (choices of the previous sweep; cells < lo are final)."""
    # E[m]: index of the cap-th chooser among cells < hi (prev); cells >= that index may not use m
    sel = np.nonzero(prev[:hi] >= 0)[0]
    order = sel[np.argsort(prev[sel], kind="stable")]
    mats_sorted = prev[order]
    first = np.searchsorted(mats_sorted, np.arange(n), side="left")
    cnt = np.searchsorted(mats_sorted, np.arange(n), side="right") - first
    E = np.full(n, INF, np.int64)
    has = cnt >= cap
    E[has] = order[first[has] + cap - 1]
    ci = idx[lo:hi]
    i = cell[lo:hi, None]
    valid = i <= E[ci]
    left = np.where(xs[lo:hi] > 0, prev[np.maximum(cell[lo:hi] - 1, 0)], -1)
    up = np.where(ys[lo:hi] > 0, prev[np.maximum(cell[lo:hi] - gw, 0)], -1)
    valid &= ci != left[:, None]
    valid &= ci != up[:, None]
    pen = np.zeros(ci.shape, np.float32)
    for nb in (left, up):
        has_nb = nb >= 0
        d = mlab[ci] - mlab[np.maximum(nb, 0)][:, None, :]
        dist = np.sqrt((d * d).sum(-1, dtype=np.float32))
        pen += np.where(has_nb[:, None], np.float32(1.0) / (dist + np.float32(1.0)), np.float32(0))
    total = d2[lo:hi] + pen * w
    total = np.where(valid, total, np.float32(np.inf))
    j = np.argmin(total, axis=1)
    ch = ci[np.arange(hi - lo), j]
    ch = np.where(valid.any(1), ch, -2)
    return ch


mode = sys.argv[2] if len(sys.argv) > 2 else "full"
if mode == "full":
    prev = np.full(C, -1, np.int64)
    p = 0
    t0 = time.time()
    for it in range(100000):
        new = prev.copy()
        new[p:limit] = sweep(prev, p, limit)
        diff = np.nonzero(new[p:limit] != prev[p:limit])[0]
        np_ = limit if len(diff) == 0 else p + int(diff[0])
        ok = np.array_equal(new[:np_], truth[:np_])
        if it < 20 or it % 20 == 0 or np_ == limit:
            print(f"sweep {it}: verified prefix {np_} (+{np_ - p}) changed {len(diff)} truth-equal {ok} correct cells {(new[:limit] == truth[:limit]).sum()}  {time.time() - t0:.1f}s")
        prev, p = new, np_
        if p >= limit:
            break
else:
    # windowed: only [p, p+W) is swept; cells beyond the window stay unplaced (-1) and do not count
    W = int(mode)
    prev = np.full(C, -1, np.int64)
    p = 0
    sweeps = 0
    adv = []
    while p < limit:
        hi = min(limit, p + W)
        new = prev.copy()
        new[p:hi] = sweep(prev, p, hi)
        new[hi:] = -1
        diff = np.nonzero(new[p:hi] != prev[p:hi])[0]
        np_ = hi if len(diff) == 0 else p + int(diff[0])
        # a cell whose sweep result equals the previous one given an agreed prefix is final; also the first differing
        # cell itself is final (its inputs — all earlier cells — were final): advance past it
        if np_ < hi:
            np_ += 1
        sweeps += 1
        adv.append(np_ - p)
        assert np.array_equal(new[:np_], truth[:np_]), (p, np_)
        prev, p = new, np_
    adv = np.array(adv)
    print(f"W={W}: {sweeps} sweeps for {limit} cells = {limit / sweeps:.2f} cells/sweep; median {np.median(adv)}, "
          f"p10 {np.percentile(adv, 10)}, p90 {np.percentile(adv, 90)}")
