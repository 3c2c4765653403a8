"""Experiment (CPU): windowed Jacobi iteration of the greedy placement over the WHOLE grid (usage resets included).
Window = W consecutive raster cells behind the committed prefix; every sweep recomputes all of them from the previous
sweep's choices; the leading cells whose predecessors did not change are final and are committed."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "mosaic-art-rust_b200"))
import oracle as orc
from mosaic_b200 import synth

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
cfg = synth.CONFIGS[name]
gw, gh, n = cfg["grid_w"], cfg["grid_h"], cfg["n_materials"]
cap, w = cfg["max_usage"], np.float32(cfg["adjacency_weight"])
z = np.load(f"/tmp/jacobi_{name}.npz")
mlab, idx, d2, truth, fb = z["mlab"], z["idx"].astype(np.int64), z["d2"], z["truth"], z["fb"]
C = gw * gh
K = idx.shape[1]

H32 = True
nmark=[0]
first_marker=[0]
def run(W):
    first_marker[0]=0
    usage = np.zeros(n, np.int64)
    cur = np.full(C, -1, np.int64)       # committed or tentative choices
    p = 0
    sweeps = 0
    adv = []
    nfb = 0
    while p < C:
        hi = min(C, p + W)
        cells = np.arange(p, hi)
        ci = idx[p:hi]                                        # (w, K)
        prev = cur[p:hi].copy()
        # tentative choosers before each cell
        eq = (ci[:, :, None] == prev[None, None, :])          # (w, K, w): candidate == choice of window cell j
        lower = np.tril(np.ones((hi - p, hi - p), bool), -1)  # j < i
        cnt = (eq & lower[:, None, :]).sum(-1)
        valid = usage[ci] + cnt < cap
        x = cells % gw
        left = np.where(x > 0, cur[np.maximum(cells - 1, 0)], -1)
        up = np.where(cells >= gw, cur[np.maximum(cells - gw, 0)], -1)
        valid &= ci != left[:, None]
        valid &= ci != up[:, None]
        pen = np.zeros(ci.shape, np.float32)
        for nb in (left, up):
            d = mlab[ci] - mlab[np.maximum(nb, 0)][:, None, :]
            dist = np.sqrt((d * d).sum(-1, dtype=np.float32))
            pen += np.where((nb >= 0)[:, None], np.float32(1) / (dist + np.float32(1)), np.float32(0))
        total = np.where(valid, d2[p:hi] + pen * w, np.float32(np.inf))
        j = np.argmin(total, axis=1)
        new = np.where(valid.any(1), ci[np.arange(hi - p), j], -2)
        if H32:
            # the kernel's tentative evaluation only sees the first 32 candidates: a cell they cannot decide is a marker,
            # resolved (with every candidate) once it is the first cell of the window
            t32 = total[:, :32]
            m32 = t32.min(1)
            undecided = ~np.isfinite(m32) | (d2[p:hi, 32] < m32)
            undecided[0] = False   # the first cell is resolved with every candidate (serial path)
            nmark[0] += int(undecided[0:1].sum())
            new = np.where(undecided, -2, new)
            first_marker[0] += int(not np.isfinite(m32[0]) or d2[p, 32] < m32[0])
        sweeps += 1
        diff = np.nonzero(new != prev)[0]
        t = (hi - p) if len(diff) == 0 else int(diff[0]) + 1
        # the first cell is final by construction; a final -2 is a fallback cell: handled serially
        if new[0] == -2:
            assert not valid[0].any()
            usage[:] = 0
            nfb += 1
            c0 = p
            ch = -1
            for m in idx[c0]:
                if m != left[0] and m != up[0]:
                    ch = m
                    break
            if ch < 0:
                ch = idx[c0, 0]
            cur[c0] = ch
            usage[ch] += 1
            cur[p + 1:hi] = -1   # everything speculative is void after a reset
            assert cur[c0] == truth[c0]
            p += 1
            adv.append(1)
            continue
        # commit the final prefix, but stop before a tentative -2 that is not first
        stop = t
        m2 = np.nonzero(new[:t] == -2)[0]
        if len(m2):
            stop = int(m2[0])
        cur[p:hi] = new
        fin = new[:stop]
        np.add.at(usage, fin, 1)
        assert np.array_equal(cur[p:p + stop], truth[p:p + stop]), (p, stop)
        p += stop
        adv.append(stop)
    adv = np.array(adv)
    print(f"{name} W={W}: {sweeps} sweeps for {C} cells = {C / sweeps:.2f} cells/sweep, fallbacks {nfb}; "
          f"advance median {np.median(adv)}, mean {adv.mean():.2f}, share of sweeps advancing 1: {(adv == 1).mean():.2f}, serial (generic) resolutions {first_marker[0]}", flush=True)

for W in [int(a) for a in sys.argv[2:]] or [32, 64]:
    t0 = time.time()
    run(W)
    print(f"  ({time.time() - t0:.0f}s)")
