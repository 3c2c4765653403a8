import os, sys, faulthandler
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mosaic-art-rust_b200")]
import numpy as np
import mosaic_b200
from mosaic_b200 import synth
from oracle import oracle as orc
faulthandler.dump_traceback_later(40, exit=True)
n, gw, gh, cap, wgt = 300, 33, 9, 3, 0.3
yy, xx = np.mgrid[0:gh, 0:gw]
cl = np.stack([50 + 30 * np.sin(xx / 5.0), 40 * np.cos(yy / 4.0), 30 * np.sin((xx + yy) / 6.0)], -1).reshape(-1, 3).astype(np.float32)
ml = synth.lab_points(n, seed=n)
with mosaic_b200.Context(0) as ctx:
    ctx.set_chain(int(os.environ.get("CHAIN_V", "3")))
    ctx.set_material_lab(ml); ctx.set_cell_lab(cl, gw, gh)
    idx, d2 = ctx.candidates()
    print("candidates ok", flush=True)
    placed, nfb = ctx.place(cap, wgt)
    print("placed", nfb, flush=True)
    rp, rn, rfb, _ = orc.place(idx, d2, ml, gw, gh, cap, wgt)
    print("match", (placed == rp).all(), nfb, rn, np.nonzero(placed != rp)[0][:10], flush=True)
