import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mosaic-art-rust_b200")]
import numpy as np
import mosaic_b200
from mosaic_b200 import synth
from oracle import oracle as orc
orc.build()
n, gw, gh, mu, w = [float(a) if "." in a else int(a) for a in sys.argv[1:6]]
mat_lab = synth.lab_points(n, seed=n)
yy, xx = np.mgrid[0:gh, 0:gw]
cell_lab = np.stack([50 + 30 * np.sin(xx / 5.0), 40 * np.cos(yy / 4.0), 30 * np.sin((xx + yy) / 6.0)], -1).reshape(-1, 3).astype(np.float32)
with mosaic_b200.Context(0) as ctx:
    ctx.set_material_lab(mat_lab); ctx.set_cell_lab(cell_lab, gw, gh)
    idx, d2 = ctx.candidates()
    placed, nfb = ctx.place(mu, w)
    rplaced, rnfb, rfb, _ = orc.place(idx, d2, mat_lab, gw, gh, mu, w)
    print("nfb", nfb, "ref", rnfb, "mismatch cells", np.nonzero(placed != rplaced)[0][:20])
    print("gpu fb", ctx.fallback_cells()[:20], "ref fb", rfb[:20])
