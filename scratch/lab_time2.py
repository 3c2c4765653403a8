import json, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mosaic-art-rust_b200")]
import numpy as np
import mosaic_b200
from mosaic_b200 import synth
from oracle import oracle as orc
mats = synth.materials(10000); tgt = synth.target(7680, 4320)
with mosaic_b200.Context(0) as ctx:
    ctx.set_materials(mats); ctx.set_target(tgt, 200, 120)
    tm, tc = [], []
    for _ in range(6):
        ctx.set_materials(mats); ctx.set_target(tgt, 200, 120)
        lab = ctx.material_lab(); tm.append(ctx.stage_ms("material_lab"))
        cl = ctx.cell_lab(); tc.append(ctx.stage_ms("cell_lab"))
    ref = orc.material_lab(mats[:3000], threads=16)
    tol = np.maximum(1e-4 * np.abs(ref), 1e-4)
    refc = orc.cell_lab(tgt, 200, 120, threads=16)
    tolc = np.maximum(1e-4 * np.abs(refc), 1e-4)
    print(json.dumps(dict(material_ms=round(min(tm), 4), material_gbs=round(mats.nbytes / min(tm) / 1e6, 1),
                          cell_ms=round(min(tc), 4), cell_gbs=round(tgt.nbytes / min(tc) / 1e6, 1),
                          mat_err_over_tol=float((np.abs(lab[:3000] - ref) / tol).max()),
                          cell_err_over_tol=float((np.abs(cl - refc) / tolc).max()))))
