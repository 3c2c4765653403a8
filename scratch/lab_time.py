import json, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mosaic-art-rust_b200")]
import numpy as np
import mosaic_b200
from mosaic_b200 import synth
from oracle import oracle as orc
orc.build()
cfg = synth.CONFIGS["C3"]
mats = synth.materials(cfg["n_materials"]); tgt = synth.target(cfg["target_w"], cfg["target_h"])
with mosaic_b200.Context(0) as ctx:
    ctx.set_materials(mats); ctx.set_target(tgt, cfg["grid_w"], cfg["grid_h"])
    tm, tc = [], []
    for _ in range(4):
        mlab = ctx.material_lab(); tm.append(ctx.stage_ms("material_lab"))
        clab = ctx.cell_lab(); tc.append(ctx.stage_ms("cell_lab"))
    rm = orc.material_lab(mats, threads=os.cpu_count()); rc = orc.cell_lab(tgt, cfg["grid_w"], cfg["grid_h"], threads=os.cpu_count())
    def err(a, r):
        tol = np.maximum(1e-4 * np.abs(r), 1e-4)
        return float((np.abs(a - r) / tol).max()), float(np.abs(a - r).max())
    print(json.dumps(dict(material_ms=round(min(tm), 4), cell_ms=round(min(tc), 4), material_err_over_tol=err(mlab, rm), cell_err_over_tol=err(clab, rc))))
