import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mosaic-art-rust_b200")]
import numpy as np
import mosaic_b200
from mosaic_b200 import synth
cfg = synth.CONFIGS["C3"]
mats = synth.materials(cfg["n_materials"]); tgt = synth.target(cfg["target_w"], cfg["target_h"])
with mosaic_b200.Context(0) as ctx:
    ctx.set_materials(mats); ctx.set_target(tgt, cfg["grid_w"], cfg["grid_h"])
    ctx.place(cfg["max_usage"], cfg["adjacency_weight"], fetch=False)
    for s in (0.3, 0.3, 0.0):
        ctx.render(s, fetch=False)
        print(s, ctx.stage_ms("render"))
