"""GPU scratch: placement time on C3 for several adjacency weights (w=0: every cell takes the bound-only path)."""
import json, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mosaic-art-rust_b200")]
import numpy as np
import mosaic_b200
from mosaic_b200 import synth
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
cfg = synth.CONFIGS[name]
mats = synth.materials(cfg["n_materials"]); tgt = synth.target(cfg["target_w"], cfg["target_h"])
with mosaic_b200.Context(0) as ctx:
    ctx.set_materials(mats); ctx.set_target(tgt, cfg["grid_w"], cfg["grid_h"])
    ctx.candidates(fetch=False)
    cells = cfg["grid_w"] * cfg["grid_h"]
    for w in (0.0, 0.05, 0.3, 1.0):
        for mu in (3, 1000):
            ts = []
            for _ in range(4):
                placed, nfb = ctx.place(mu, w)
                ts.append(ctx.stage_ms("place"))
            ms = min(ts)
            print(json.dumps(dict(config=name, w=w, max_usage=mu, ms=round(ms, 4), ns_per_cell=round(ms * 1e6 / cells, 1), cyc_per_cell=round(ms * 1e6 / cells * 1.965, 1), fallbacks=nfb)), flush=True)
