"""GPU scratch: annealing time on a config."""
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
    placed, _ = ctx.place(cfg["max_usage"], cfg["adjacency_weight"])
    for iters in (0, 1000, 2000, 20000):
        ts = []
        for _ in range(3):
            ctx.set_placement(placed)
            ctx.anneal(iters, 42)
            ts.append(ctx.stage_ms("anneal"))
        print(json.dumps(dict(config=name, iterations=iters, ms=round(min(ts), 4))), flush=True)
    for t0, decay in ((100.0, 0.99995), (0.01, 0.999)):
        ctx.set_placement(placed)
        _, res = ctx.anneal(20000, 42, t0=t0, decay=decay)
        print(json.dumps(dict(t0=t0, ms=round(ctx.stage_ms("anneal"), 4), accepted=res.accepted_count)), flush=True)
