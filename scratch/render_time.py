import json, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mosaic-art-rust_b200")]
import numpy as np
import mosaic_b200
from mosaic_b200 import synth
mats = synth.materials(10000); tgt = synth.target(7680, 4320)
with mosaic_b200.Context(0) as ctx:
    ctx.set_materials(mats); ctx.set_target(tgt, 200, 120)
    ctx.cell_lab(fetch=False)
    placed, _ = ctx.place(3, 0.3)
    ts = []
    for _ in range(6):
        ctx.render(0.3, fetch=False); ts.append(ctx.stage_ms("render"))
    print(json.dumps(dict(render_ms=round(min(ts), 4), all=[round(t, 4) for t in ts])))
