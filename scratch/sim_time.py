import json, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mosaic-art-rust_b200")]
import numpy as np
import mosaic_b200
from mosaic_b200 import synth
for n in (10000, 50000):
    lab = synth.lab_points(n)
    with mosaic_b200.Context(0) as ctx:
        ctx.set_material_lab(lab)
        ts = []
        for _ in range(5):
            ctx.similarity_tri(fetch=False); ts.append(ctx.stage_ms("similarity"))
        total = n * (n - 1) // 2
        print(json.dumps(dict(n=n, ms=round(min(ts), 4), gbs=round(total * 4 / min(ts) / 1e6, 1))), flush=True)
