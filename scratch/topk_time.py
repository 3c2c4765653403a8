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
    ctx.material_lab(); ctx.cell_lab()
    ts = []
    for _ in range(6):
        ctx.set_cell_lab(ctx.cell_lab(), 200, 120) if False else None
        idx, d2 = ctx.candidates(fetch=True) if _ == 0 else ctx.candidates(fetch=False)
        if _ == 0: keep = (idx, d2)
        ts.append(ctx.stage_ms("topk"))
    mlab, clab = ctx.material_lab(), ctx.cell_lab()
    sample = np.arange(0, 24000, 97)
    ridx, rd2 = orc.topk(mlab, clab[sample], 100, threads=16)
    ok = bool((ridx == keep[0][sample]).all() and (rd2.view(np.uint32) == keep[1][sample].view(np.uint32)).all())
    print(json.dumps(dict(topk_ms=round(min(ts), 4), all=[round(t, 4) for t in ts], sampled_bit_exact=ok)))
