import json, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mosaic-art-rust_b200")]
import mosaic_b200
from mosaic_b200 import synth
for name in sys.argv[1:] or ["C3"]:
    cfg = synth.CONFIGS[name]
    mats = synth.materials(cfg["n_materials"]); tgt = synth.target(cfg["target_w"], cfg["target_h"])
    with mosaic_b200.Context(0) as ctx:
        ctx.set_materials(mats); ctx.set_target(tgt, cfg["grid_w"], cfg["grid_h"])
        ts = []
        for _ in range(4):
            ctx.candidates(fetch=False); ts.append(ctx.stage_ms("topk"))
        print(json.dumps(dict(config=name, topk_ms=round(min(ts), 4))), flush=True)
