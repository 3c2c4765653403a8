# stage times of the 24,000-tile mosaic (C3) when the 10,000 materials have individual sizes
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mosaic-art-rust_b200")]
import numpy as np
import mosaic_b200
from mosaic_b200 import synth

n, gw, gh = 10000, 200, 120
sizes = [(64, 64), (96, 64), (48, 80), (128, 128), (40, 40), (160, 90), (38, 36)]
groups = {s: synth.materials((n + len(sizes) - 1) // len(sizes), w=s[0], h=s[1], seed=synth.SEED_MATERIALS + (k << 24)) for k, s in enumerate(sizes)}
mats = [groups[sizes[i % len(sizes)]][i // len(sizes)] for i in range(n)]
tgt = synth.target(7680, 4320)
nbytes = sum(m.size for m in mats)
with mosaic_b200.Context(0) as ctx:
    res = {}
    for name in ("ragged", "uniform64"):
        if name == "ragged":
            ctx.set_materials_ragged(mats)
        else:
            ctx.set_materials(synth.materials(n))
        ctx.set_target(tgt, gw, gh)
        best = {}
        for it in range(4):
            ctx.set_target(tgt, gw, gh)
            if name == "ragged":
                ctx.set_materials_ragged(mats)
            else:
                ctx.set_materials(synth.materials(n)) if it == 0 else None
            ctx.cell_lab(fetch=False)  # consume the pending uploads: the stage times below are device times of whole stages
            ctx.sync()
            t0 = time.perf_counter()
            rep = ctx.generate(max_usage=3, adjacency_weight=0.3, color_strength=0.3, enable_optimization=True,
                               optimization_iterations=1000, seed=42, build_similarity=True)
            ctx.sync()
            wall = (time.perf_counter() - t0) * 1e3
            cur = {s: ctx.stage_ms(s) for s in mosaic_b200.STAGES}
            cur["wall_ms"] = wall
            for k, v in cur.items():
                best[k] = min(best.get(k, 1e9), v)
        res[name] = {k: round(v, 4) for k, v in best.items()}
    res["ragged"]["material_bytes"] = int(nbytes)
    res["ragged"]["sizes"] = [list(s) for s in sizes]
    print(json.dumps(res), flush=True)
