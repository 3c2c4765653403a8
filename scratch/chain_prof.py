"""GPU scratch: per-path cycle profile of the placement chain + stage times on every config."""
import json, sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mosaic-art-rust_b200")]
import numpy as np
import mosaic_b200
from mosaic_b200 import synth

for name in (sys.argv[1:] or ["C1", "C2", "C3"]):
    cfg = synth.CONFIGS[name]
    mats = synth.materials(cfg["n_materials"]); tgt = synth.target(cfg["target_w"], cfg["target_h"])
    with mosaic_b200.Context(0) as ctx:
        ctx.set_materials(mats); ctx.set_target(tgt, cfg["grid_w"], cfg["grid_h"])
        ctx.candidates(fetch=False)
        if os.environ.get('CHAIN_V'): ctx.set_chain(int(os.environ['CHAIN_V']))
        ctx.chain_profile(True)
        for _ in range(3):
            placed, nfb = ctx.place(cfg["max_usage"], cfg["adjacency_weight"])
        ms_prof = ctx.stage_ms("place")
        prof = ctx.chain_profile(False)
        for _ in range(3):
            placed2, _ = ctx.place(cfg["max_usage"], cfg["adjacency_weight"])
        ms = ctx.stage_ms("place")
        assert (placed == placed2).all()
        cells = cfg["grid_w"] * cfg["grid_h"]
        out = dict(config=name, cells=cells, ms=ms, ms_profiled=ms_prof, ns_per_cell=ms * 1e6 / cells, fallbacks=nfb, **prof)
        for k in ("fast", "exact", "generic", "fallback", "slow", "reset"):
            if prof.get("n_" + k) and ("cyc_" + k) in prof:
                out["cyc_per_" + k] = prof["cyc_" + k] / prof["n_" + k]
        print(json.dumps(out), flush=True)
