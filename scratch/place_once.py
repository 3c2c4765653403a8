"""GPU scratch: a few placements on one config (for ncu captures of the chain kernel)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mosaic-art-rust_b200")]
import mosaic_b200
from mosaic_b200 import synth
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
w = float(sys.argv[2]) if len(sys.argv) > 2 else None
cfg = synth.CONFIGS[name]
mats = synth.materials(cfg["n_materials"]); tgt = synth.target(cfg["target_w"], cfg["target_h"])
with mosaic_b200.Context(0) as ctx:
    ctx.set_materials(mats); ctx.set_target(tgt, cfg["grid_w"], cfg["grid_h"])
    ctx.candidates(fetch=False)
    for _ in range(3):
        ctx.place(cfg["max_usage"], cfg["adjacency_weight"] if w is None else w)
    print("place ms", ctx.stage_ms("place"))
