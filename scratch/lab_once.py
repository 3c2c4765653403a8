import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mosaic-art-rust_b200")]
import mosaic_b200
from mosaic_b200 import synth
mats = synth.materials(10000); tgt = synth.target(7680, 4320)
with mosaic_b200.Context(0) as ctx:
    for _ in range(2):
        ctx.set_materials(mats); ctx.set_target(tgt, 200, 120)
        ctx.material_lab(); ctx.cell_lab()
    print(ctx.stage_ms("material_lab"), ctx.stage_ms("cell_lab"))
