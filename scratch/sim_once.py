import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mosaic-art-rust_b200")]
import mosaic_b200
from mosaic_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
with mosaic_b200.Context(0) as ctx:
    ctx.set_material_lab(synth.lab_points(n))
    for _ in range(2):
        ctx.similarity_tri(fetch=False)
    print(ctx.stage_ms("similarity"))
