"""GPU scratch: the two large BASELINE configs at full size, with property / sampled parity checks.
   C4: similarity database of 50,000 materials (5 GB packed matrix).
   C5: 1000x1000 grid from a 30720x17280 target vs 100,000 materials (annealing off)."""
import json, sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mosaic-art-rust_b200")]
import numpy as np
import mosaic_b200
from mosaic_b200 import synth
from oracle import oracle as orc
orc.build()
which = sys.argv[1:] or ["C4", "C5"]
out = {}

if "C4" in which:
    n = 50000
    lab = synth.lab_points(n)
    with mosaic_b200.Context(0) as ctx:
        ctx.set_material_lab(lab)
        for _ in range(3):
            ctx.similarity_tri(fetch=False)
        ms = ctx.stage_ms("similarity")
        total = n * (n - 1) // 2
        t0 = time.time(); sim = ctx.similarity_tri(); t_d2h = time.time() - t0
        t0 = time.time(); ref = orc.similarity_tri(lab, threads=os.cpu_count()); t_cpu = time.time() - t0
        equal = bool((sim.view(np.uint32) == ref.view(np.uint32)).all())
        out["C4"] = dict(n=n, entries=total, ms=round(ms, 4), achieved_gbs=round(total * 4 / ms / 1e6, 1),
                         bit_equal_all_entries=equal, cpu_ms=round(t_cpu * 1e3, 1), cpu_threads=os.cpu_count(),
                         fetch_with_d2h_ms=round(t_d2h * 1e3, 1))
        print(json.dumps(out["C4"]), flush=True)
        assert equal

if "C5" in which:
    cfg = synth.CONFIGS["C5"]
    gw, gh, n = cfg["grid_w"], cfg["grid_h"], cfg["n_materials"]
    t0 = time.time()
    mats = synth.materials(n); tgt = synth.target(cfg["target_w"], cfg["target_h"])
    t_gen = time.time() - t0
    with mosaic_b200.Context(0) as ctx:
        ctx.set_materials(mats); ctx.set_target(tgt, gw, gh)
        rep = None
        for _ in range(2):
            rep = ctx.generate(cfg["max_usage"], cfg["adjacency_weight"], cfg["color_strength"], False, 0, 42, build_similarity=False)
        st = {s: round(ctx.stage_ms(s), 3) for s in ("material_lab", "cell_lab", "topk", "place", "render")}
        mlab, clab = ctx.material_lab(), ctx.cell_lab()
        idx, d2 = ctx.candidates()
        placed, nfb = ctx.place(cfg["max_usage"], cfg["adjacency_weight"])
        # (1) sampled candidate parity: 256 cells against the oracle's brute force over all 100k materials
        rng = np.random.default_rng(5)
        sample = np.sort(rng.choice(gw * gh, 256, replace=False))
        ridx, rd2 = orc.topk(mlab, clab[sample], 100, threads=os.cpu_count())
        cand_ok = bool((ridx == idx[sample]).all() and (rd2.view(np.uint32) == d2[sample].view(np.uint32)).all())
        # (2) the whole greedy replay on the CPU from the GPU's candidate lists: bit-exact placements
        t0 = time.time()
        rplaced, rnfb, rfb, _ = orc.place(idx, d2, mlab, gw, gh, cfg["max_usage"], cfg["adjacency_weight"])
        t_place_cpu = time.time() - t0
        place_ok = bool((placed == rplaced).all() and nfb == rnfb and (ctx.fallback_cells() == rfb[:4096]).all())
        # (3) sampled Lab parity + rendered rows of 2 grid rows
        lab_ref = orc.material_lab(mats[:2000], threads=os.cpu_count())
        lab_ok = bool((np.abs(mlab[:2000] - lab_ref) <= np.maximum(1e-4 * np.abs(lab_ref), 1e-4)).all())
        img = ctx.render(cfg["color_strength"])
        th = cfg["target_h"] // gh
        rows = 2
        ref_img = orc.render(np.ascontiguousarray(tgt[:rows * th]), gw, rows, mats, placed[:rows * gw], cfg["color_strength"], threads=os.cpu_count())
        render_ok = bool(np.abs(img[:rows * th].astype(np.int16) - ref_img.astype(np.int16)).max() <= 1)
        used = np.bincount(placed, minlength=n)
        out["C5"] = dict(cells=gw * gh, materials=n, stage_ms=st, total_ms=round(sum(st.values()), 2), fallbacks=int(nfb),
                         ns_per_cell_place=round(st["place"] * 1e6 / (gw * gh), 1), candidates_sample_bit_exact=cand_ok,
                         placement_bit_exact_vs_cpu_replay=place_ok, cpu_replay_ms=round(t_place_cpu * 1e3, 1),
                         lab_sample_ok=lab_ok, render_rows_within_1lsb=render_ok, distinct_materials=int((used > 0).sum()),
                         host_input_generation_s=round(t_gen, 1))
        print(json.dumps(out["C5"]), flush=True)
        assert cand_ok and place_ok and lab_ok and render_ok
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "large.json"), "a") as f:
    f.write(json.dumps(out) + "\n")
