"""Writes tests/golden/pipeline_small.npz: the outputs of every stage of the hot path on one small seeded case.

Where the vectors come from.  The reference is a Rust crate and cannot be built or run in this image (no rustc / cargo,
crates not vendored, no network), so these vectors are NOT outputs of the reference: they are outputs of the CPU
restatement (oracle/mosaic_oracle.cpp, strict build), which tests/test_oracle_golden.py pins to every known answer the
reference's own tests hold (those literal vectors are in reference_known_answers.json next to this file).  What the
committed file adds: (1) the restatement cannot drift unnoticed — a compiler upgrade, a flag change or an edit that moves
one bit of one stage fails tests/test_zz_golden_fixtures.py on the CPU; (2) the CUDA path is compared with vectors that exist
independently of the checker's build on the GPU box.

    python tests/golden/make_golden.py            # rewrites the file; commit it together with the change that moved it

Inputs are not stored (they come from the counter-based generators of mosaic_b200/synth.py); their SHA-256 is,
so a change of the generators is told apart from a change of the arithmetic."""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "mosaic-art-rust_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

CONFIG = dict(n_materials=48, target_w=250, target_h=170, grid_w=12, grid_h=8, max_usage=1, adjacency_weight=0.3,
              color_strength=0.3, sa_iterations=400, greedy_iterations=300, seed=42, k=48)
FILE = os.path.join(HERE, "pipeline_small.npz")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def compute():
    """Every stage through the oracle, each fed with the previous stage's oracle output."""
    from mosaic_b200 import synth
    from oracle import oracle as orc

    c = CONFIG
    mats = synth.materials(c["n_materials"])
    tgt = synth.target(c["target_w"], c["target_h"])
    gw, gh = c["grid_w"], c["grid_h"]
    out = {"inputs_sha256": np.array([sha(mats), sha(tgt)])}
    out["material_lab"] = orc.material_lab(mats)
    out["cell_lab"] = orc.cell_lab(tgt, gw, gh)
    idx, d2 = orc.topk(out["material_lab"], out["cell_lab"], c["k"])
    out["cand_idx"], out["cand_d2"] = idx, d2
    placed, nfb, fb, usage = orc.place(idx, d2, out["material_lab"], gw, gh, c["max_usage"], c["adjacency_weight"])
    out["placed"], out["n_fallbacks"], out["fallback_cells"], out["usage"] = placed, np.array(nfb), fb, usage
    out["similarity"] = orc.similarity_tri(out["material_lab"])
    grid, res = orc.anneal(out["material_lab"], placed.reshape(gh, gw), c["sa_iterations"], c["seed"])
    out["annealed"] = grid.reshape(-1).astype(np.int32)
    out["anneal_costs"] = np.array([res.initial_cost, res.final_cost, res.best_cost], np.float32)
    out["anneal_counts"] = np.array([res.improved_count, res.accepted_count, res.iterations], np.uint64)
    ggrid, gres = orc.anneal_greedy(out["material_lab"], placed.reshape(gh, gw), c["greedy_iterations"], c["seed"])
    out["greedy_grid"] = ggrid.reshape(-1).astype(np.int32)
    out["greedy_costs"] = np.array([gres.initial_cost, gres.final_cost, gres.best_cost], np.float32)
    out["greedy_counts"] = np.array([gres.improved_count, gres.accepted_count, gres.iterations], np.uint64)
    out["image"] = orc.render(tgt, gw, gh, mats, out["annealed"], c["color_strength"])
    return out


def main():
    out = compute()
    np.savez_compressed(FILE, config=np.array(sorted(CONFIG.items()), dtype=object).astype(str), **out)
    print(f"{FILE}: {os.path.getsize(FILE)} bytes; fallbacks {int(out['n_fallbacks'])}, "
          f"accepted {int(out['anneal_counts'][1])} of {int(out['anneal_counts'][2])}")


if __name__ == "__main__":
    main()
