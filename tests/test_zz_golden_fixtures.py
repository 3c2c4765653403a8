"""The committed vectors under tests/golden/ (see its README): the reference's own known answers as data, and every stage
of a small seeded pipeline.  The same checks run against two implementations of one small interface — the CPU oracle
(`-m "not gpu"`: it must meet the reference's vectors and reproduce the committed pipeline bit for bit) and the CUDA path
through the C ABI (`-m gpu`)."""
import json
import os

import numpy as np
import pytest

from mosaic_b200 import synth
from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
KNOWN = json.load(open(os.path.join(GOLDEN, "reference_known_answers.json")))


def _pipeline():
    z = np.load(os.path.join(GOLDEN, "pipeline_small.npz"), allow_pickle=False)
    cfg = {k: (float(v) if "." in v else int(v)) for k, v in z["config"]}
    return cfg, z


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


# ------------------------------------------------------------------------------------------------ two implementations
class OracleImpl:
    """the CPU restatement, stage by stage"""
    name = "oracle"

    def material_lab(self, mats):
        return orc.material_lab(mats)

    def cell_lab(self, tgt, gw, gh):
        return orc.cell_lab(tgt, gw, gh)

    def candidates(self, mat_lab, cell_lab, gw, gh, k):
        return orc.topk(mat_lab, cell_lab, k)

    def place(self, mat_lab, cell_lab, gw, gh, k, max_usage, weight):
        idx, d2 = orc.topk(mat_lab, cell_lab, k)
        placed, nfb, fb, _ = orc.place(idx, d2, mat_lab, gw, gh, max_usage, weight)
        return placed, nfb, fb

    def similarity(self, mat_lab):
        return orc.similarity_tri(mat_lab)

    def anneal(self, mat_lab, grid, gw, gh, iterations, seed):
        g, r = orc.anneal(mat_lab, grid.reshape(gh, gw), iterations, seed)
        return g.reshape(-1), r

    def optimize_greedy(self, mat_lab, grid, gw, gh, iterations, seed):
        g, r = orc.anneal_greedy(mat_lab, grid.reshape(gh, gw), iterations, seed)
        return g.reshape(-1), r

    def render(self, mats, tgt, gw, gh, placed, strength):
        return orc.render(tgt, gw, gh, mats, placed, strength)


class CudaImpl:
    """the product: libmosaic_b200.so through the ctypes mirror of the C ABI (the calls of tests/test_gpu_parity.py)"""
    name = "cuda"

    def __init__(self, ctx):
        self.ctx = ctx

    def material_lab(self, mats):
        self.ctx.set_materials(mats)
        return self.ctx.material_lab()

    def cell_lab(self, tgt, gw, gh):
        self.ctx.set_target(tgt, gw, gh)
        return self.ctx.cell_lab()

    def candidates(self, mat_lab, cell_lab, gw, gh, k):
        self.ctx.set_material_lab(mat_lab)
        self.ctx.set_cell_lab(cell_lab, gw, gh)
        return self.ctx.candidates(k)

    def place(self, mat_lab, cell_lab, gw, gh, k, max_usage, weight):
        self.ctx.set_material_lab(mat_lab)
        self.ctx.set_cell_lab(cell_lab, gw, gh)
        self.ctx.candidates(k)
        placed, nfb = self.ctx.place(max_usage, weight)
        return placed, nfb, self.ctx.fallback_cells()

    def similarity(self, mat_lab):
        self.ctx.set_material_lab(mat_lab)
        return self.ctx.similarity_tri()

    def _load_grid(self, mat_lab, grid, gw, gh):
        self.ctx.set_material_lab(mat_lab)
        self.ctx.set_cell_lab(np.zeros((gw * gh, 3), np.float32), gw, gh)
        self.ctx.set_placement(np.ascontiguousarray(grid, np.int32))

    def anneal(self, mat_lab, grid, gw, gh, iterations, seed):
        self._load_grid(mat_lab, grid, gw, gh)
        return self.ctx.anneal(iterations, seed)

    def optimize_greedy(self, mat_lab, grid, gw, gh, iterations, seed):
        self._load_grid(mat_lab, grid, gw, gh)
        return self.ctx.optimize_greedy(iterations, seed)

    def render(self, mats, tgt, gw, gh, placed, strength):
        self.ctx.set_materials(mats)
        self.ctx.set_target(tgt, gw, gh)
        self.ctx.cell_lab(fetch=False)
        self.ctx.set_placement(np.ascontiguousarray(placed, np.int32))
        return self.ctx.render(strength)


# ------------------------------------------------------------------------------------------------ the checks
def check_reference_known_answers(impl):
    for case in KNOWN["average_lab"]:
        im = case["image"]
        img = np.zeros((1, im["h"], im["w"], 3), np.uint8)
        img[:] = im["rgb"]
        lab = impl.material_lab(img)[0]
        for got, want in zip(lab, case["expect"]):
            if want is not None:
                assert abs(float(got) - want) < case["tolerance"], (impl.name, case["cite"], lab)
    for case in KNOWN["similarity"]:
        lab = np.array(case["lab"], np.float32)
        sim = impl.similarity(lab)
        assert (_bits(sim) == _bits(np.sqrt(np.array(case["derived_squares"], np.float32)))).all(), (impl.name, case["cite"])
        assert np.allclose(sim, case["derived_packed"], rtol=1e-6)
    for case in KNOWN["nearest"]:
        pts = np.array(case["points"], np.float32)                       # all of them ranked (k = n), the nearest first
        idx, d2 = impl.candidates(pts, np.array([case["query"]], np.float32), 1, 1, len(pts))
        assert idx[0, 0] == case["expect_index"] and d2[0, 0] == np.float32(case["derived_d2"]), (impl.name, case["cite"])
    for case in KNOWN["total_cost"]:
        lab = np.array(case["lab"], np.float32)
        grid = np.array(case["grid"], np.int32)
        gh, gw = grid.shape
        _, res = impl.anneal(lab, grid.reshape(-1), gw, gh, 0, 0)      # zero iterations: the energy of the grid as given
        cost = np.float32(res.initial_cost)
        if "derived_bits" in case:
            assert int(cost.view(np.uint32)) == int(case["derived_bits"], 16), (impl.name, case["cite"], cost)
        else:
            assert cost == np.float32(case["derived"]), (impl.name, case["cite"], cost)


def check_pipeline(impl, exact_lab):
    """exact_lab: the implementation's Lab means are expected to be the committed bits (the oracle) instead of within
    1e-4 of them (the CUDA path: MUFU cube root, tree-order sums)"""
    cfg, z = _pipeline()
    gw, gh, k = cfg["grid_w"], cfg["grid_h"], cfg["k"]
    mats = synth.materials(cfg["n_materials"])
    tgt = synth.target(cfg["target_w"], cfg["target_h"])
    import hashlib
    sha = [hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest() for a in (mats, tgt)]
    assert sha == list(z["inputs_sha256"]), "the synthetic input generators changed: regenerate tests/golden (make_golden.py)"

    # K1 from pixels
    for got, want in ((impl.material_lab(mats), z["material_lab"]), (impl.cell_lab(tgt, gw, gh), z["cell_lab"])):
        assert got.shape == want.shape
        if exact_lab:
            assert (_bits(got) == _bits(want)).all()
        else:
            assert (np.abs(got - want) <= np.maximum(1e-4 * np.abs(want), 1e-4)).all()
    # everything behind the Lab means, fed the COMMITTED Lab means: identical
    mlab, clab = z["material_lab"], z["cell_lab"]
    idx, d2 = impl.candidates(mlab, clab, gw, gh, k)
    assert (idx == z["cand_idx"]).all() and (_bits(d2) == _bits(z["cand_d2"])).all()
    placed, nfb, fb = impl.place(mlab, clab, gw, gh, k, cfg["max_usage"], cfg["adjacency_weight"])
    assert nfb == int(z["n_fallbacks"]) and nfb > 0                                  # the usage reset fired
    assert (np.asarray(fb) == z["fallback_cells"]).all()
    assert (placed == z["placed"]).all()
    assert (_bits(impl.similarity(mlab)) == _bits(z["similarity"])).all()
    for fn, grid_key, cost_key, count_key, iters in ((impl.anneal, "annealed", "anneal_costs", "anneal_counts", cfg["sa_iterations"]),
                                                     (impl.optimize_greedy, "greedy_grid", "greedy_costs", "greedy_counts",
                                                      cfg["greedy_iterations"])):
        grid, res = fn(mlab, z["placed"], gw, gh, iters, cfg["seed"])
        assert (np.asarray(grid).reshape(-1) == z[grid_key]).all(), grid_key
        costs = np.array([res.initial_cost, res.final_cost, res.best_cost], np.float32)
        assert (_bits(costs) == _bits(z[cost_key])).all(), (cost_key, costs, z[cost_key])
        assert [int(res.improved_count), int(res.accepted_count), int(res.iterations)] == [int(v) for v in z[count_key]], count_key
    # K4 from the committed grid
    img = impl.render(mats, tgt, gw, gh, z["annealed"], cfg["color_strength"])
    assert img.shape == z["image"].shape
    diff = np.abs(img.astype(np.int16) - z["image"].astype(np.int16))
    if exact_lab:
        assert diff.max() == 0
    else:
        assert diff.max() <= 1 and (diff > 0).mean() < 1e-3


# ------------------------------------------------------------------------------------------------ CPU: the oracle
def test_oracle_meets_the_reference_known_answers():
    check_reference_known_answers(OracleImpl())
    for case in KNOWN["average_rgb"]:
        assert tuple(orc.average_rgb(np.array(case["pixels"], np.uint8))) == tuple(case["expect"]), case["cite"]
    for case in KNOWN["packed_index"]:
        assert [orc.tri_index(i, j, case["n"]) for i, j in case["pairs"]] == case["expect"], case["cite"]
    for case in KNOWN["penalty"]:
        grid = np.array(case["grid"], np.int32)
        p = orc.penalty(np.array(case["lab"], np.float32), grid, case["position"][0], case["position"][1], case["candidate"],
                        case["weight"])
        assert p > 0 and np.float32(p) == np.float32(case["derived"]), case["cite"]
    for case in KNOWN["color_adjustment_new"]:
        assert orc.color_adjustment_new(*case["args"]).tolist() == case["expect"], case["cite"]
    for case in KNOWN["adjust_pixel_identity"]:
        out = orc.adjust_pixel(np.array(case["adjustment"], np.float32), tuple(case["pixel"]))
        assert np.abs(out.astype(int) - np.array(case["pixel"])).max() <= case["tolerance"], case["cite"]


def test_oracle_reproduces_the_committed_pipeline():
    check_pipeline(OracleImpl(), exact_lab=True)


def test_committed_pipeline_is_what_the_script_writes():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cfg, z = _pipeline()
    assert cfg == {k: v for k, v in mod.CONFIG.items()}
    fresh = mod.compute()
    for key, val in fresh.items():
        assert np.array_equal(np.asarray(val), z[key]), key


# ------------------------------------------------------------------------------------------------ GPU: the CUDA path
@pytest.mark.gpu
def test_cuda_meets_the_reference_known_answers(ctx):
    check_reference_known_answers(CudaImpl(ctx))


@pytest.mark.gpu
def test_cuda_reproduces_the_committed_pipeline(ctx):
    check_pipeline(CudaImpl(ctx), exact_lab=False)


# ------------------------------------------------------------------------------------------------ the adapter, checked on the CPU
class _StandInContext:
    """The methods of mosaic_b200.Context that CudaImpl calls, over the oracle, with the binding's shapes and return types —
    so that the adapter and the checks above are exercised on the CPU too (a mistake in the TEST must not show up for the
    first time on the GPU box).  Test infrastructure only."""

    def __init__(self):
        self.mats = self.tgt = self.mlab = self.clab = self.placed = self.cand = None
        self.gw = self.gh = 0

    def set_materials(self, mats):
        self.mats, self.mlab = np.ascontiguousarray(mats, np.uint8), None

    def material_lab(self, fetch=True):
        if self.mlab is None:
            self.mlab = orc.material_lab(self.mats)
        return self.mlab.copy() if fetch else None

    def set_target(self, tgt, gw, gh):
        self.tgt, self.gw, self.gh, self.clab = tgt, gw, gh, None

    def cell_lab(self, fetch=True):
        if self.clab is None:
            self.clab = orc.cell_lab(self.tgt, self.gw, self.gh)
        return self.clab.copy() if fetch else None

    def set_material_lab(self, lab):
        self.mlab, self.mats = np.ascontiguousarray(lab, np.float32), None

    def set_cell_lab(self, lab, gw, gh):
        lab = np.ascontiguousarray(lab, np.float32)
        assert lab.shape == (gw * gh, 3)
        self.clab, self.gw, self.gh = lab, gw, gh

    def candidates(self, k=100, fetch=True):
        self.cand = orc.topk(self.mlab, self.clab, min(k if 0 < k <= 100 else 100, len(self.mlab)))
        return self.cand

    def place(self, max_usage=3, adjacency_weight=0.3, fetch=True):
        placed, nfb, self.fb, _ = orc.place(self.cand[0], self.cand[1], self.mlab, self.gw, self.gh, max_usage, adjacency_weight)
        self.placed = placed
        return placed, nfb

    def fallback_cells(self, cap=4096):
        return self.fb

    def similarity_tri(self, fetch=True):
        return orc.similarity_tri(self.mlab)

    def set_placement(self, placed):
        placed = np.ascontiguousarray(placed, np.int32).ravel()
        assert placed.shape == (self.gw * self.gh,)
        self.placed = placed

    def anneal(self, iterations=1000, seed=0, t0=100.0, decay=0.99995, fetch=True):
        g, r = orc.anneal(self.mlab, self.placed.reshape(self.gh, self.gw), iterations, seed, t0, decay)
        self.placed = g.reshape(-1).astype(np.int32)
        return self.placed, r

    def optimize_greedy(self, max_iterations=1000, seed=0, fetch=True):
        g, r = orc.anneal_greedy(self.mlab, self.placed.reshape(self.gh, self.gw), max_iterations, seed)
        self.placed = g.reshape(-1).astype(np.int32)
        return self.placed, r

    def render(self, color_strength=0.3, out=None, fetch=True):
        return orc.render(self.tgt, self.gw, self.gh, self.mats, self.placed, color_strength)


def test_cuda_adapter_over_a_stand_in_context():
    check_reference_known_answers(CudaImpl(_StandInContext()))
    check_pipeline(CudaImpl(_StandInContext()), exact_lab=False)


def test_timed_cpu_arm_computes_the_same_bits():
    """`bench.py --impl reference` / `cpu_baseline` time the -O3 -march=native build of the restatement; with
    -ffp-contract=off and no fast-math it has to produce the strict build's results bit for bit — the timed CPU arm is the
    same computation as the parity oracle, only compiled for speed."""
    try:
        orc.select("fast")
        check_pipeline(OracleImpl(), exact_lab=True)
        check_reference_known_answers(OracleImpl())
    finally:
        orc.select("strict")
