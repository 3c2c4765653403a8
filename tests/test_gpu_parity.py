"""GPU parity tests: every stage of the CUDA path, called through the C ABI, against the CPU oracle.

Stage-wise protocol: Lab means within 1e-4 relative; candidate lists, placements,
annealed grids and similarity entries bit-exact GIVEN IDENTICAL INPUTS (the oracle is fed the GPU's
Lab means); output pixels <= 1 LSB given identical placements.
"""
import numpy as np
import pytest

from oracle import oracle as orc
from mosaic_b200 import synth

pytestmark = pytest.mark.gpu

LAB_RTOL = 1e-4   # north star: Lab means <= 1e-4 relative ...
LAB_ATOL = 1e-4   # ... with an absolute floor of 1e-4 for a/b near zero


def assert_lab_close(gpu, ref):
    tol = np.maximum(LAB_RTOL * np.abs(ref), LAB_ATOL)
    err = np.abs(gpu - ref)
    assert (err <= tol).all(), f"max err {err.max()} at {np.unravel_index(err.argmax(), err.shape)}"


# ---------------------------------------------------------------- K1 lab_mean
def test_material_lab_matches_oracle(ctx):
    mats = synth.materials(257)
    ctx.set_materials(mats)
    assert_lab_close(ctx.material_lab(), orc.material_lab(mats, threads=8))


def test_material_lab_reference_goldens(ctx):
    # lib.rs:133-158: solid red -> (53.24, 80.09, 67.20) +-1.0, gray128 -> L 53.59 +-1.0, |a|,|b| < 1
    cols = [(255, 0, 0), (128, 128, 128), (0, 255, 0), (0, 0, 255), (255, 255, 255), (0, 0, 0)]
    mats = np.zeros((len(cols), 10, 10, 3), np.uint8)
    for i, c in enumerate(cols):
        mats[i] = c
    ctx.set_materials(mats)
    lab = ctx.material_lab()
    assert abs(lab[0, 0] - 53.24) < 1.0 and abs(lab[0, 1] - 80.09) < 1.0 and abs(lab[0, 2] - 67.20) < 1.0
    assert abs(lab[1, 0] - 53.59) < 1.0 and abs(lab[1, 1]) < 1.0 and abs(lab[1, 2]) < 1.0
    exp = np.array([[53.2408, 80.09243, 67.20321], [53.585022, 0, 0], [87.73472, -86.182686, 83.17931],
                    [32.29701, 79.187546, -107.86017], [100, 0, 0], [0, 0, 0]], np.float32)
    assert_lab_close(lab, exp)


@pytest.mark.parametrize("w,h", [(64, 64), (7, 5), (33, 17), (128, 96), (1, 1)])
def test_material_lab_odd_sizes(ctx, w, h):
    mats = synth.materials(19, w=w, h=h)
    ctx.set_materials(mats)
    assert_lab_close(ctx.material_lab(), orc.material_lab(mats))


@pytest.mark.parametrize("w,h,gw,gh", [(480, 288, 24, 16), (250, 130, 7, 3), (97, 61, 5, 4), (64, 64, 64, 64),
                                        (300, 200, 3, 2)])
def test_cell_lab_matches_oracle(ctx, w, h, gw, gh):
    tgt = synth.target(w, h)
    ctx.set_target(tgt, gw, gh)
    assert_lab_close(ctx.cell_lab(), orc.cell_lab(tgt, gw, gh, threads=8))


# ---------------------------------------------------------------- K2 nn_topk
@pytest.mark.parametrize("n,c", [(300, 384), (5000, 200), (100, 50), (37, 33), (1, 4), (2049, 70), (4100, 129)])
def test_candidates_bit_exact(ctx, n, c):
    mat_lab = synth.lab_points(n)
    cell_lab = (synth.lab_points(c, seed=99) * np.float32(0.9)).astype(np.float32)
    if n >= 300:  # exact duplicates and exact ties in distance
        mat_lab[10] = mat_lab[3]
        mat_lab[200] = mat_lab[3]
        cell_lab[5] = mat_lab[3]
    ctx.set_material_lab(mat_lab)
    ctx.set_cell_lab(cell_lab, c, 1)
    idx, d2 = ctx.candidates()
    ridx, rd2 = orc.topk(mat_lab, cell_lab, 100, threads=8)
    assert idx.shape == ridx.shape
    assert (idx == ridx).all()
    assert (d2.view(np.uint32) == rd2.view(np.uint32)).all()


def test_candidates_ties_quantised(ctx):
    # heavy ties: Lab values on a coarse lattice -> many equal distances; order must be (d2, index)
    rng = np.random.default_rng(5)
    mat_lab = rng.integers(0, 6, (3000, 3)).astype(np.float32) * 10
    cell_lab = rng.integers(0, 6, (256, 3)).astype(np.float32) * 10
    ctx.set_material_lab(mat_lab)
    ctx.set_cell_lab(cell_lab, 16, 16)
    idx, d2 = ctx.candidates()
    ridx, rd2 = orc.topk(mat_lab, cell_lab, 100, threads=8)
    assert (idx == ridx).all() and (d2.view(np.uint32) == rd2.view(np.uint32)).all()


# ---------------------------------------------------------------- K2b greedy placement
@pytest.mark.parametrize("n,gw,gh,max_usage,weight", [
    (300, 24, 16, 3, 0.3),     # capacity 900 > 384
    (100, 24, 16, 3, 0.3),     # capacity 300 < 384 -> fallback/reset fires
    (100, 24, 16, 3, 0.0),     # no adjacency calculator (main.rs:355-362)
    (40, 20, 20, 1, 1.0),      # k = n < 100, many resets
    (500, 50, 28, 3, 0.3),     # C1 shape
    (5, 9, 7, 0, 0.3),         # max_usage 0 -> every cell takes the fallback
    (1, 4, 4, 3, 0.3),         # single material: final nearest_one fallback
    (300, 1, 40, 2, 0.5),      # one-column grid (no Left)
    (300, 40, 1, 2, 0.5),      # one-row grid (no Up)
    # grid_w >= 33: the three-role fast kernel (producers / consumer / committer)
    (300, 40, 12, 3, 0.3),
    (100, 64, 10, 3, 0.0),     # no adjacency calculator on the fast kernel
    (40, 33, 9, 1, 1.0),       # k = n < 100, cap 1: constant resets
    (5, 36, 4, 0, 0.3),        # max_usage 0
    (1, 40, 3, 3, 0.3),        # single material
    (300, 200, 3, 2, 0.5),
    (2000, 120, 40, 3, 0.3),
    (300, 48, 20, 300, 0.3),   # 16-bit usage counters
    (300, 48, 20, 70000, 0.3), # cap > 65535: generic kernel with 32-bit counters
])
def test_place_bit_exact(ctx, n, gw, gh, max_usage, weight):
    mat_lab = synth.lab_points(n, seed=n)
    # smooth cell field: neighbours want the same materials
    yy, xx = np.mgrid[0:gh, 0:gw]
    cell_lab = np.stack([50 + 30 * np.sin(xx / 5.0), 40 * np.cos(yy / 4.0), 30 * np.sin((xx + yy) / 6.0)], -1)
    cell_lab = cell_lab.reshape(-1, 3).astype(np.float32)
    ctx.set_material_lab(mat_lab)
    ctx.set_cell_lab(cell_lab, gw, gh)
    idx, d2 = ctx.candidates()
    placed, nfb = ctx.place(max_usage, weight)
    rplaced, rnfb, rfb, _ = orc.place(idx, d2, mat_lab, gw, gh, max_usage, weight)
    assert nfb == rnfb
    assert (ctx.fallback_cells() == rfb).all()
    assert (placed == rplaced).all()


@pytest.mark.parametrize("n,gw,gh", [(300, 24, 16), (300, 48, 20), (100, 40, 12)])
def test_loaded_database_colours_drive_the_similarities(ctx, n, gw, gh):
    """main.rs:120-143: with a database loaded from a file, matching uses the tiles' Lab means but every similarity
    (matrix, adjacency penalty, annealing energy) comes from the colours stored in the file."""
    W = 200.0  # a weight at which the penalty decides between candidates, so that the two colour sets give different mosaics
    mat_lab = synth.lab_points(n, seed=n)
    db_lab = (mat_lab + np.random.default_rng(5).normal(0, 6.0, mat_lab.shape)).astype(np.float32)  # a stale file
    yy, xx = np.mgrid[0:gh, 0:gw]
    cell_lab = np.stack([50 + 30 * np.sin(xx / 5.0), 40 * np.cos(yy / 4.0), 30 * np.sin((xx + yy) / 6.0)], -1)
    cell_lab = cell_lab.reshape(-1, 3).astype(np.float32)
    ctx.set_material_lab(mat_lab)
    ctx.set_cell_lab(cell_lab, gw, gh)
    ctx.set_similarity_lab(db_lab)
    idx, d2 = ctx.candidates()
    ridx, rd2 = orc.topk(mat_lab, cell_lab, 100)                      # candidates: the tiles' colours
    assert (idx == ridx).all() and (d2.view(np.uint32) == rd2.view(np.uint32)).all()
    placed, nfb = ctx.place(3, W)
    rplaced, rnfb, rfb, _ = orc.place(idx, d2, db_lab, gw, gh, 3, W)  # penalties: the database's colours
    assert nfb == rnfb and (placed == rplaced).all()
    plain, _, _, _ = orc.place(idx, d2, mat_lab, gw, gh, 3, W)
    assert (plain != rplaced).any(), "the fixture must make the two colour sets disagree"
    assert (ctx.similarity_tri().view(np.uint32) == orc.similarity_tri(db_lab).view(np.uint32)).all()
    out, res = ctx.anneal(1500, 11)
    rgrid, rres = orc.anneal(db_lab, rplaced.reshape(gh, gw), 1500, 11)
    assert (out.reshape(gh, gw) == rgrid).all()
    assert np.float32(res.final_cost).view(np.uint32) == np.float32(rres.final_cost).view(np.uint32)
    ctx.set_similarity_lab(None)                                        # back to the tiles' own colours
    placed2, _ = ctx.place(3, W)
    assert (placed2 == plain).all()


def test_place_randomized_against_the_replay(ctx):
    """Random shapes, caps, weights and colour fields through the chain kernel (hot blocks, contended cells, the
    multi-chunk path, usage resets, narrow-grid kernel) against the oracle's replay of main.rs:335-475."""
    rng = np.random.default_rng(20261017)
    for trial in range(120):
        n = int(rng.choice([1, 2, 7, 31, 33, 64, 100, 150, 400, 1200]))
        gw = int(rng.integers(1, 90))
        gh = int(rng.integers(1, 14))
        max_usage = int(rng.choice([0, 1, 1, 2, 3, 3, 5, 40, 300]))
        weight = float(rng.choice([0.0, 0.05, 0.3, 0.3, 1.0, 2.5]))
        mat_lab = synth.lab_points(n, seed=1000 + trial)
        if trial % 4 == 0:    # clustered materials: many near-ties and contended cells
            mat_lab = (mat_lab[rng.integers(0, max(1, n // 8), n)] + rng.normal(0, 0.05, (n, 3))).astype(np.float32)
        yy, xx = np.mgrid[0:gh, 0:gw]
        smooth = np.stack([50 + 30 * np.sin(xx / 5.0 + trial), 40 * np.cos(yy / 4.0), 30 * np.sin((xx + yy) / 6.0)], -1)
        noise = rng.normal(0, float(rng.choice([0.0, 0.5, 20.0])), smooth.shape)
        cell_lab = (smooth + noise).reshape(-1, 3).astype(np.float32)
        ctx.set_material_lab(mat_lab)
        ctx.set_cell_lab(cell_lab, gw, gh)
        idx, d2 = ctx.candidates()
        placed, nfb = ctx.place(max_usage, weight)
        rplaced, rnfb, rfb, _ = orc.place(idx, d2, mat_lab, gw, gh, max_usage, weight)
        tag = f"trial {trial}: n={n} grid={gw}x{gh} cap={max_usage} w={weight}"
        assert nfb == rnfb, tag
        assert (placed == rplaced).all(), tag
        assert (ctx.fallback_cells() == rfb[:4096]).all(), tag


def test_place_full_pipeline_c1(ctx):
    cfg = synth.CONFIGS["C1"]
    mats = synth.materials(cfg["n_materials"])
    tgt = synth.target(cfg["target_w"], cfg["target_h"])
    ctx.set_materials(mats)
    ctx.set_target(tgt, cfg["grid_w"], cfg["grid_h"])
    mlab, clab = ctx.material_lab(), ctx.cell_lab()
    idx, d2 = ctx.candidates()
    ridx, rd2 = orc.topk(mlab, clab, 100, threads=8)  # oracle fed with the GPU's Lab means
    assert (idx == ridx).all() and (d2.view(np.uint32) == rd2.view(np.uint32)).all()
    placed, nfb = ctx.place(cfg["max_usage"], cfg["adjacency_weight"])
    rplaced, rnfb, rfb, _ = orc.place(ridx, rd2, mlab, cfg["grid_w"], cfg["grid_h"], cfg["max_usage"],
                                      cfg["adjacency_weight"])
    assert nfb == rnfb and (placed == rplaced).all()


# ---------------------------------------------------------------- K3 similarity
@pytest.mark.parametrize("n", [2, 3, 4, 17, 500, 1500, 2311])
def test_similarity_bit_equal(ctx, n):
    lab = synth.lab_points(n, seed=n + 1)
    ctx.set_material_lab(lab)
    sim = ctx.similarity_tri()
    ref = orc.similarity_tri(lab, threads=8)
    assert sim.shape == ref.shape
    assert (sim.view(np.uint32) == ref.view(np.uint32)).all()


def test_similarity_reference_fixtures(ctx):
    # similarity.rs:190-192 fixtures; App. B: sqrt(300), sqrt(1200); packed index n=3
    lab = np.array([[50, 0, 0], [60, 10, 10], [40, -10, -10]], np.float32)
    ctx.set_material_lab(lab)
    sim = ctx.similarity_tri()
    assert sim[0] == np.float32(17.320509) and sim[1] == np.float32(17.320509) and sim[2] == np.float32(34.641018)


# ---------------------------------------------------------------- K5 anneal
@pytest.mark.parametrize("n,gw,gh,iters,seed", [(300, 24, 16, 2000, 42), (50, 8, 8, 5000, 0), (300, 2, 2, 500, 7),
                                                 (300, 1, 1, 100, 1), (300, 1, 30, 1000, 3), (20, 30, 20, 3000, 9)])
def test_anneal_bit_exact(ctx, n, gw, gh, iters, seed):
    mat_lab = synth.lab_points(n, seed=n + 3)
    rng = np.random.default_rng(seed)
    grid = rng.integers(0, n, (gh, gw)).astype(np.int32)
    ctx.set_material_lab(mat_lab)
    ctx.set_cell_lab(np.zeros((gw * gh, 3), np.float32), gw, gh)
    ctx.set_placement(grid)
    placed, res = ctx.anneal(iters, seed)
    rgrid, rres = orc.anneal(mat_lab, grid, iters, seed)
    assert (placed.reshape(gh, gw) == rgrid).all()
    for f in ("initial_cost", "final_cost", "best_cost"):
        assert np.float32(getattr(res, f)).view(np.uint32) == np.float32(getattr(rres, f)).view(np.uint32), f
    assert res.accepted_count == rres.accepted_count and res.improved_count == rres.improved_count
    assert res.iterations == iters


@pytest.mark.parametrize("n,gw,gh,iters,seed,holes", [(300, 24, 16, 4000, 42, 0.0), (50, 8, 8, 5000, 0, 0.0), (300, 2, 2, 500, 7, 0.0),
                                                       (300, 1, 1, 100, 1, 0.0), (20, 30, 20, 6000, 9, 0.1), (6, 40, 33, 20000, 5, 0.05)])
def test_optimize_greedy_bit_exact(ctx, n, gw, gh, iters, seed, holes):
    """MosaicOptimizer::optimize_greedy (optimizer.rs:132-198): improvements only, four position draws per iteration."""
    mat_lab = synth.lab_points(n, seed=n + 5)
    rng = np.random.default_rng(seed)
    grid = rng.integers(0, n, (gh, gw)).astype(np.int32)
    grid[rng.random((gh, gw)) < holes] = -1
    ctx.set_material_lab(mat_lab)
    ctx.set_cell_lab(np.zeros((gw * gh, 3), np.float32), gw, gh)
    ctx.set_placement(grid)
    placed, res = ctx.optimize_greedy(iters, seed)
    rgrid, rres = orc.anneal_greedy(mat_lab, grid, iters, seed)
    assert (placed.reshape(gh, gw) == rgrid).all()
    for f in ("initial_cost", "final_cost", "best_cost"):
        assert np.float32(getattr(res, f)).view(np.uint32) == np.float32(getattr(rres, f)).view(np.uint32), f
    assert res.accepted_count == rres.accepted_count == res.improved_count == rres.improved_count
    assert res.iterations == iters
    assert res.best_cost == res.final_cost and res.final_cost <= res.initial_cost
    # a second call continues from the improved grid (the context's placement is the working grid)
    placed2, res2 = ctx.optimize_greedy(iters, seed + 1)
    rgrid2, rres2 = orc.anneal_greedy(mat_lab, rgrid, iters, seed + 1)
    assert (placed2.reshape(gh, gw) == rgrid2).all() and res2.accepted_count == rres2.accepted_count


def test_anneal_with_empty_cells_and_duplicates(ctx):
    # empty cells are skipped without cooling (optimizer.rs:71-73); equal materials give delta 0 (adjacency.rs:156-158)
    n, gw, gh = 6, 12, 9
    mat_lab = synth.lab_points(n, seed=2)
    mat_lab[4] = mat_lab[1]   # two materials with identical Lab: sim 0 between different indices
    rng = np.random.default_rng(8)
    grid = rng.integers(0, n, (gh, gw)).astype(np.int32)
    grid[rng.random((gh, gw)) < 0.15] = -1
    ctx.set_material_lab(mat_lab)
    ctx.set_cell_lab(np.zeros((gw * gh, 3), np.float32), gw, gh)
    for t0, decay in ((100.0, 0.99995), (0.05, 0.9995)):
        ctx.set_placement(grid)
        placed, res = ctx.anneal(3000, 11, t0=t0, decay=decay)
        rgrid, rres = orc.anneal(mat_lab, grid, 3000, 11, t0=t0, decay=decay)
        assert (placed.reshape(gh, gw) == rgrid).all()
        assert res.accepted_count == rres.accepted_count and res.improved_count == rres.improved_count
        for f in ("initial_cost", "final_cost", "best_cost"):
            assert np.float32(getattr(res, f)).view(np.uint32) == np.float32(getattr(rres, f)).view(np.uint32), f


@pytest.mark.parametrize("t0,decay", [(1.0, 0.999), (0.2, 0.9999), (0.003, 1.0), (5.0, 0.99)])
def test_anneal_temperature_sweep(ctx, t0, decay):
    # acceptance probabilities from ~1 down to ~0: exercises both exp bounds and the exact exp in between
    n, gw, gh = 500, 40, 30
    mat_lab = synth.lab_points(n, seed=21)
    grid = np.random.default_rng(2).integers(0, n, (gh, gw)).astype(np.int32)
    ctx.set_material_lab(mat_lab)
    ctx.set_cell_lab(np.zeros((gw * gh, 3), np.float32), gw, gh)
    ctx.set_placement(grid)
    placed, res = ctx.anneal(6000, 77, t0=t0, decay=decay)
    rgrid, rres = orc.anneal(mat_lab, grid, 6000, 77, t0=t0, decay=decay)
    assert (placed.reshape(gh, gw) == rgrid).all()
    assert res.accepted_count == rres.accepted_count and res.improved_count == rres.improved_count
    assert np.float32(res.final_cost).view(np.uint32) == np.float32(rres.final_cost).view(np.uint32)


def test_anneal_large_grid_in_hbm(ctx):
    # 300x300 = 90,000 cells x 2 B + ring does not fit with the Lab table -> grid-only variant; 70,000 materials -> i32 grid
    n, gw, gh = 70000, 300, 300
    mat_lab = synth.lab_points(n, seed=5)
    grid = np.random.default_rng(4).integers(0, n, (gh, gw)).astype(np.int32)
    ctx.set_material_lab(mat_lab)
    ctx.set_cell_lab(np.zeros((gw * gh, 3), np.float32), gw, gh)
    ctx.set_placement(grid)
    placed, res = ctx.anneal(3000, 3)
    rgrid, rres = orc.anneal(mat_lab, grid, 3000, 3)
    assert (placed.reshape(gh, gw) == rgrid).all()
    assert np.float32(res.final_cost).view(np.uint32) == np.float32(rres.final_cost).view(np.uint32)
    assert np.float32(res.initial_cost).view(np.uint32) == np.float32(rres.initial_cost).view(np.uint32)


def test_anneal_cold_rejects(ctx):
    # low temperature: the acceptance test (exp + uniform draw) actually rejects -> exercises the RNG stream
    n, gw, gh = 200, 16, 12
    mat_lab = synth.lab_points(n, seed=11)
    grid = np.random.default_rng(1).integers(0, n, (gh, gw)).astype(np.int32)
    ctx.set_material_lab(mat_lab)
    ctx.set_cell_lab(np.zeros((gw * gh, 3), np.float32), gw, gh)
    ctx.set_placement(grid)
    placed, res = ctx.anneal(4000, 5, t0=0.01, decay=0.999)
    rgrid, rres = orc.anneal(mat_lab, grid, 4000, 5, t0=0.01, decay=0.999)
    assert 0 < rres.accepted_count < 4000
    assert (placed.reshape(gh, gw) == rgrid).all()
    assert res.accepted_count == rres.accepted_count and res.improved_count == rres.improved_count
    assert np.float32(res.final_cost).view(np.uint32) == np.float32(rres.final_cost).view(np.uint32)


def test_anneal_randomized(ctx):
    """Random grids, material counts, temperatures and seeds: grid, counters and cost bits against the oracle."""
    rng = np.random.default_rng(7)
    for trial in range(40):
        n = int(rng.choice([2, 3, 17, 64, 300, 2000]))
        gw, gh = int(rng.integers(1, 70)), int(rng.integers(1, 40))
        iters = int(rng.choice([0, 1, 50, 700, 3000]))
        t0 = float(rng.choice([100.0, 1.0, 0.05, 0.002]))
        decay = float(rng.choice([0.99995, 0.999, 0.9]))
        seed = int(rng.integers(0, 2**62))
        mat_lab = synth.lab_points(n, seed=trial)
        grid = rng.integers(0, n, (gh, gw)).astype(np.int32)
        if trial % 3 == 0:
            grid[rng.random((gh, gw)) < 0.1] = -1            # empty cells (optimizer.rs:71-73)
        ctx.set_material_lab(mat_lab)
        ctx.set_cell_lab(np.zeros((gw * gh, 3), np.float32), gw, gh)
        ctx.set_placement(grid)
        placed, res = ctx.anneal(iters, seed, t0=t0, decay=decay)
        rgrid, rres = orc.anneal(mat_lab, grid, iters, seed, t0=t0, decay=decay)
        tag = f"trial {trial}: n={n} grid={gw}x{gh} iters={iters} t0={t0} decay={decay}"
        assert (placed.reshape(gh, gw) == rgrid).all(), tag
        assert (res.accepted_count, res.improved_count) == (rres.accepted_count, rres.improved_count), tag
        for a, b in ((res.initial_cost, rres.initial_cost), (res.final_cost, rres.final_cost), (res.best_cost, rres.best_cost)):
            assert np.float32(a).view(np.uint32) == np.float32(b).view(np.uint32), tag


# ---------------------------------------------------------------- K4 render
@pytest.mark.parametrize("w,h,gw,gh,strength", [(480, 288, 24, 16, 0.3), (480, 288, 24, 16, 0.0), (250, 130, 7, 3, 1.0),
                                                 (512, 512, 8, 8, 0.4), (1920, 1080, 50, 28, 0.3),
                                                 (760, 432, 20, 12, 0.3),   # 38x36 tiles (the 24,000-tile shape)
                                                 (760, 432, 20, 12, 0.0),
                                                 (800, 450, 10, 5, 0.5),    # 80x90 tiles: upscaling, 7-tap windows
                                                 (768, 320, 12, 10, 0.3),   # 64x32 tiles: horizontal axis skipped
                                                 (900, 510, 30, 30, 0.3),   # 30x17 tiles (the giant-mosaic shape): 15/25 taps
                                                 (640, 640, 40, 32, 0.6)])  # 16x20 tiles: 25/21 taps
def test_render_within_one_lsb(ctx, w, h, gw, gh, strength):
    n = 64
    mats = synth.materials(n)
    tgt = synth.target(w, h)
    placed = np.random.default_rng(3).integers(0, n, gw * gh).astype(np.int32)
    placed[5] = -1  # an empty cell stays black
    ctx.set_materials(mats)
    ctx.set_target(tgt, gw, gh)
    ctx.cell_lab(fetch=False)
    ctx.set_placement(placed)
    out = ctx.render(strength)
    ref = orc.render(tgt, gw, gh, mats, placed, strength, threads=8)
    assert out.shape == ref.shape
    diff = np.abs(out.astype(np.int16) - ref.astype(np.int16))
    assert diff.max() <= 1, f"max diff {diff.max()}"
    # tighter: the arithmetic is restated operation by operation, so in practice it is bit-exact
    assert (diff > 0).mean() < 1e-3


def test_render_randomized_shapes(ctx):
    """Random material and tile sizes (all kernel paths: dot-product with 4/5/7-word windows, per-cell fast and
    generic kernels, skipped axes) and strengths, <= 1 LSB against the oracle."""
    rng = np.random.default_rng(11)
    for trial in range(24):
        mw, mh = int(rng.choice([16, 20, 32, 48, 64, 64, 96])), int(rng.choice([16, 24, 32, 64, 64, 80]))
        tw, th = int(rng.integers(4, 72)), int(rng.integers(4, 72))
        gw, gh = int(rng.integers(1, 9)), int(rng.integers(1, 7))
        strength = float(rng.choice([0.0, 0.3, 0.7, 1.0]))
        n = 24
        mats = synth.materials(n, w=mw, h=mh, seed=100 + trial)
        tgt = synth.target(gw * tw + int(rng.integers(0, 3)), gh * th + int(rng.integers(0, 3)), seed=200 + trial)
        gw_, gh_ = gw, gh
        if tgt.shape[1] // gw_ != tw or tgt.shape[0] // gh_ != th:
            continue
        placed = rng.integers(-1, n, gw * gh).astype(np.int32)
        ctx.set_materials(mats)
        ctx.set_target(tgt, gw, gh)
        ctx.cell_lab(fetch=False)
        ctx.set_placement(placed)
        out = ctx.render(strength)
        ref = orc.render(tgt, gw, gh, mats, placed, strength, threads=8)
        diff = np.abs(out.astype(np.int16) - ref.astype(np.int16))
        assert diff.max() <= 1, f"trial {trial}: {mw}x{mh} -> {tw}x{th}, strength {strength}: max diff {diff.max()}"


def test_render_identity_resize(ctx):
    # material already tile-sized: both resampling passes are skipped (fast_image_resize skips equal axes)
    mats = synth.materials(8, w=20, h=18)
    tgt = synth.target(200, 180)
    placed = (np.arange(100) % 8).astype(np.int32)
    ctx.set_materials(mats)
    ctx.set_target(tgt, 10, 10)
    ctx.cell_lab(fetch=False)
    ctx.set_placement(placed)
    out = ctx.render(0.0)
    ref = orc.render(tgt, 10, 10, mats, placed, 0.0)
    assert (out == ref).all()


# ---------------------------------------------------------------- whole path
def test_generate_matches_stagewise(ctx):
    cfg = synth.CONFIGS["C2"]
    mats = synth.materials(cfg["n_materials"])
    tgt = synth.target(cfg["target_w"], cfg["target_h"])
    gw, gh = cfg["grid_w"], cfg["grid_h"]
    ctx.set_materials(mats)
    ctx.set_target(tgt, gw, gh)
    placed = np.zeros(gw * gh, np.int32)
    img = np.zeros(ctx.output_shape(), np.uint8)
    rep = ctx.generate(cfg["max_usage"], cfg["adjacency_weight"], cfg["color_strength"], True, cfg["sa_iterations"],
                       cfg["seed"], build_similarity=True, placed_out=placed, image_out=img)
    # oracle, stage by stage, fed with the GPU Lab means
    mlab, clab = ctx.material_lab(), ctx.cell_lab()
    ridx, rd2 = orc.topk(mlab, clab, 100, threads=8)
    rplaced, rnfb, _, _ = orc.place(ridx, rd2, mlab, gw, gh, cfg["max_usage"], cfg["adjacency_weight"])
    assert rep.n_fallbacks == rnfb and rnfb > 0  # C2: capacity 4500 < 4800 cells
    rgrid, rres = orc.anneal(mlab, rplaced.reshape(gh, gw), cfg["sa_iterations"], cfg["seed"])
    assert rep.annealed == 1 and rep.opt.accepted_count == rres.accepted_count
    assert (placed == rgrid.ravel()).all()
    ref = orc.render(tgt, gw, gh, mats, placed, cfg["color_strength"], threads=8)
    assert np.abs(img.astype(np.int16) - ref.astype(np.int16)).max() <= 1


@pytest.mark.parametrize("max_usage,weight,iters,pinned", [(3, 0.3, 600, True), (1, 1.0, 0, True), (300, 0.3, 200, False),
                                                           (3, 0.3, 0, False)])
def test_generate_with_banded_upload_matches_stagewise(ctx, max_usage, weight, iters, pinned):
    """A host target of >= 8 MB goes up in bands; mosaic_generate then matches band by band and runs the placement
    chain on its own stream while the later bands are still being uploaded and matched (api.cu,
    run_match_place_banded).  Same results as the stage-by-stage oracle replay, and as the unbanded stage calls."""
    n, gw, gh = 1500, 96, 60                       # tile 24 x 24; capacity 4500 < 5760 cells at cap 3: resets fire
    mats = synth.materials(n)
    tgt = synth.target(2304 + 5, 1440 + 3, seed=77)  # 10 MB; remainder pixels right / below the grid
    assert tgt.nbytes >= 8 << 20
    if pinned:  # page-locked host buffers: the copies really are asynchronous and the chain really waits for its records
        import torch
        keep = [torch.empty(mats.shape, dtype=torch.uint8, pin_memory=True), torch.empty(tgt.shape, dtype=torch.uint8, pin_memory=True)]
        keep[0].numpy()[...] = mats
        keep[1].numpy()[...] = tgt
        mats, tgt = keep[0].numpy(), keep[1].numpy()
    ctx.set_materials(mats)
    ctx.set_target(tgt, gw, gh)
    placed = np.zeros(gw * gh, np.int32)
    img = np.zeros(ctx.output_shape(), np.uint8)
    rep = ctx.generate(max_usage, weight, 0.3, iters > 0, iters, 5, build_similarity=False, placed_out=placed, image_out=img)
    mlab, clab = ctx.material_lab(), ctx.cell_lab()
    assert_lab_close(clab, orc.cell_lab(tgt, gw, gh, threads=8))
    idx, d2 = ctx.candidates()                      # the lists the banded run left behind
    ridx, rd2 = orc.topk(mlab, clab, 100, threads=8)
    assert (idx == ridx).all() and (d2.view(np.uint32) == rd2.view(np.uint32)).all()
    rplaced, rnfb, rfb, _ = orc.place(ridx, rd2, mlab, gw, gh, max_usage, weight)
    assert rep.n_fallbacks == rnfb
    assert (ctx.fallback_cells() == rfb).all()
    if iters > 0 and weight > 0:
        rgrid, rres = orc.anneal(mlab, rplaced.reshape(gh, gw), iters, 5)
        assert rep.annealed == 1 and rep.opt.accepted_count == rres.accepted_count
        assert (placed == rgrid.ravel()).all()
    else:
        assert (placed == rplaced).all()
    ref = orc.render(tgt, gw, gh, mats, placed, 0.3, threads=8)
    assert np.abs(img.astype(np.int16) - ref.astype(np.int16)).max() <= 1
    # the unbanded stage calls on the same inputs
    ctx.set_target(tgt, gw, gh)
    ctx.cell_lab(fetch=False)
    p2, nfb2 = ctx.place(max_usage, weight)
    assert nfb2 == rnfb and (p2 == rplaced).all()


def test_generate_with_banded_upload_narrow_grid(ctx):
    """A >= 8 MB host target on a grid the chain kernel does not cover (grid_w < 33): the bands still feed the cell Lab
    means and the top-K as they land, the generic placement kernel runs afterwards."""
    n, gw, gh = 400, 24, 40
    mats = synth.materials(n)
    tgt = synth.target(1920, 1480, seed=31)  # 8.5 MB, tile 80 x 37
    assert tgt.nbytes >= 8 << 20
    ctx.set_materials(mats)
    ctx.set_target(tgt, gw, gh)
    placed = np.zeros(gw * gh, np.int32)
    rep = ctx.generate(2, 0.4, 0.0, False, 0, 0, build_similarity=False, placed_out=placed)
    mlab, clab = ctx.material_lab(), ctx.cell_lab()
    ridx, rd2 = orc.topk(mlab, clab, 100, threads=8)
    rplaced, rnfb, _, _ = orc.place(ridx, rd2, mlab, gw, gh, 2, 0.4)
    assert rep.n_fallbacks == rnfb and (placed == rplaced).all()


def test_fast_ieee_helpers_exhaustive(ctx):
    # sqrt/rcp sequences of the serial chains == correctly rounded builtins on all 2^32 bit patterns
    assert ctx.selftest() == (0, 0)


def test_errors_are_loud(ctx):
    import mosaic_b200
    with pytest.raises(mosaic_b200.MosaicError):
        ctx.material_lab()          # nothing set
    with pytest.raises(mosaic_b200.MosaicError):
        ctx.set_target(np.zeros((4, 4, 3), np.uint8), 8, 8)   # grid finer than the image
