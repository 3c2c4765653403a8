"""The wave placement chain (greedy_wave.cu; selected with mosaic_debug_set_chain(ctx, 3)) against the oracle's replay
of main.rs:335-475 — the same cases the serial chains are held to: bit-exact ids, fallback count, fallback cells.

A window of 32 consecutive cells is evaluated at once, one warp per cell, from the committed state and the previous
sweep's tentative choices; the prefix that did not change is final and is committed (greedy_wave.cu has the argument)."""
import numpy as np
import pytest

from mosaic_b200 import synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def smooth_cells(gw, gh, phase=0.0):
    yy, xx = np.mgrid[0:gh, 0:gw]
    lab = np.stack([50 + 30 * np.sin(xx / 5.0 + phase), 40 * np.cos(yy / 4.0), 30 * np.sin((xx + yy) / 6.0)], -1)
    return lab.reshape(-1, 3).astype(np.float32)


def check(ctx, mat_lab, cell_lab, gw, gh, max_usage, weight, tag=""):
    ctx.set_material_lab(mat_lab)
    ctx.set_cell_lab(cell_lab, gw, gh)
    idx, d2 = ctx.candidates()
    placed, nfb = ctx.place(max_usage, weight)
    rplaced, rnfb, rfb, _ = orc.place(idx, d2, mat_lab, gw, gh, max_usage, weight)
    assert nfb == rnfb, tag
    assert (placed == rplaced).all(), tag
    assert (ctx.fallback_cells() == rfb[:4096]).all(), tag


# grids of 32 columns and more take the wave chain (the cell above a window cell is always committed); narrower grids,
# a cap of 0 and caps beyond 16 bits fall back to the serial kernels
@pytest.mark.parametrize("n,gw,gh,max_usage,weight", [
    (500, 50, 28, 3, 0.3), (1500, 80, 60, 3, 0.25), (300, 33, 9, 3, 0.3), (300, 63, 9, 1, 1.0), (300, 64, 9, 2, 0.3),
    (300, 65, 20, 3, 0.0), (200, 90, 11, 0, 0.3), (120, 70, 9, 300, 0.3), (40, 100, 7, 70000, 0.3), (1, 40, 3, 3, 0.3),
    (2000, 200, 30, 3, 0.3), (150, 37, 40, 1, 2.5), (300, 20, 10, 3, 0.3), (400, 32, 12, 2, 0.3), (5000, 128, 40, 3, 0.3),
    (60, 45, 30, 254, 0.3), (60, 45, 30, 255, 0.3)])
def test_wave_chain_bit_exact(ctx, n, gw, gh, max_usage, weight):
    ctx.set_chain(3)
    check(ctx, synth.lab_points(n, seed=n), smooth_cells(gw, gh), gw, gh, max_usage, weight)


def test_wave_chain_randomized(ctx):
    ctx.set_chain(3)
    rng = np.random.default_rng(20261018)
    for trial in range(60):
        n = int(rng.choice([1, 2, 7, 31, 33, 64, 100, 150, 400, 1200]))
        gw = int(rng.integers(32, 120))
        gh = int(rng.integers(1, 14))
        max_usage = int(rng.choice([0, 1, 1, 2, 3, 3, 5, 40, 300]))
        weight = float(rng.choice([0.0, 0.05, 0.3, 0.3, 1.0, 2.5]))
        mat_lab = synth.lab_points(n, seed=3000 + trial)
        if trial % 4 == 0:  # clustered materials: many near-ties and threatened shortlists
            mat_lab = (mat_lab[rng.integers(0, max(1, n // 8), n)] + rng.normal(0, 0.05, (n, 3))).astype(np.float32)
        cells = smooth_cells(gw, gh, trial) + rng.normal(0, float(rng.choice([0.0, 0.5, 20.0])), (gw * gh, 3))
        check(ctx, mat_lab, cells.astype(np.float32), gw, gh, max_usage, weight,
              f"trial {trial}: n={n} grid={gw}x{gh} cap={max_usage} w={weight}")


def test_wave_chain_whole_pipeline_with_banded_upload(ctx):
    """mosaic_generate with a host target (bands, gated chain on its own stream) through the wave chain."""
    ctx.set_chain(3)
    n, gw, gh = 1500, 96, 60
    mats = synth.materials(n)
    tgt = synth.target(2304 + 5, 1440 + 3, seed=77)
    ctx.set_materials(mats)
    ctx.set_target(tgt, gw, gh)
    placed = np.zeros(gw * gh, np.int32)
    rep = ctx.generate(3, 0.3, 0.3, False, 0, 1, placed_out=placed)
    mlab, clab = ctx.material_lab(), ctx.cell_lab()
    ridx, rd2 = orc.topk(mlab, clab, 100, threads=8)
    rplaced, rnfb, _, _ = orc.place(ridx, rd2, mlab, gw, gh, 3, 0.3)
    assert rep.n_fallbacks == rnfb and rnfb > 0
    assert (placed == rplaced).all()
