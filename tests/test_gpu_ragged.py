"""Materials of individual sizes ("arbitrary material sizes -> variable-size K1/K4 paths"): the
reference opens files of any size, takes the Lab mean of each as it is (main.rs:263-301) and resizes each to the tile
(main.rs:710-738).  Same stage-wise protocol as test_gpu_parity.py, through mosaic_set_materials_ragged_rgb8."""
import numpy as np
import pytest

from oracle import oracle as orc
from mosaic_b200 import synth

pytestmark = pytest.mark.gpu

# (w, h): upscaled and downscaled on both axes, mixed, one axis equal to the tile's (skipped by the resampler),
# both equal (copied), tiny, and sources large enough that the intermediate leaves shared memory (700 rows x 60 x 3)
SIZES = [(64, 64), (20, 12), (20, 50), (100, 12), (7, 5), (1, 1), (33, 17), (128, 96), (20, 40), (90, 12), (300, 200),
         (48, 700), (3, 900)]


def ragged_materials(n, sizes=SIZES, seed=0):
    out = []
    for i in range(n):
        w, h = sizes[i % len(sizes)]
        out.append(synth.materials(1, w=w, h=h, start=i)[0])  # material i of the (w, h) generator: distinct colours
    return out


def lab_close(gpu, ref):
    tol = np.maximum(1e-4 * np.abs(ref), 1e-4)
    err = np.abs(gpu - ref)
    assert (err <= tol).all(), f"max err {err.max()} at {np.unravel_index(err.argmax(), err.shape)}"


def test_ragged_material_lab(ctx):
    mats = ragged_materials(40)
    ctx.set_materials_ragged(mats)
    ref = np.concatenate([orc.material_lab(m[None]) for m in mats])
    lab_close(ctx.material_lab(), ref)


def test_ragged_material_lab_large_image(ctx):
    big = synth.target(1600, 1200, seed=5)  # 1.9 M pixels through one warp: the two-level sums keep a/b
    ctx.set_materials_ragged([big, big[:37, :91].copy()])
    ref = np.concatenate([orc.material_lab(big[None], threads=8), orc.material_lab(big[None, :37, :91].copy())])
    lab_close(ctx.material_lab(), ref)


def oracle_render_ragged(tgt, gw, gh, mats, placed, strength):
    """orc.render takes uniform materials: one call per distinct size with the other cells empty, merged by cell."""
    th, tw = tgt.shape[0] // gh, tgt.shape[1] // gw
    out = np.zeros((gh * th, gw * tw, 3), np.uint8)
    placed = np.asarray(placed).reshape(gh, gw)
    for shape in sorted({m.shape for m in mats}):
        ids = [i for i, m in enumerate(mats) if m.shape == shape]
        remap = {m: k for k, m in enumerate(ids)}
        sub = np.vectorize(lambda v: remap.get(v, -1))(placed).astype(np.int32)
        if (sub < 0).all():
            continue
        part = orc.render(tgt, gw, gh, np.stack([mats[i] for i in ids]), sub.ravel(), strength, threads=8)
        for cy, cx in zip(*np.nonzero(sub >= 0)):
            out[cy * th:(cy + 1) * th, cx * tw:(cx + 1) * tw] = part[cy * th:(cy + 1) * th, cx * tw:(cx + 1) * tw]
    return out


@pytest.mark.parametrize("w,h,gw,gh,strength", [(480, 288, 24, 24, 0.3), (480, 288, 24, 24, 0.0), (250, 130, 7, 3, 1.0),
                                                 (400, 360, 20, 30, 0.5)])
def test_ragged_render_within_one_lsb(ctx, w, h, gw, gh, strength):
    # tiles: 20x12 (three of the sizes above share one or both axes with it), 35x43, 20x12
    mats = ragged_materials(39)
    tgt = synth.target(w, h, seed=3)
    rng = np.random.default_rng(gw * gh)
    placed = rng.integers(-1, len(mats), gw * gh).astype(np.int32)
    ctx.set_materials_ragged(mats)
    ctx.set_target(tgt, gw, gh)
    ctx.cell_lab(fetch=False)
    ctx.set_placement(placed)
    out = ctx.render(strength)
    ref = oracle_render_ragged(tgt, gw, gh, mats, placed, strength)
    diff = np.abs(out.astype(np.int16) - ref.astype(np.int16))
    assert diff.max() <= 1, f"max diff {diff.max()} at {np.unravel_index(diff.argmax(), diff.shape)}"


def test_ragged_generate_matches_stagewise(ctx):
    """The whole path on ragged materials: placements from the GPU's Lab means replayed by the oracle, annealing,
    render — and the same mosaic as the uniform entry point when all materials happen to have one size."""
    mats = ragged_materials(260)
    tgt = synth.target(640, 384)
    gw, gh = 32, 24
    ctx.set_materials_ragged(mats)
    ctx.set_target(tgt, gw, gh)
    placed = np.zeros(gw * gh, np.int32)
    img = np.zeros(ctx.output_shape(), np.uint8)
    rep = ctx.generate(max_usage=3, adjacency_weight=0.3, color_strength=0.3, enable_optimization=True,
                       optimization_iterations=700, seed=9, build_similarity=True, placed_out=placed, image_out=img)
    mlab, clab = ctx.material_lab(), ctx.cell_lab()
    lab_close(mlab, np.concatenate([orc.material_lab(m[None]) for m in mats]))
    idx, d2 = orc.topk(mlab, clab, 100)
    rplaced, rnfb, _, _ = orc.place(idx, d2, mlab, gw, gh, 3, 0.3)
    assert rep.n_fallbacks == rnfb
    rgrid, rres = orc.anneal(mlab, rplaced.reshape(gh, gw), 700, 9)
    assert (placed.reshape(gh, gw) == rgrid).all() and rep.opt.accepted_count == rres.accepted_count
    ref = oracle_render_ragged(tgt, gw, gh, mats, placed, 0.3)
    assert np.abs(img.astype(np.int16) - ref.astype(np.int16)).max() <= 1

    uni = synth.materials(120)
    ctx.set_materials(uni)
    ctx.set_target(tgt, gw, gh)
    p1, i1 = np.zeros(gw * gh, np.int32), np.zeros(ctx.output_shape(), np.uint8)
    ctx.generate(max_usage=8, adjacency_weight=0.3, color_strength=0.3, enable_optimization=True,
                 optimization_iterations=300, seed=1, build_similarity=False, placed_out=p1, image_out=i1)
    ctx.set_materials_ragged(list(uni))
    ctx.set_target(tgt, gw, gh)
    p2, i2 = np.zeros(gw * gh, np.int32), np.zeros(ctx.output_shape(), np.uint8)
    ctx.generate(max_usage=8, adjacency_weight=0.3, color_strength=0.3, enable_optimization=True,
                 optimization_iterations=300, seed=1, build_similarity=False, placed_out=p2, image_out=i2)
    # (the Lab means of the two entry points differ in the last bits — one sums 16 pixels per lane step, the other
    # folds 64-pixel partial sums — so identical placements are expected, not guaranteed; the images must agree
    # wherever the placements do)
    same = p1 == p2
    assert same.mean() > 0.99
    th, tw = i1.shape[0] // gh, i1.shape[1] // gw
    for c in np.nonzero(same)[0][:200]:
        cy, cx = divmod(int(c), gw)
        assert (i1[cy * th:(cy + 1) * th, cx * tw:(cx + 1) * tw] == i2[cy * th:(cy + 1) * th, cx * tw:(cx + 1) * tw]).all()


def test_ragged_generate_with_banded_upload(ctx):
    """Materials of individual sizes together with a >= 8 MB host target: the upload stream carries the ragged block
    and the size tables first, then the target in bands that are matched and placed as they land (api.cu)."""
    mats = ragged_materials(700)
    tgt = synth.target(2304, 1440, seed=13)
    gw, gh = 96, 60
    assert tgt.nbytes >= 8 << 20
    ctx.set_materials_ragged(mats)
    ctx.set_target(tgt, gw, gh)
    placed = np.zeros(gw * gh, np.int32)
    img = np.zeros(ctx.output_shape(), np.uint8)
    rep = ctx.generate(max_usage=9, adjacency_weight=0.3, color_strength=0.3, enable_optimization=True,
                       optimization_iterations=400, seed=2, build_similarity=False, placed_out=placed, image_out=img)
    mlab, clab = ctx.material_lab(), ctx.cell_lab()
    idx, d2 = orc.topk(mlab, clab, 100, threads=8)
    rplaced, rnfb, _, _ = orc.place(idx, d2, mlab, gw, gh, 9, 0.3)
    assert rep.n_fallbacks == rnfb
    rgrid, rres = orc.anneal(mlab, rplaced.reshape(gh, gw), 400, 2)
    assert (placed.reshape(gh, gw) == rgrid).all() and rep.opt.accepted_count == rres.accepted_count
    ref = oracle_render_ragged(tgt, gw, gh, mats, placed, 0.3)
    assert np.abs(img.astype(np.int16) - ref.astype(np.int16)).max() <= 1


def test_ragged_errors_are_loud(ctx):
    import mosaic_b200
    with pytest.raises(mosaic_b200.MosaicError):
        ctx.set_materials_ragged([np.zeros((0, 4, 3), np.uint8)])
