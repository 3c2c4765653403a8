"""The progress surface: mosaic_set_progress_callback feeds what the GUI's `(f32, String)` channel
(src/gui/app_full.rs:1597-1616, fractions as in :1716-2193) and the CLI's progress lines consume, from stage events of
the device — host functions behind the stages of mosaic_generate."""
import numpy as np
import pytest

from mosaic_b200 import synth

pytestmark = pytest.mark.gpu


def run(ctx, with_host_image, events):
    n, gw, gh = 1500, 96, 60
    ctx.set_materials(synth.materials(n))
    ctx.set_target(synth.target(2304, 1440, seed=5), gw, gh)
    placed = np.zeros(gw * gh, np.int32)
    img = np.zeros(ctx.output_shape(), np.uint8) if with_host_image else None
    ctx.generate(3, 0.3, 0.3, True, 500, 9, build_similarity=True, placed_out=placed, image_out=img)
    return placed, img, list(events)


@pytest.mark.parametrize("with_host_image", [False, True])
def test_progress_events_follow_the_device(ctx, with_host_image):
    events = []
    ctx.set_progress_callback(lambda frac, stage, msg: events.append((frac, stage, msg)))
    placed, img, ev = run(ctx, with_host_image, events)
    msgs = [m for _, _, m in ev]
    assert "Loaded 1500 tiles" in msgs and "Building similarity database..." in msgs
    assert "Processing grid cells..." in msgs and "Grid processing completed" in msgs
    assert "Starting optimization phase..." in msgs and "Optimization phase completed" in msgs and msgs[-1] == "Mosaic image ready"
    # the placement stage reports the rows the chain has committed (app_full.rs:1810-1827), increasing up to 100 %
    rows = [float(m.split(": ")[1].rstrip("%")) for m in msgs if m.startswith("Processing grid: ")]
    assert len(rows) >= 2 and rows == sorted(rows) and rows[-1] == 100.0
    # events of different streams may overtake each other by a little (the matrix and the rendered rows travel beside the
    # placement path): the reported fraction never goes backwards
    fr = [f for f, _, _ in ev]
    assert fr == sorted(fr) and fr[0] == pytest.approx(0.40) and fr[-1] == pytest.approx(0.99)
    # the callback changes nothing: same result without it, and no events afterwards
    ctx.set_progress_callback(None)
    events.clear()
    placed2, img2, ev2 = run(ctx, with_host_image, events)
    assert ev2 == [] and (placed == placed2).all()
    if with_host_image:
        assert (img == img2).all()


def test_page_locked_buffers_of_the_library_and_of_the_caller(ctx):
    """mosaic_host_alloc / mosaic_host_register: the host-buffer path with page-locked memory (asynchronous uploads, rows
    downloaded under the chain, patched tiles written straight into the caller's image) gives the pageable path's result."""
    import mosaic_b200
    n, gw, gh = 1200, 96, 60
    mats, tgt = synth.materials(n), synth.target(2304, 1440, seed=11)
    kw = dict(max_usage=3, adjacency_weight=0.3, color_strength=0.3, enable_optimization=True, optimization_iterations=400, seed=4)

    def run(m, t, img):
        ctx.set_materials(m)
        ctx.set_target(t, gw, gh)
        placed = np.zeros(gw * gh, np.int32)
        ctx.generate(placed_out=placed, image_out=img, **kw)
        return placed

    ref_img = np.zeros((gh * (1440 // gh), gw * (2304 // gw), 3), np.uint8)
    ref_placed = run(mats, tgt, ref_img)                       # pageable everything
    pm, pt, pi = (mosaic_b200.PinnedArray(a.shape) for a in (mats, tgt, ref_img))
    pm.array[...] = mats
    pt.array[...] = tgt
    assert (run(pm.array, pt.array, pi.array) == ref_placed).all() and (pi.array == ref_img).all()
    own_img = np.zeros_like(ref_img)
    with mosaic_b200.register(mats.copy()) as rm, mosaic_b200.register(tgt.copy()) as rt, mosaic_b200.register(own_img) as ri:
        assert (run(rm, rt, ri) == ref_placed).all()
        ctx.sync()
        assert (ri == ref_img).all()
    for p in (pm, pt, pi):
        p.close()
