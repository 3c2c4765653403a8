"""CPU tests pinning the oracle against every numeric known-answer the reference's own tests hold for
the hot path — citations are file:line under /root/reference/src/ — plus
known-answer tests of the restated third-party algorithms (ChaCha / rand / Lanczos / HSV)."""
import ctypes

import numpy as np
import pytest

from oracle import oracle as orc


def solid(c, w=10, h=10):
    img = np.zeros((h, w, 3), np.uint8)
    img[:] = c
    return img


# ---------------------------------------------------------------- lib.rs
def test_lab_red_golden():  # lib.rs:133-144
    lab = orc.average_lab(solid((255, 0, 0)))
    assert abs(lab[0] - 53.24) < 1.0 and abs(lab[1] - 80.09) < 1.0 and abs(lab[2] - 67.20) < 1.0


def test_lab_gray_golden():  # lib.rs:146-158
    lab = orc.average_lab(solid((128, 128, 128)))
    assert abs(lab[0] - 53.59) < 1.0 and abs(lab[1]) < 1.0 and abs(lab[2]) < 1.0


def test_lab_cli_tolerance():  # main.rs:822-825 (+-5.0 on red)
    lab = orc.average_lab(solid((255, 0, 0), 4, 4))
    assert abs(lab[0] - 53.24) < 5.0


@pytest.mark.parametrize("rgb,exp", [
    ((255, 0, 0), (53.2408, 80.09243, 67.20321)), ((0, 255, 0), (87.73472, -86.182686, 83.17931)),
    ((0, 0, 255), (32.29701, 79.187546, -107.86017)), ((128, 128, 128), (53.585022, 0, 0)),
    ((255, 255, 255), (100, 0, 0)), ((0, 0, 0), (0, 0, 0))])
def test_lab_single_pixel_values(rgb, exp):  # SURVEY App. B (palette formula, tighter than the +-1.0 goldens)
    lab = orc.rgb8_to_lab(*rgb)
    assert np.allclose(lab, exp, rtol=1e-4, atol=1e-4)


def test_lab_kdtree_fixture_points():  # lib.rs:212-227: nearest of 4 literal Lab points
    pts = np.array([[50, 0, 0], [60, 10, 10], [40, -10, -10], [55, 5, 5]], np.float32)
    idx, d2 = orc.topk(pts, np.array([[51, 1, 1]], np.float32), 1)
    assert idx[0, 0] == 0 and d2[0, 0] == np.float32(3.0)


def test_average_rgb_exact():  # main.rs:866-889
    assert tuple(orc.average_rgb(solid((255, 0, 0)))) == (255, 0, 0)
    img = np.array([[[255, 0, 0], [0, 255, 0]], [[0, 0, 255], [255, 255, 255]]], np.uint8)
    assert tuple(orc.average_rgb(img)) == (127, 127, 127)


def test_lab_mean_is_sequential_f32_sum():  # lib.rs:109-111 fold order
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (13, 11, 3), dtype=np.uint8)
    acc = np.zeros(3, np.float32)
    for y in range(13):
        for x in range(11):
            acc = (acc + orc.rgb8_to_lab(*img[y, x])).astype(np.float32)
    exp = (acc / np.float32(13 * 11)).astype(np.float32)
    assert (orc.average_lab(img).view(np.uint32) == exp.view(np.uint32)).all()


def test_cell_lab_ignores_remainder_pixels():  # main.rs:495-496, 536-540
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (23, 31, 3), dtype=np.uint8)
    lab = orc.cell_lab(img, 4, 3)  # tile 7x7; columns 28..30 and rows 21..22 are never read
    img2 = img.copy()
    img2[:, 28:] = 0
    img2[21:] = 0
    assert (lab == orc.cell_lab(img2, 4, 3)).all()
    assert (lab[5] == orc.average_lab(img[7:14, 7:14])).all()


# ---------------------------------------------------------------- similarity.rs / adjacency.rs
def test_packed_index():  # similarity.rs:100, App. B
    assert [orc.tri_index(0, 1, 3), orc.tri_index(0, 2, 3), orc.tri_index(1, 2, 3)] == [0, 1, 2]
    assert [orc.tri_index(*p, 4) for p in [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]] == list(range(6))


def test_similarity_fixture():  # similarity.rs:190-192
    lab = np.array([[50, 0, 0], [60, 10, 10], [40, -10, -10]], np.float32)
    tri = orc.similarity_tri(lab)
    assert tri[0] == np.float32(17.320509) and tri[1] == np.float32(17.320509) and tri[2] == np.float32(34.641018)
    assert orc.lab_distance(lab[0], lab[1]) == orc.lab_distance(lab[1], lab[0])  # symmetry (similarity.rs:185-250)
    assert orc.lab_distance(lab[0], lab[0]) == 0.0


def test_similarity_threads_equal_serial():
    lab = np.random.default_rng(3).normal(0, 30, (211, 3)).astype(np.float32)
    assert (orc.similarity_tri(lab, 1) == orc.similarity_tri(lab, 5)).all()


def test_adjacency_penalty_fixture():  # adjacency.rs:243-271 -> 1/(17.320509+1)
    lab = np.array([[50, 0, 0], [60, 10, 10], [40, -10, -10]], np.float32)
    grid = np.array([[0, 1]], np.int32)
    p = orc.penalty(lab, grid, 1, 0, 2, 1.0)
    assert p > 0 and np.float32(p) == np.float32(0.054583635)


def test_total_cost_identical_colours():  # adjacency.rs:273-290 -> exactly 1.0
    lab = np.array([[50, 0, 0], [50, 0, 0]], np.float32)
    assert orc.total_cost(lab, np.array([[0, 1]], np.int32)) == np.float32(1.0)


def test_total_cost_optimizer_fixture():  # optimizer.rs:230-245 -> 0.24974766 (0x3E7FBDDA)
    lab = np.array([[50, 0, 0], [60, 10, 10], [40, -10, -10], [55, 5, 5]], np.float32)
    tc = orc.total_cost(lab, np.array([[0, 1], [2, 3]], np.int32))
    assert np.float32(tc).view(np.uint32) == 0x3E7FBDDA


def test_swap_delta_consistent_with_total_cost():
    rng = np.random.default_rng(2)
    lab = rng.normal(0, 30, (40, 3)).astype(np.float32)
    grid = rng.integers(0, 40, (6, 7)).astype(np.int32)
    for _ in range(200):
        x1, y1, x2, y2 = rng.integers(0, 7), rng.integers(0, 6), rng.integers(0, 7), rng.integers(0, 6)
        d = orc.swap_delta(lab, grid, (x1, y1), (x2, y2))
        g2 = grid.copy()
        g2[y1, x1], g2[y2, x2] = grid[y2, x2], grid[y1, x1]
        assert abs(float(orc.total_cost(lab, g2)) - float(orc.total_cost(lab, grid)) - float(d)) < 1e-4


# ---------------------------------------------------------------- greedy (main.rs:309-475)
def _lists(lab, cells_lab):
    return orc.topk(lab, cells_lab, 100)


def test_usage_cap_and_fallback_reset():  # main.rs:1345-1378 semantics
    lab = np.array([[10, 0, 0], [90, 0, 0]], np.float32)
    cells = np.tile(np.array([[12, 0, 0]], np.float32), (4, 1))  # everybody wants material 0
    idx, d2 = _lists(lab, cells)
    placed, nfb, fb, usage = orc.place(idx, d2, lab, 4, 1, 1, 0.0)
    # cell0 -> 0; cell1: 0 is used up and adjacent -> 1; cell2: both at cap -> fallback, reset, first placeable = 0
    assert placed.tolist() == [0, 1, 0, 1] and nfb == 1 and fb.tolist() == [2]
    assert usage.tolist() == [1, 1]


def test_same_image_never_adjacent_when_avoidable():  # main.rs:923-1024
    lab = np.array([[10, 0, 0], [11, 0, 0], [12, 0, 0]], np.float32)
    cells = np.tile(np.array([[10, 0, 0]], np.float32), (9, 1))
    idx, d2 = _lists(lab, cells)
    placed, _, _, _ = orc.place(idx, d2, lab, 3, 3, 100, 0.0)
    g = placed.reshape(3, 3)
    assert (g[:, 1:] != g[:, :-1]).all() and (g[1:] != g[:-1]).all()


def test_single_material_final_fallback():  # main.rs:456-474
    lab = np.array([[10, 0, 0]], np.float32)
    cells = np.zeros((4, 3), np.float32)
    idx, d2 = _lists(lab, cells)
    placed, nfb, _, _ = orc.place(idx, d2, lab, 2, 2, 3, 0.3)
    assert placed.tolist() == [0, 0, 0, 0] and nfb == 3


def test_max_usage_zero():  # lib.rs:288-306: max 0 -> can_use_image always false -> fallback every cell
    lab = np.array([[10, 0, 0], [20, 0, 0], [30, 0, 0]], np.float32)
    idx, d2 = _lists(lab, np.zeros((6, 3), np.float32))
    placed, nfb, _, _ = orc.place(idx, d2, lab, 3, 2, 0, 0.3)
    assert nfb == 6 and (placed >= 0).all()


def test_strict_less_keeps_first_of_equal_scores():  # main.rs:401-407
    lab = np.array([[10, 0, 0], [10, 0, 0], [50, 0, 0]], np.float32)  # 0 and 1 tie exactly
    idx, d2 = _lists(lab, np.array([[10, 0, 0]], np.float32))
    assert idx[0].tolist() == [0, 1, 2]
    placed, _, _, _ = orc.place(idx, d2, lab, 1, 1, 3, 0.0)
    assert placed[0] == 0


# ---------------------------------------------------------------- RNG [3P]: rand 0.8.5 / rand_chacha 0.3.1
def test_chacha_block_known_answers():
    z = np.zeros(8, np.uint32)
    assert orc.chacha_block(z, 0, 20)[:4].tobytes().hex() == "76b8e0ada0f13d90405d6ae55386bd28"  # RFC 7539 zero key
    assert orc.chacha_block(z, 0, 12)[:4].tobytes().hex().startswith("9bf49a6a0755f953811fce125f2683d5")
    assert orc.chacha_block(z, 0, 8)[:2].tobytes().hex() == "3e00ef2f895f40d6"


def test_stdrng_seed_stream():  # SURVEY §8(c) known answers
    r = orc.Rng(0)
    assert [r.next_u32() for _ in range(4)] == [3442241407, 3140108210, 2384947579, 3321986196]
    assert orc.Rng(0).next_u64() == 13486662071293341567
    r = orc.Rng(42)
    assert [r.next_u32() for _ in range(4)] == [572990626, 2261546851, 1068323197, 2330987027]


def test_blockrng_u64_is_two_consecutive_words_even_across_refill():
    a, b = orc.Rng(7), orc.Rng(7)
    words = [a.next_u32() for _ in range(200)]
    for _ in range(63):
        b.next_u32()
    v = b.next_u64()  # index 63: low = word 63, high = word 64 (first of the next 64-word buffer)
    assert v == words[63] | (words[64] << 32)
    assert b.next_u32() == words[65]


def test_gen_range_and_f32():
    r = orc.Rng(1)
    xs = [r.gen_range(7) for _ in range(2000)]
    assert min(xs) == 0 and max(xs) == 6
    r2, r3 = orc.Rng(3), orc.Rng(3)
    for n in (1 << 40, 200, 120, 1000003, (1 << 63) + 5):  # rand 0.8.5 sample_single: widening multiply + zone
        zone = ((n << (64 - n.bit_length())) - 1) & (2 ** 64 - 1)
        while True:
            m = r2.next_u64() * n
            if (m & (2 ** 64 - 1)) <= zone:
                break
        assert r3.gen_range(n) == m >> 64
    f = [float(r.gen_f32()) for _ in range(2000)]
    assert 0.0 <= min(f) and max(f) < 1.0
    r4, r5 = orc.Rng(9), orc.Rng(9)
    assert r4.gen_f32() == np.float32((r5.next_u32() >> 8) * 2.0 ** -24)


def test_expf_correctly_rounded():
    libm = ctypes.CDLL("libm.so.6")
    libm.expf.restype = ctypes.c_float
    libm.expf.argtypes = [ctypes.c_float]
    rng = np.random.default_rng(0)
    xs = np.concatenate([-rng.random(20000, dtype=np.float32) * 20, -rng.random(5000, dtype=np.float32) * 1e-3,
                         rng.standard_normal(5000).astype(np.float32) * 30, np.float32([0.0, -0.0, -103.9, 88.7])])
    import math
    wrong = glibc_diff = 0
    for x in xs:
        got = np.float32(orc.expf(x))
        with np.errstate(over="ignore"):
            exp = np.float32(math.exp(float(x))) if x < 700 else np.float32(np.inf)
        wrong += int(got != exp)
        glibc_diff += int(got != np.float32(libm.expf(float(x))))
    assert wrong == 0                      # = round-to-nearest of the double-precision exp
    assert glibc_diff < len(xs) * 2e-3     # the platform libm Rust would call differs in <0.2% by 1 ulp


# ---------------------------------------------------------------- optimizer.rs invariants
LAB4 = np.array([[50, 0, 0], [60, 10, 10], [40, -10, -10], [55, 5, 5]], np.float32)


def test_optimizer_empty_grid_default():  # optimizer.rs:295-307
    g, res = orc.anneal(LAB4, np.zeros((0, 0), np.int32), 100, 0)
    assert res.iterations == 0 and res.initial_cost == 0 and res.accepted_count == 0


def test_optimizer_single_cell():  # optimizer.rs:309-331
    g, res = orc.anneal(LAB4, np.array([[1]], np.int32), 100, 0)
    assert res.accepted_count == 0 and res.iterations == 100 and g[0, 0] == 1


def test_optimizer_zero_iterations():  # optimizer.rs:439-457
    grid = np.array([[0, 1], [2, 3]], np.int32)
    g, res = orc.anneal(LAB4, grid, 0, 0)
    assert (g == grid).all() and res.initial_cost == res.final_cost == res.best_cost


def test_optimizer_invariants_and_determinism():  # optimizer.rs:247-264, 333-437
    rng = np.random.default_rng(0)
    lab = rng.normal(0, 30, (30, 3)).astype(np.float32)
    grid = rng.integers(0, 30, (8, 9)).astype(np.int32)
    g1, r1 = orc.anneal(lab, grid, 3000, 5)
    g2, r2 = orc.anneal(lab, grid, 3000, 5)
    assert (g1 == g2).all() and r1.accepted_count == r2.accepted_count
    assert sorted(g1.ravel().tolist()) == sorted(grid.ravel().tolist())  # swaps only permute
    assert np.isfinite(r1.final_cost) and r1.final_cost >= 0 and r1.best_cost <= r1.initial_cost
    assert abs(float(orc.total_cost(lab, g1)) - r1.final_cost) < 1e-2
    assert r1.iterations == 3000
    g3, r3 = orc.anneal(lab, grid, 3000, 6)
    assert not (g1 == g3).all()


def test_optimizer_skips_empty_cells():  # optimizer.rs:71-73
    grid = np.array([[0, -1], [-1, 3]], np.int32)
    g, res = orc.anneal(LAB4, grid, 500, 1)
    assert (g.ravel() < 0).sum() == 2 and g[0, 1] == -1 and g[1, 0] == -1


# ---------------------------------------------------------------- color_adjustment.rs
def test_color_adjustment_clamps():  # color_adjustment.rs:169-176, 444-452
    assert orc.color_adjustment_new(2.0, -1.0, 400.0, 3.0).tolist() == [1.0, 0.0, 180.0, 2.0]
    assert orc.color_adjustment_new(-2.0, 5.0, -500.0, 10.0).tolist() == [-1.0, 2.0, -180.0, 2.0]


def test_identity_adjustment_within_one():  # color_adjustment.rs:178-192
    adj = np.array([0, 1, 0, 1], np.float32)
    for px in [(128, 64, 192), (0, 0, 0), (255, 255, 255), (17, 200, 90)]:
        out = orc.adjust_pixel(adj, px)
        assert np.abs(out.astype(int) - np.array(px)).max() <= 1


def test_brightness_contrast_direction():  # color_adjustment.rs:194-260
    px = (100, 100, 100)
    assert orc.adjust_pixel(np.float32([0.2, 1, 0, 1]), px)[0] > 100 > orc.adjust_pixel(np.float32([-0.2, 1, 0, 1]), px)[0]
    hi, lo = orc.adjust_pixel(np.float32([0, 1.5, 0, 1]), (200, 200, 200)), orc.adjust_pixel(np.float32([0, 1.5, 0, 1]), (50, 50, 50))
    assert hi[0] > 200 and lo[0] < 50


def test_hue_and_saturation():  # color_adjustment.rs:262-330, 455-464
    red = (255, 0, 0)
    out = orc.adjust_pixel(np.float32([0, 1, 120, 1]), red)  # +120 degrees: red -> green
    assert out[1] > 200 and out[0] < 50
    out = orc.adjust_pixel(np.float32([0, 1, 0, 0]), (200, 100, 50))  # saturation 0 -> gray
    assert out[0] == out[1] == out[2]
    out = orc.adjust_pixel(np.float32([0, 1, -30, 1]), red)  # wraparound to 330 degrees
    assert out[0] > 200 and out[2] > 100 and out[1] < 20


def test_optimal_adjustment_zero_strength_and_half():  # color_adjustment.rs:397-433
    z = orc.optimal_adjustment((100, 50, 50), (200, 100, 100), 0.0)
    assert z.tolist() == [0.0, 1.0, 0.0, 1.0]
    full = orc.optimal_adjustment((100, 100, 100), (200, 200, 200), 1.0)
    half = orc.optimal_adjustment((100, 100, 100), (200, 200, 200), 0.5)
    assert abs(half[0] - full[0] * 0.5) < 0.01 and full[0] > 0


def test_hsv_roundtrip():
    rng = np.random.default_rng(4)
    for _ in range(500):
        r, g, b = rng.random(3).astype(np.float32)
        h, s, v = orc.rgb_to_hsv(r, g, b)
        back = orc.hsv_to_rgb(h, s, v)
        assert np.allclose(back, [r, g, b], atol=2e-6)
    assert orc.rgb_to_hsv(1.0, 0.0, 0.5)[0] < 0  # raw (unnormalised) hue on the magenta side


# ---------------------------------------------------------------- resize (main.rs:710-738) [3P fast_image_resize]
def test_resize_solid_and_dims():  # main.rs:891-921
    out = orc.resize_rgb8(solid((255, 0, 0), 64, 64), 38, 36)
    assert out.shape == (36, 38, 3) and (out == np.array([255, 0, 0])).all()


def test_resize_coefficients_structure():
    start, size, coeffs, prec = orc.resize_coeffs(64, 38)
    assert coeffs.shape[1] == 13 and prec == 15  # support 3*64/38 = 5.05 -> window 2*6+1
    sums = np.array([coeffs[o, :size[o]].sum() for o in range(38)])
    assert np.abs(sums - (1 << prec)).max() <= 8  # each row sums to ~1.0 in fixed point
    assert (start >= 0).all() and (start + size <= 64).all()


@pytest.mark.parametrize("dw,dh", [(38, 38), (38, 36), (30, 17), (24, 18), (64, 32)])
def test_resize_matches_pillow_family(dw, dh):
    # sanity check only: Pillow's LANCZOS is the same algorithm with 22-bit coefficients
    from PIL import Image
    src = np.random.default_rng(dw * dh).integers(0, 256, (64, 64, 3), dtype=np.uint8)
    a = orc.resize_rgb8(src, dw, dh)
    b = np.asarray(Image.fromarray(src).resize((dw, dh), Image.LANCZOS))
    d = np.abs(a.astype(int) - b.astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 0.05


def test_resize_same_size_is_copy():
    src = np.random.default_rng(0).integers(0, 256, (20, 30, 3), dtype=np.uint8)
    assert (orc.resize_rgb8(src, 30, 20) == src).all()


def test_render_composes_tiles():  # main.rs:1380-1424: 100x100 target, 4x4 grid -> 100x100 output
    rng = np.random.default_rng(0)
    tgt = rng.integers(0, 256, (100, 100, 3), dtype=np.uint8)
    mats = rng.integers(0, 256, (5, 16, 16, 3), dtype=np.uint8)
    placed = (np.arange(16) % 5).astype(np.int32)
    out = orc.render(tgt, 4, 4, mats, placed, 0.0)
    assert out.shape == (100, 100, 3)
    assert (out[25:50, 50:75] == orc.resize_rgb8(mats[placed[6]], 25, 25)).all()
    out2 = orc.render(tgt, 4, 4, mats, placed, 0.3, threads=4)
    assert out2.shape == out.shape and (out2 != out).any()


def test_kdtree_topk_equals_brute_force():
    """The CPU arm of bench.py searches a bucketed k-d tree (the cost structure of the reference's kiddo search): same
    (d2 bits, index) lists as the brute force the GPU parity tests use, including exact ties and duplicates."""
    from mosaic_b200 import synth
    rng = np.random.default_rng(3)
    for n, c, k in [(5000, 800, 100), (500, 200, 100), (37, 50, 100), (1, 4, 100), (3000, 300, 7)]:
        m, q = synth.lab_points(n, seed=n), synth.lab_points(c, seed=c + 1)
        a, b = orc.topk(m, q, k, threads=4), orc.topk_kdtree(m, q, k, threads=4)
        assert (a[0] == b[0]).all() and (a[1].view(np.uint32) == b[1].view(np.uint32)).all()
    m = (rng.integers(0, 6, (4000, 3)) * 10).astype(np.float32)   # quantised colours: ties everywhere
    q = (rng.integers(0, 6, (300, 3)) * 10).astype(np.float32)
    a, b = orc.topk(m, q, 100, threads=4), orc.topk_kdtree(m, q, 100, threads=4)
    assert (a[0] == b[0]).all() and (a[1].view(np.uint32) == b[1].view(np.uint32)).all()
    m, q = np.zeros((500, 3), np.float32), np.ones((5, 3), np.float32)  # every material the same colour: one big leaf
    assert (orc.topk(m, q, 100)[0] == orc.topk_kdtree(m, q, 100)[0]).all()


# ---------------------------------------------------------------- independent implementations found in this image
def test_chacha_block_against_an_independent_chacha20():
    """The block function with ARBITRARY keys and 64-bit counters against the ChaCha20 of the `cryptography` package
    (OpenSSL): its 16-byte nonce is state words 12..15, so `counter (u64 LE) || 8 zero bytes` is the djb layout rand_chacha
    uses (64-bit block counter, stream 0).  The 12-round variant the annealer draws from shares everything but the loop count
    (pinned by the zero-key vectors above)."""
    ciphers = pytest.importorskip("cryptography.hazmat.primitives.ciphers")
    rng = np.random.default_rng(12)
    for trial in range(40):
        key = rng.integers(0, 2 ** 32, 8, dtype=np.uint64).astype(np.uint32)
        counter = int(rng.integers(0, 2 ** 62)) if trial % 3 else int(rng.integers(0, 1000)) + (trial % 2) * (2 ** 32 - 500)
        nonce = counter.to_bytes(8, "little") + bytes(8)
        enc = ciphers.Cipher(ciphers.algorithms.ChaCha20(key.tobytes(), nonce), mode=None).encryptor()
        want = enc.update(bytes(64 * 3))                                    # three consecutive blocks of key stream
        got = b"".join(orc.chacha_block(key, counter + i, 20).tobytes() for i in range(3))
        if (counter & 0xFFFFFFFF) + 2 >= 2 ** 32:
            want = want[:64 * ((2 ** 32 - (counter & 0xFFFFFFFF)))]         # (OpenSSL's counter is 32-bit: compare up to its wrap)
            got = got[:len(want)]
        assert got == want, (trial, counter)


def test_hsv_against_colorsys():
    """The hexcone conversions against Python's colorsys (same model, hue in turns instead of degrees, different operation
    order): agreement to float rounding guards the restatement against a wrong sector or sign, not against the last ulp."""
    import colorsys
    rng = np.random.default_rng(8)
    samples = [tuple(rng.random(3).astype(np.float32)) for _ in range(2000)]
    samples += [(0.0, 0.0, 0.0), (1.0, 1.0, 1.0), (0.5, 0.5, 0.5), (1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0), (1.0, 1.0, 0.0),
                (0.0, 1.0, 1.0), (1.0, 0.0, 1.0), (0.2, 0.2, 0.7), (0.7, 0.2, 0.2000001)]
    for r, g, b in samples:
        h, s, v = orc.rgb_to_hsv(r, g, b)
        ch, cs, cv = colorsys.rgb_to_hsv(float(r), float(g), float(b))
        assert abs(float(s) - cs) < 1e-5 and abs(float(v) - cv) < 1e-6
        if cs > 1e-4:
            dh = (float(h) % 360.0 - ch * 360.0 + 180.0) % 360.0 - 180.0
            assert abs(dh) < 1e-2 * max(1.0, 1e-2 / cs), (r, g, b, h, ch * 360.0)
        back = orc.hsv_to_rgb(float(h), float(s), float(v))
        want = colorsys.hsv_to_rgb((float(h) % 360.0) / 360.0, float(s), float(v))
        assert np.allclose(back, want, atol=2e-6), (r, g, b)
        for shift in (-400.0, -123.0, 77.0, 360.0, 725.0):                  # hues outside [0, 360): into_positive_degrees
            assert np.allclose(orc.hsv_to_rgb(float(h) + shift, float(s), float(v)),
                               colorsys.hsv_to_rgb(((float(h) + shift) % 360.0) / 360.0, float(s), float(v)), atol=3e-5)


def test_lab_against_float64_formula_and_opencv():
    """sRGB -> CIE L*a*b* (D65) over the whole cube INCLUDING the dark colours, where the sRGB decode takes its linear segment
    (c <= 0.04045) and f(t) its linear branch (t <= (6/29)^3) — the reference's own vectors (red, grey, black) never enter
    those branches.  (1) the published formula evaluated in float64 with numpy: the f32 restatement agrees to 2e-4;
    (2) OpenCV's float path, an independent implementation (other matrix digits, table-interpolated gamma and cube root, good
    to ~0.5): agreement to 0.6 rules out a wrong white point, matrix or branch, which move Lab by whole units."""
    rng = np.random.default_rng(6)
    cols = np.concatenate([rng.integers(0, 256, (3000, 3)), rng.integers(0, 12, (500, 3)), rng.integers(0, 40, (500, 3)),
                           [[0, 0, 0], [1, 1, 1], [10, 10, 10], [11, 11, 11], [2, 0, 0], [0, 3, 0], [0, 0, 9], [255, 255, 255]]]).astype(np.uint8)
    got = np.array([orc.rgb8_to_lab(*c) for c in cols], np.float32)
    c = cols.astype(np.float64) / 255.0
    lin = np.where(c <= 0.04045, c / 12.92, ((c + 0.055) / 1.055) ** 2.4)
    m = np.array([[0.4124564, 0.3575761, 0.1804375], [0.2126729, 0.7151522, 0.0721750], [0.0193339, 0.1191920, 0.9503041]])
    xyz = lin @ m.T / np.array([0.95047, 1.0, 1.08883])
    f = np.where(xyz > (6 / 29) ** 3, np.cbrt(xyz), (841 / 108) * xyz + 4 / 29)
    want = np.stack([116 * f[:, 1] - 16, 500 * (f[:, 0] - f[:, 1]), 200 * (f[:, 1] - f[:, 2])], 1)
    assert np.abs(got - want).max() < 2e-4, np.abs(got - want).max()
    dark = cols.max(1) < 11                                              # every channel on the linear segment of the decode
    assert dark.sum() > 300 and (got[dark, 0] < 3.5).all() and (got[dark, 0] >= 0).all()
    assert ((xyz <= (6 / 29) ** 3).any(1)).sum() > 300                   # f(t)'s linear branch is really taken
    cv2 = pytest.importorskip("cv2")
    cv = cv2.cvtColor((cols.astype(np.float32) / np.float32(255.0)).reshape(-1, 1, 3), cv2.COLOR_RGB2Lab).reshape(-1, 3)
    assert np.abs(got - cv).max() < 0.6, np.abs(got - cv).max()


def test_topk_against_scipy_kdtree():
    """`nearest_n::<SquaredEuclidean>(query, k)` = the k smallest squared Euclidean distances, ascending (main.rs:348-352):
    the brute force and the bucketed k-d tree of the oracle against scipy's cKDTree (float64) — same neighbours in the same
    order wherever consecutive distances are further apart than float rounding, same distances to 1e-5."""
    spatial = pytest.importorskip("scipy.spatial")
    rng = np.random.default_rng(21)
    pts = np.stack([rng.uniform(0, 100, 3000), rng.uniform(-100, 100, 3000), rng.uniform(-100, 100, 3000)], 1).astype(np.float32)
    qs = np.stack([rng.uniform(0, 100, 200), rng.uniform(-100, 100, 200), rng.uniform(-100, 100, 200)], 1).astype(np.float32)
    k = 100
    dist, ref = spatial.cKDTree(pts.astype(np.float64)).query(qs.astype(np.float64), k=k)
    for idx, d2 in (orc.topk(pts, qs, k), orc.topk_kdtree(pts, qs, k)):
        assert np.allclose(d2, dist ** 2, rtol=1e-5)
        assert (np.diff(d2.astype(np.float64), axis=1) >= 0).all()
        gaps = np.diff(dist ** 2, axis=1)
        clear = np.concatenate([np.ones((len(qs), 1), bool), gaps > 1e-3], 1) & np.concatenate([gaps > 1e-3, np.ones((len(qs), 1), bool)], 1)
        assert clear.mean() > 0.9 and (idx[clear] == ref[clear]).all()
