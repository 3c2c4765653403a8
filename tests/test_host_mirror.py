"""The C++ host mirror of the reference's interface (include/mosaic_b200.hpp) — the reference is a Rust crate and this
image has no rustc, so the host side above the C ABI is C++ with the reference's names.  tests/cpp/test_host_mirror.cpp
restates the reference's own unit tests (similarity.rs, adjacency.rs, optimizer.rs, lib.rs) against it; this file
builds that binary, runs it, and checks MosaicGenerator end to end against the ctypes path (which the stage-wise parity
tests pin to the oracle), including the saved / reloaded / stale similarity-database cases of main.rs:120-143."""
import os
import subprocess

import numpy as np
import pytest

import mosaic_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "mosaic-art-rust_b200")
SRC = os.path.join(ROOT, "tests", "cpp", "test_host_mirror.cpp")
BIN = os.path.join(ROOT, "tests", "cpp", "_build", "test_host_mirror")


@pytest.fixture(scope="module")
def binary(lib_built):
    deps = [SRC, os.path.join(ROOT, "include", "mosaic_b200.hpp"), os.path.join(ROOT, "include", "mosaic_b200.h"),
            os.path.join(PKG, "libmosaic_b200.so")]
    if not os.path.exists(BIN) or any(os.path.getmtime(d) > os.path.getmtime(BIN) for d in deps):
        os.makedirs(os.path.dirname(BIN), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                               SRC, "-L", PKG, "-lmosaic_b200", "-Wl,-rpath," + PKG, "-o", BIN])
    return BIN


def _run(args, **kw):
    r = subprocess.run(args, capture_output=True, text=True, timeout=600, **kw)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    return r.stdout


def test_mirror_builds_and_host_only_tests_pass(binary, tmp_path):
    """Compiles with -Wall -Wextra -Werror against the shipped header, links every entry point the mirror uses, and runs
    the reference tests that need no device (bookkeeping, file format, loud failure without a GPU)."""
    out = _run([binary, "--cpu"], env=dict(os.environ, TMPDIR=str(tmp_path)))
    assert "0 failed checks" in out


@pytest.mark.gpu
def test_reference_unit_tests_through_the_mirror(binary, tmp_path):
    out = _run([binary], env=dict(os.environ, TMPDIR=str(tmp_path)))
    assert "0 failed checks" in out and "test_optimization_numerical_stability" in out


def _write_case(d, mats, tgt, gw, gh, max_usage, weight, strength, optimize, iters, seed, rebuild, db_path):
    mats.tofile(os.path.join(d, "materials.bin"))
    tgt.tofile(os.path.join(d, "target.bin"))
    n, mh, mw, _ = mats.shape
    with open(os.path.join(d, "meta.txt"), "w") as f:
        f.write(f"{n} {mw} {mh} {tgt.shape[1]} {tgt.shape[0]} {gw} {gh} {max_usage} {weight!r} {strength!r} "
                f"{int(optimize)} {iters} {seed} {int(rebuild)} {db_path or '-'}\n")


def _read_result(d, gw, gh):
    w, h, nfb, annealed, acc, imp, dblen = (int(v) for v in open(os.path.join(d, "report.txt")).read().split())
    placed = np.fromfile(os.path.join(d, "placed.bin"), np.int32)
    img = np.fromfile(os.path.join(d, "image.bin"), np.uint8).reshape(h, w, 3)
    assert placed.size == gw * gh
    return placed, img, dict(n_fallbacks=nfb, annealed=annealed, accepted=acc, improved=imp, db_len=dblen)


@pytest.mark.gpu
def test_mosaic_generator_matches_the_c_abi_path(binary, ctx, tmp_path):
    import mosaic_b200
    from mosaic_b200 import synth

    n, gw, gh = 300, 40, 24
    mats = synth.materials(n)
    tgt = synth.target(640, 384)
    W = 200.0  # a weight at which the penalty decides between candidates (so that stale database colours change the mosaic)
    args = dict(max_usage=3, weight=W, strength=0.3, optimize=True, iters=800, seed=42)

    def through_ctypes(sim_lab=None):
        ctx.set_materials(mats)
        ctx.set_target(tgt, gw, gh)
        ctx.material_lab()
        ctx.set_similarity_lab(sim_lab)
        placed = np.zeros(gw * gh, np.int32)
        img = np.zeros(ctx.output_shape(), np.uint8)
        rep = ctx.generate(max_usage=3, adjacency_weight=W, color_strength=0.3, enable_optimization=True,
                           optimization_iterations=800, seed=42, build_similarity=False, placed_out=placed, image_out=img)
        return placed, img, rep

    ref_placed, ref_img, ref_rep = through_ctypes()

    # 1. no database file
    d = str(tmp_path)
    _write_case(d, mats, tgt, gw, gh, rebuild=False, db_path=None, **args)
    _run([binary, "--generate", d])
    placed, img, rep = _read_result(d, gw, gh)
    assert (placed == ref_placed).all() and (img == ref_img).all()
    assert rep["n_fallbacks"] == ref_rep.n_fallbacks and rep["annealed"] == 1 and rep["accepted"] == ref_rep.opt.accepted_count

    # 2. the file does not exist: built, saved (main.rs:120-134); second run loads it (135-137): same mosaic both times
    db = os.path.join(d, "similarity_db.json")
    _write_case(d, mats, tgt, gw, gh, rebuild=False, db_path=db, **args)
    _run([binary, "--generate", d])
    assert os.path.exists(db)
    p1, i1, _ = _read_result(d, gw, gh)
    _run([binary, "--generate", d])
    p2, i2, r2 = _read_result(d, gw, gh)
    assert (p1 == ref_placed).all() and (p2 == ref_placed).all() and (i1 == ref_img).all() and (i2 == ref_img).all()
    assert r2["db_len"] == n

    # 3. a stale file: other colours for the known paths, two paths missing (topped up), one foreign path —
    #    similarities follow the FILE's colours, matching the tiles' (main.rs:135-143)
    sdb = mosaic_b200.SimilarityDb.load(db)
    lab, sim = sdb.arrays()
    paths = [sdb.path(i) for i in range(n)]
    sdb.close()
    stale = (lab + np.random.default_rng(3).normal(0, 8.0, lab.shape)).astype(np.float32)
    keep = [i for i in range(n) if i not in (7, 123)]
    file_paths = [paths[i] for i in keep] + ["elsewhere/foreign.png"]
    file_lab = np.concatenate([stale[keep], np.array([[10.0, 20.0, -30.0]], np.float32)])
    m = len(file_paths)
    mosaic_b200.SimilarityDb.save(db, file_paths, file_lab, np.zeros(m * (m - 1) // 2, np.float32))
    _run([binary, "--generate", d])
    p3, i3, r3 = _read_result(d, gw, gh)
    assert r3["db_len"] == n + 1
    expect_lab = stale.copy()
    tile_lab = ctx.material_lab()
    expect_lab[[7, 123]] = tile_lab[[7, 123]]
    e_placed, e_img, _ = through_ctypes(expect_lab)
    assert (p3 == e_placed).all() and (i3 == e_img).all()
    assert (e_placed != ref_placed).any(), "the stale colours must change the mosaic for this test to mean anything"

    # 4. --rebuild-similarity-db ignores the stale file and overwrites it
    _write_case(d, mats, tgt, gw, gh, rebuild=True, db_path=db, **args)
    _run([binary, "--generate", d])
    p4, i4, r4 = _read_result(d, gw, gh)
    assert (p4 == ref_placed).all() and (i4 == ref_img).all() and r4["db_len"] == n


# ---------------------------------------------------------------- the C++ host CLI (reference flag names)
CLI = os.path.join(PKG, "host", "mosaic-b200")


@pytest.fixture(scope="module")
def cli(lib_built):
    subprocess.check_call(["make", "-C", PKG, "host/mosaic-b200"], stdout=subprocess.DEVNULL)
    return CLI


def _write_ppm(path, img):
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
        f.write(np.ascontiguousarray(img, np.uint8).tobytes())


def _read_ppm(path):
    with open(path, "rb") as f:
        assert f.readline() == b"P6\n"
        w, h = (int(v) for v in f.readline().split())
        assert f.readline() == b"255\n"
        return np.frombuffer(f.read(), np.uint8).reshape(h, w, 3)


def test_cli_flags_and_errors(cli, tmp_path):
    r = subprocess.run([cli], capture_output=True, text=True)
    assert r.returncode == 2 and "--target" in r.stderr
    r = subprocess.run([cli, "-t", "a", "-m", "b", "-o", "c", "--no-such-flag", "1"], capture_output=True, text=True)
    assert r.returncode == 2 and "unexpected argument" in r.stderr
    r = subprocess.run([cli, "-t", str(tmp_path / "none.ppm"), "-m", str(tmp_path), "-o", str(tmp_path / "o.ppm")],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "cannot read target" in r.stderr
    # clap's value parsers: a malformed number is a usage error (exit code 2), never an exception out of main
    for flag, value in (("--grid-w", "abc"), ("--max-materials", "-5"), ("--aspect-tolerance", "xyz"), ("--grid-h", "99999999999"),
                        ("--seed", "1e3"), ("--aspect-tolerance", "0x1p3"), ("--optimization-iterations", "12x"),
                        ("--adjacency-penalty-weight", ""), ("--max-usage-per-image", " 3")):
        r = subprocess.run([cli, "-t", "a", "-m", "b", "-o", "c.png", flag, value], capture_output=True, text=True)
        assert r.returncode == 2 and "invalid value" in r.stderr and flag in r.stderr, (flag, value, r.returncode, r.stderr)
    r = subprocess.run([cli, "-t", str(tmp_path / "none.ppm"), "-m", str(tmp_path), "-o", "c.png", "--aspect-tolerance", "-0.1",
                        "--grid-w", "+5", "--color-adjustment-strength", "1e-1"], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot read target" in r.stderr       # well-formed values get past the parser


def _reference_load_tiles(files, target_aspect, tolerance, max_materials):
    """load_tiles of the reference (main.rs:166-261) over `files` = [(name, aspect or None if the decode fails)] in the
    order the CLI takes them (sorted)."""
    f32 = np.float32
    entries = [(n, a) for n, a in files
               if n.rsplit(".", 1)[-1].lower() in ("png", "jpg", "jpeg", "webp", "ppm") and "." in n]
    tiles = [n for n, a in entries if a is not None and abs(f32(a) - f32(target_aspect)) <= f32(tolerance)]
    if not tiles:
        tiles = [n for n, a in entries[:min(len(entries), 2 * max_materials)] if a is not None]
    return tiles[:max_materials]


def test_cli_selects_the_tiles_load_tiles_selects(cli, tmp_path):
    """The material selection alone (--list-tiles stops before anything needs a device): extension filter, aspect filter
    with the f32 comparison of lib.rs:121-123, failures dropped BEFORE the truncation, the no-match fallback."""
    from PIL import Image
    rng = np.random.default_rng(5)
    tgt = rng.integers(0, 256, (90, 160, 3), dtype=np.uint8)         # 16:9
    Image.fromarray(tgt).save(tmp_path / "target.png")

    def listed(mdir, *flags):
        r = subprocess.run([cli, "--target", str(tmp_path / "target.png"), "--material-src", str(mdir), "--list-tiles", *flags],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        return [os.path.basename(line.split()[1]) for line in r.stdout.splitlines() if line.startswith("tile: ")], r

    # ---- a directory with matching, non-matching, damaged and foreign files
    mdir = tmp_path / "mixed"
    mdir.mkdir()
    files = []

    def add(name, w, h, damage=None, fmt=None):
        img = Image.fromarray(rng.integers(0, 256, (h, w, 3), dtype=np.uint8))
        path = mdir / name
        img.save(path, format=fmt) if fmt else img.save(path)
        if damage == "body":                                          # header intact, data cut off
            data = path.read_bytes()
            path.write_bytes(data[:len(data) // 2])
        elif damage == "all":
            path.write_bytes(b"not an image")
        files.append((name, None if damage else w / h))

    add("a_wide.png", 64, 36)
    add("b_square.png", 48, 48)
    add("c_wide_cut.png", 64, 36, damage="body")                      # matches the aspect, cannot be decoded: takes no slot
    add("d_wide.JPG", 128, 72, fmt="JPEG")                            # extension matched case-insensitively
    add("e_garbage.jpeg", 10, 10, damage="all")
    add("f_almost.webp", 66, 36, fmt="WEBP")                          # 1.833: inside the default tolerance 0.1 of 1.778
    add("g_outside.png", 68, 36)                                      # 1.889: outside
    add("h_edge.png", 169, 90)                                        # 1.8778: |d| = 0.1 to within f32 rounding
    add("i_wide.png", 32, 18)
    add("j_wide.png", 16, 9)
    (mdir / "notes.txt").write_text("text")
    (mdir / "noext").write_bytes((mdir / "a_wide.png").read_bytes())
    (mdir / "k_dir.png").mkdir()                                      # is_file() fails
    files.sort()
    mixed_files = list(files)
    for max_materials in (500, 4, 3, 1):
        got, _ = listed(mdir, "--max-materials", str(max_materials))
        assert got == _reference_load_tiles(files, 160 / 90, 0.1, max_materials), max_materials
    got, r = listed(mdir, "--max-materials", "3")
    assert got == ["a_wide.png", "d_wide.JPG", "f_almost.webp"]      # c_wide_cut did not use up one of the three
    assert "c_wide_cut.png" in r.stderr and "Loaded 3 tiles" in r.stdout
    got, _ = listed(mdir, "--aspect-tolerance", "0.0")
    assert got == _reference_load_tiles(files, 160 / 90, 0.0, 500) == ["a_wide.png", "d_wide.JPG", "i_wide.png", "j_wide.png"]

    # ---- nothing matches: the fallback takes the first min(len, 2 * max) files, drops failures, keeps max
    sq = tmp_path / "square"
    sq.mkdir()
    files = []
    for i in range(9):
        name = f"s{i}.png"
        Image.fromarray(rng.integers(0, 256, (20, 20, 3), dtype=np.uint8)).save(sq / name)
        files.append((name, 1.0))
    (sq / "s1.png").write_bytes((sq / "s1.png").read_bytes()[:40])   # damaged inside the sampled range
    files[1] = ("s1.png", None)
    for max_materials in (2, 3, 4, 100):
        got, r = listed(sq, "--max-materials", str(max_materials))
        assert got == _reference_load_tiles(files, 160 / 90, 0.1, max_materials), max_materials
        assert "No tiles matched target aspect ratio 1.778" in r.stdout
    got, _ = listed(sq, "--max-materials", "2")
    assert got == ["s0.png", "s2.png"]

    # ---- every matching file is damaged: the reference's `tiles.is_empty()` is true AFTER decoding -> fallback
    bad = tmp_path / "allbad"
    bad.mkdir()
    Image.fromarray(rng.integers(0, 256, (36, 64, 3), dtype=np.uint8)).save(bad / "wide.png")
    (bad / "wide.png").write_bytes((bad / "wide.png").read_bytes()[:60])
    Image.fromarray(rng.integers(0, 256, (30, 30, 3), dtype=np.uint8)).save(bad / "square.png")
    got, r = listed(bad)
    assert got == ["square.png"] and "No tiles matched" in r.stdout

    # ---- the reference's own edge cases: main.rs:1987-2022
    got, r = listed(mdir, "--max-materials", "0")                     # test_zero_max_materials: nothing, and no fallback
    assert got == [] and "No tiles matched" not in r.stdout and "Loaded 0 tiles" in r.stdout
    got, r = listed(mdir, "--aspect-tolerance=-0.1")                  # test_negative_aspect_tolerance: nothing matches ...
    assert "No tiles matched" in r.stdout                             # ... so load_tiles falls back to the unfiltered files
    assert got == _reference_load_tiles(mixed_files, 160 / 90, -0.1, 500) and len(got) == 8
    got, r = listed(mdir, "--aspect-tolerance", "1000")               # test_extreme_aspect_tolerance: everything matches
    assert "No tiles matched" not in r.stdout
    assert got == _reference_load_tiles(mixed_files, 160 / 90, 1000.0, 500) and len(got) == 8

    # ---- PPM headers that lie (found with the CLI under AddressSanitizer: 99999999999 x 99999999999 went to atoi and resize)
    pp = tmp_path / "ppm"
    pp.mkdir()
    px = bytes(range(48))
    for name, data in {"ok.ppm": b"P6\n4 4\n255\n" + px, "comment.ppm": b"P6\n# c\n4 # x\n4\n255\n" + px, "nows.ppm": b"P6 4 4 255 " + px,
                       "crlf.ppm": b"P6\r\n4 4\r\n255\r\n" + px, "short.ppm": b"P6\n4 4\n255\n" + px[:10],
                       "huge.ppm": b"P6\n99999999999 99999999999\n255\n" + px, "big.ppm": b"P6\n65536 65536\n255\n" + px,
                       "neg.ppm": b"P6\n-4 4\n255\n" + px, "zero.ppm": b"P6\n0 4\n255\n", "maxval.ppm": b"P6\n4 4\n65535\n" + px + px,
                       "p3.ppm": b"P3\n1 1\n255\n1 2 3\n", "empty.ppm": b"", "p6only.ppm": b"P6", "nonum.ppm": b"P6\nabc def\n255\n" + px,
                       "eofincomment.ppm": b"P6\n# never ends"}.items():
        (pp / name).write_bytes(data)
    got, r = listed(pp, "--aspect-tolerance", "100")
    assert got == ["comment.ppm", "crlf.ppm", "nows.ppm", "ok.ppm"]
    for name in ("huge.ppm", "big.ppm", "short.ppm", "neg.ppm", "zero.ppm", "maxval.ppm", "p3.ppm", "empty.ppm"):
        assert name in r.stderr

    # ---- an empty directory lists nothing
    empty = tmp_path / "empty"
    empty.mkdir()
    got, _ = listed(empty)
    assert got == []


@pytest.mark.gpu
def test_cli_end_to_end(cli, ctx, tmp_path):
    from mosaic_b200 import synth

    n, gw, gh = 60, 20, 12
    mats = synth.materials(n + 5)
    tgt = synth.target(320, 192)
    mdir = tmp_path / "materials"
    mdir.mkdir()
    for i in range(n + 5):
        _write_ppm(mdir / f"m{i:04d}.ppm", mats[i])
    (mdir / "notes.txt").write_text("not an image")
    _write_ppm(tmp_path / "target.ppm", tgt)
    out = tmp_path / "out.ppm"
    db = tmp_path / "similarity_db.json"
    # square materials against a 5:3 target: the aspect filter matches nothing -> fallback takes min(len, 2*max) files,
    # then truncate to --max-materials (main.rs:211-257)
    stdout = _run([cli, "--target", str(tmp_path / "target.ppm"), "--material-src", str(mdir), "--output", str(out),
                   "--grid-w", str(gw), "--grid-h", str(gh), "--max-materials", str(n), "--max-usage-per-image", "4",
                   "--adjacency-penalty-weight", "0.25", "--color-adjustment-strength", "0.4",
                   "--optimization-iterations", "300", "--similarity-db", str(db), "--seed", "7"])
    assert "No tiles matched target aspect ratio" in stdout and f"Loaded {n} tiles" in stdout
    assert db.exists()
    ctx.set_materials_ragged(list(mats[:n]))  # the CLI hands the files over as materials of individual sizes
    ctx.set_target(tgt, gw, gh)
    img = np.zeros(ctx.output_shape(), np.uint8)
    ctx.generate(max_usage=4, adjacency_weight=0.25, color_strength=0.4, enable_optimization=True,
                 optimization_iterations=300, seed=7, build_similarity=False, image_out=img)
    assert (_read_ppm(out) == img).all()
    # --enable-optimization false, database reloaded from the file
    stdout = _run([cli, "-t", str(tmp_path / "target.ppm"), "-m", str(mdir), "-o", str(out), "--grid-w", str(gw),
                   "--grid-h", str(gh), "--max-materials", str(n), "--max-usage-per-image", "4",
                   "--adjacency-penalty-weight", "0.25", "--color-adjustment-strength", "0.4",
                   "--enable-optimization", "false", "--similarity-db", str(db)])
    assert "Optimization complete" not in stdout
    ctx.generate(max_usage=4, adjacency_weight=0.25, color_strength=0.4, enable_optimization=False,
                 optimization_iterations=0, seed=0, build_similarity=False, image_out=img)
    assert (_read_ppm(out) == img).all()

    # a directory of mixed sizes: the aspect filter keeps the 5:3 files (80x48, 50x30, 200x120), drops the squares
    mixed = tmp_path / "mixed"
    mixed.mkdir()
    kept = []
    for i in range(90):
        w, h = [(64, 64), (80, 48), (50, 30), (200, 120), (30, 70)][i % 5]
        m = synth.materials(1, w=w, h=h, start=i)[0]
        # PNG and PPM side by side, the extension matched case-insensitively (main.rs:175-184); same sorted order
        if i % 3 == 0:
            _write_ppm(mixed / f"p{i:04d}.ppm", m)
        else:
            mosaic_b200.png_encode(mixed / f"p{i:04d}.{'PNG' if i % 3 == 1 else 'png'}", m)
        if abs(w / h - 320 / 192) <= 0.1:
            kept.append(m)
    mosaic_b200.png_encode(tmp_path / "target.png", tgt)
    out_png = tmp_path / "sub" / "out.png"
    out_png.parent.mkdir()
    stdout = _run([cli, "-t", str(tmp_path / "target.png"), "-m", str(mixed), "-o", str(out_png), "--grid-w", str(gw),
                   "--grid-h", str(gh), "--max-materials", "50", "--optimization-iterations", "200",
                   "--similarity-db", str(tmp_path / "db2.json"), "--seed", "3"])
    assert "No tiles matched" not in stdout and "Loaded 50 tiles" in stdout
    ctx.set_materials_ragged(kept[:50])
    ctx.set_target(tgt, gw, gh)
    ctx.generate(max_usage=3, adjacency_weight=0.3, color_strength=0.3, enable_optimization=True,
                 optimization_iterations=200, seed=3, build_similarity=False, image_out=img)
    assert (mosaic_b200.png_decode(out_png) == img).all()   # decode(file) -> device -> encode(file): lossless end to end
    try:
        from PIL import Image
        assert (np.asarray(Image.open(out_png).convert("RGB")) == img).all()   # an independent decoder reads the same pixels
    except ImportError:
        pass


def test_cli_material_selection_follows_load_tiles(cli, tmp_path):
    """The selection logic of load_tiles (main.rs:166-261) runs before any device work, so it is checked on CPU: PPM
    headers with comments, the aspect filter (lib.rs:121-123), the no-match fallback to min(len, 2 * max) files, the
    truncation to --max-materials, files that are not images.  Without a GPU the run then stops with the library's loud
    error (no CPU fallback); with one it completes."""
    rng = np.random.default_rng(0)
    tgt = rng.integers(0, 256, (60, 100, 3), dtype=np.uint8)                   # aspect 1.667
    with open(tmp_path / "target.ppm", "wb") as f:
        f.write(b"P6\n# a comment line\n100 60\n# another\n255\n" + tgt.tobytes())
    mdir = tmp_path / "m"
    mdir.mkdir()
    sizes = [(50, 30), (64, 64), (25, 15), (30, 70), (100, 60), (10, 6), (48, 48)]  # 4 of 7 match 5:3 within 0.1
    for i in range(21):
        w, h = sizes[i % 7]
        _write_ppm(mdir / f"f{i:02d}.ppm", rng.integers(0, 256, (h, w, 3), dtype=np.uint8))
    (mdir / "readme.txt").write_text("x")
    (mdir / "broken.ppm").write_bytes(b"P5\n2 2\n255\n0000")                    # wrong magic: reported, skipped
    (mdir / "broken.png").write_bytes(b"\x89PNG but not really")                 # bad signature: reported, skipped
    for i, (w, h) in enumerate(sizes):                                           # the same sizes once more, as PNG files
        mosaic_b200.png_encode(mdir / f"g{i:02d}.{'PNG' if i % 2 else 'png'}", rng.integers(0, 256, (h, w, 3), dtype=np.uint8))

    def run(material_dir, *extra):
        r = subprocess.run([cli, "-t", str(tmp_path / "target.ppm"), "-m", str(material_dir), "-o", str(tmp_path / "o.ppm"),
                            "--grid-w", "5", "--grid-h", "3", "--similarity-db", str(tmp_path / "db.json"), *extra],
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 or "no usable CUDA device" in r.stderr, r.stderr[-1500:]
        return r

    r = run(mdir)
    assert "Target aspect ratio: 1.667" in r.stdout
    assert "Loaded 16 tiles" in r.stdout and "No tiles matched" not in r.stdout   # (3 rounds of PPM + 1 of PNG) x 4 matching sizes
    assert "broken.ppm" in r.stderr and "not a PNG / JPEG / WebP / binary PPM" in r.stderr
    assert "broken.png" in r.stderr
    r = run(mdir, "--max-materials", "5")
    assert "Loaded 5 tiles" in r.stdout                                            # truncate (main.rs:255-257)
    r = run(mdir, "--aspect-tolerance", "0.001", "--max-materials", "4")
    assert "Loaded 4 tiles" in r.stdout   # only exact 5:3 files pass the filter (50x30, 25x15, 100x60, 10x6 -> 12), truncated
    sq = tmp_path / "squares"
    sq.mkdir()
    for i in range(9):
        _write_ppm(sq / f"s{i}.ppm", rng.integers(0, 256, (16, 16, 3), dtype=np.uint8))
    r = run(sq, "--max-materials", "3")
    assert "No tiles matched target aspect ratio 1.667" in r.stdout               # fallback: min(9, 2 * 3) = 6 files ...
    assert "Loaded 3 tiles" in r.stdout                                            # ... truncated to 3
    empty = tmp_path / "empty"
    empty.mkdir()
    r = subprocess.run([cli, "-t", str(tmp_path / "target.ppm"), "-m", str(empty), "-o", str(tmp_path / "o.ppm")],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "no material images found" in r.stderr
