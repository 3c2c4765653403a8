"""WebP in and out on the host side (include/mosaic_webp.hpp through the C ABI): what
`image::open(path)?.to_rgb8()` (main.rs:263-301; main.rs:181 lists `webp` among the material extensions) and
`output_img.save("x.webp")` (main.rs:668-672) do.

CPU only.  Pillow reads and writes WebP through libwebp, an independent implementation:
* VP8L (lossless) files written by libwebp at every effort level — predictor / cross-colour / subtract-green / palette
  transforms with pixel bundling, colour cache, meta prefix codes, backward references — must decode to the pixels that
  went in;
* VP8 (lossy) files must decode to EXACTLY libwebp's pixels: RFC 6386 fixes the YUV planes, and the decoder takes libwebp's
  default way to RGB (fancy upsampling, its fixed-point conversion).  (Against the reference's own decoder, image-webp,
  parity is unpinned — see the header.)
* files this library writes (lossless VP8L) must decode, with libwebp, to the pixels that went in."""
import numpy as np
import pytest

import mosaic_b200

PIL = pytest.importorskip("PIL.Image")
from PIL import features  # noqa: E402

pytestmark = pytest.mark.skipif(not features.check("webp"), reason="this Pillow has no WebP support to compare with")


def images():
    rng = np.random.default_rng(20261018)
    yy, xx = np.mgrid[0:70, 0:101]
    smooth = np.stack([(xx * 2) % 256, (yy * 3 + xx) % 256, (xx * yy // 7) % 256], -1).astype(np.uint8)
    noise = rng.integers(0, 256, (53, 67, 3), dtype=np.uint8)
    ph = rng.integers(0, 256, (200, 300, 3)).astype(np.float64)
    for _ in range(4):  # blurred noise: photo-like statistics (every intra mode, both filter kinds of edges)
        ph = (ph + np.roll(ph, 1, 0) + np.roll(ph, 1, 1) + np.roll(ph, -1, 0) + np.roll(ph, -1, 1)) / 5
    photo = ((ph - ph.min()) / (ph.max() - ph.min()) * 255).astype(np.uint8)
    return smooth, noise, photo


def roundtrip(tmp_path, arr, name, **kw):
    f = tmp_path / f"{name}.webp"
    PIL.fromarray(arr).save(f, **kw)
    want = np.asarray(PIL.open(f).convert("RGB"))
    got = mosaic_b200.webp_decode(f)
    assert got.shape == want.shape, name
    assert (got == want).all(), f"{name} {kw}: {np.abs(got.astype(int) - want.astype(int)).max()} max difference to libwebp"
    return got


@pytest.mark.parametrize("quality,method", [(0, 0), (50, 3), (75, 4), (100, 6)])
def test_lossless_decode(tmp_path, quality, method):
    smooth, noise, photo = images()
    rng = np.random.default_rng(quality)
    kw = dict(lossless=True, quality=quality, method=method)
    for name, a in (("smooth", smooth), ("noise", noise), ("photo", photo), ("grey", smooth[:, :, 1])):
        got = roundtrip(tmp_path, a, name, **kw)
        assert (got == (a if a.ndim == 3 else np.repeat(a[:, :, None], 3, 2))).all()  # lossless: the pixels that went in
    pal = rng.integers(0, 256, (300, 3), dtype=np.uint8)
    for ncol in (2, 3, 4, 5, 16, 17, 200):  # colour indexing with 8, 4, 2 and 1 pixels per packed pixel, and without
        a = pal[rng.integers(0, ncol, (41, 59))]
        assert (roundtrip(tmp_path, a, f"pal{ncol}", **kw) == a).all()
    big = np.tile(smooth, (5, 6, 1))[:333, :517]  # long backward references, several prefix-code groups
    assert (roundtrip(tmp_path, big, "big", **kw) == big).all()
    rgba = np.dstack([noise, rng.integers(0, 256, noise.shape[:2], dtype=np.uint8)])
    assert (roundtrip(tmp_path, rgba, "rgba", exact=True, **kw) == noise).all()  # alpha dropped, not blended


@pytest.mark.parametrize("quality,method", [(80, 4), (10, 0), (50, 2), (95, 6), (100, 6), (0, 3)])
def test_lossy_decode_equals_libwebp(tmp_path, quality, method):
    for name, a in zip(("smooth", "noise", "photo"), images()):
        roundtrip(tmp_path, a, name, quality=quality, method=method)


@pytest.mark.parametrize("shape", [(1, 1), (1, 9), (9, 1), (3, 16), (16, 3), (15, 17), (16, 16), (17, 33), (32, 48), (2, 2), (31, 1)])
def test_small_sizes(tmp_path, shape):
    _, noise, photo = images()
    for q in (30, 90):
        roundtrip(tmp_path, np.ascontiguousarray(photo[:shape[0], :shape[1]]), "p", quality=q)
        roundtrip(tmp_path, np.ascontiguousarray(noise[:shape[0], :shape[1]]), "n", quality=q)
    roundtrip(tmp_path, np.ascontiguousarray(noise[:shape[0], :shape[1]]), "l", lossless=True)


def test_lossy_variants(tmp_path):
    smooth, noise, photo = images()
    rng = np.random.default_rng(5)
    rgba = np.dstack([photo[:100, :150], rng.integers(0, 256, (100, 150), dtype=np.uint8)])
    rgba[:20, :20, 3] = 0
    for q in (20, 75):  # extended container (VP8X + ALPH): the colour planes decode alike, alpha is dropped
        roundtrip(tmp_path, rgba, "rgba", quality=q)
        roundtrip(tmp_path, rgba, "rgba_exact", quality=q, exact=True)
        roundtrip(tmp_path, rgba, "rgba_aq", quality=q, alpha_quality=30)
    roundtrip(tmp_path, photo[:, :, 0], "grey", quality=70)
    big = np.tile(photo, (4, 4, 1))[:777, :1111]
    big = (big.astype(int) + rng.integers(-20, 20, big.shape)).clip(0, 255).astype(np.uint8)
    for q in (5, 40, 85):
        roundtrip(tmp_path, big, "big", quality=q, method=4)
    edges = np.zeros((120, 160, 3), np.uint8)
    edges[::7] = 255
    edges[:, ::5, 1] = 200
    edges[30:60, 40:100] = (255, 0, 0)
    for q in (1, 25, 60, 99):  # saturating reconstruction, strong loop filtering
        roundtrip(tmp_path, edges, "edges", quality=q)
    f = tmp_path / "icc.webp"  # ICCP chunk in front of the image chunk
    PIL.fromarray(smooth).save(f, quality=80, icc_profile=b"\0" * 132)
    assert (mosaic_b200.webp_decode(f) == np.asarray(PIL.open(f).convert("RGB"))).all()


def test_encode_is_lossless(tmp_path):
    smooth, noise, photo = images()
    rng = np.random.default_rng(3)
    cases = {"smooth": smooth, "noise": noise, "photo": photo, "flat": np.full((9, 13, 3), 77, np.uint8),
             "one": np.array([[[1, 2, 3]]], np.uint8), "grey": np.repeat(rng.integers(0, 256, (31, 17, 1), dtype=np.uint8), 3, 2),
             "two": (rng.integers(0, 2, (20, 20, 3)) * 255).astype(np.uint8), "uniform": rng.integers(0, 256, (600, 900, 3), dtype=np.uint8),
             "skew": rng.exponential(3, (300, 300, 3)).clip(0, 255).astype(np.uint8)}
    for name, a in cases.items():
        f = tmp_path / f"{name}.webp"
        mosaic_b200.webp_encode(f, a)
        im = PIL.open(f)
        assert im.format == "WEBP" and im.size == (a.shape[1], a.shape[0]), name
        assert (np.asarray(im.convert("RGB")) == a).all(), name  # libwebp reads what went in
        assert (mosaic_b200.webp_decode(f) == a).all(), name
        assert (mosaic_b200.image_decode(f) == a).all(), name


def test_errors_and_dispatch(tmp_path):
    smooth, _, _ = images()
    f = tmp_path / "a.webp"
    PIL.fromarray(smooth).save(f, quality=80)
    data = f.read_bytes()
    assert (mosaic_b200.image_decode(f) == mosaic_b200.webp_decode(f)).all()  # told apart by the first bytes
    (tmp_path / "notwebp.webp").write_bytes(b"RIFF\x04\0\0\0WAVE")
    with pytest.raises(mosaic_b200.MosaicError):
        mosaic_b200.webp_decode(tmp_path / "notwebp.webp")
    (tmp_path / "cut.webp").write_bytes(data[:25])
    with pytest.raises(mosaic_b200.MosaicError):
        mosaic_b200.webp_decode(tmp_path / "cut.webp")
    with pytest.raises(mosaic_b200.MosaicError):
        mosaic_b200.webp_decode(tmp_path / "missing.webp")
    anim = tmp_path / "anim.webp"
    frames = [PIL.fromarray(np.roll(smooth, 5 * i, 1)) for i in range(3)]
    frames[0].save(anim, save_all=True, append_images=frames[1:], duration=50)
    with pytest.raises(mosaic_b200.MosaicError, match="animated"):
        mosaic_b200.webp_decode(anim)
    lossless = tmp_path / "l.webp"
    PIL.fromarray(smooth).save(lossless, lossless=True)
    raw = bytearray(lossless.read_bytes())
    for cut in (30, len(raw) // 2, len(raw) - 3):  # truncated lossless streams: an error, never a crash
        (tmp_path / "t.webp").write_bytes(bytes(raw[:cut]))
        with pytest.raises(mosaic_b200.MosaicError):
            mosaic_b200.webp_decode(tmp_path / "t.webp")
    rng = np.random.default_rng(1)
    for trial in range(200):  # corrupt bytes: a clean error or some image, never a crash
        bad = bytearray(raw if trial % 2 else data)
        for _ in range(3):
            bad[int(rng.integers(12, len(bad)))] = int(rng.integers(0, 256))
        (tmp_path / "bad.webp").write_bytes(bytes(bad))
        try:
            mosaic_b200.webp_decode(tmp_path / "bad.webp")
        except mosaic_b200.MosaicError:
            pass


def test_cli_selects_webp_materials(tmp_path):
    """The host CLI's load_tiles (main.rs:166-261) over a directory of WebP files (lossy, lossless, with alpha): extension
    matched case-insensitively, dimensions from the bitstream header for the aspect filter; `-o x.webp` asks for a WebP
    mosaic.  (Without a GPU the run then stops with the library's loud error; with one it completes.)"""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "mosaic-art-rust_b200")
    subprocess.check_call(["make", "-C", pkg, "host/mosaic-b200"], stdout=subprocess.DEVNULL)
    cli = os.path.join(pkg, "host", "mosaic-b200")
    rng = np.random.default_rng(3)
    PIL.fromarray(rng.integers(0, 256, (60, 100, 3), dtype=np.uint8)).save(tmp_path / "target.WEBP", quality=90)
    mdir = tmp_path / "m"
    mdir.mkdir()
    sizes = [(50, 30), (64, 64), (25, 15), (30, 70), (100, 60)]  # 3 of 5 match 5:3
    for i in range(10):
        w, h = sizes[i % 5]
        a = rng.integers(0, 256, (h, w, 4 if i % 4 == 3 else 3), dtype=np.uint8)
        PIL.fromarray(a).save(mdir / f"f{i:02d}.{'WebP' if i % 2 else 'webp'}", quality=85, lossless=bool(i % 3 == 0))
    (mdir / "broken.webp").write_bytes(b"RIFF\x10\0\0\0WEBPVP8 \x04\0\0\0abcd")
    r = subprocess.run([cli, "-t", str(tmp_path / "target.WEBP"), "-m", str(mdir), "-o", str(tmp_path / "o.webp"), "--grid-w", "5",
                        "--grid-h", "3", "--similarity-db", str(tmp_path / "db.json")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 or "no usable CUDA device" in r.stderr, r.stderr[-1500:]
    assert "Target aspect ratio: 1.667" in r.stdout
    assert "Loaded 6 tiles" in r.stdout          # 2 rounds x 3 matching sizes
    assert "broken.webp" in r.stderr
    if r.returncode == 0:
        out = np.asarray(PIL.open(tmp_path / "o.webp").convert("RGB"))
        assert out.shape == (60, 100, 3)
        assert (out == mosaic_b200.webp_decode(tmp_path / "o.webp")).all()
