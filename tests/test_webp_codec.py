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
    for cut in (len(data) // 3, len(data) // 2, len(data) - 40):  # a lossy stream that ends early: an error, not a picture of zeros
        (tmp_path / "t.webp").write_bytes(data[:cut])
        with pytest.raises(mosaic_b200.MosaicError, match="truncated"):
            mosaic_b200.webp_decode(tmp_path / "t.webp")
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



def _libwebp_encode(rgb, quality=75.0, **settings):
    """libwebp's own encoder through its public C API (WebPConfig / WebPPicture / WebPEncode, encode.h), for the bitstream
    features Pillow's `save` has no switch for: the SIMPLE loop filter, filter sharpness, several token partitions, the
    segment count.  Uses the libwebp that Pillow bundles; None when it cannot be found."""
    import ctypes as C
    import glob
    import os
    import PIL as pil_pkg
    libdir = os.path.join(os.path.dirname(os.path.dirname(pil_pkg.__file__)), "pillow.libs")
    libs = sorted(glob.glob(os.path.join(libdir, "libwebp-*.so*")))
    if not libs:
        return None
    for dep in glob.glob(os.path.join(libdir, "libsharpyuv-*.so*")):
        C.CDLL(dep, mode=C.RTLD_GLOBAL)
    L = C.CDLL(libs[0])
    ints = ["lossless", None, "method", "image_hint", "target_size", None, "segments", "sns_strength", "filter_strength", "filter_sharpness",
            "filter_type", "autofilter", "alpha_compression", "alpha_filtering", "alpha_quality", "pass_", "show_compressed", "preprocessing",
            "partitions", "partition_limit", "emulate_jpeg_size", "thread_level", "low_memory", "near_lossless", "exact", "use_delta_palette",
            "use_sharp_yuv", "qmin", "qmax"]
    ints[1], ints[5] = "quality", "target_PSNR"

    class Config(C.Structure):
        _fields_ = [(n, C.c_float if n in ("quality", "target_PSNR") else C.c_int) for n in ints] + [("pad", C.c_int * 8)]

    writer_t = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_uint8), C.c_size_t, C.c_void_p)

    class Picture(C.Structure):
        _fields_ = [("use_argb", C.c_int), ("colorspace", C.c_int), ("width", C.c_int), ("height", C.c_int), ("y", C.c_void_p), ("u", C.c_void_p),
                    ("v", C.c_void_p), ("y_stride", C.c_int), ("uv_stride", C.c_int), ("a", C.c_void_p), ("a_stride", C.c_int), ("pad1", C.c_uint32 * 2),
                    ("argb", C.c_void_p), ("argb_stride", C.c_int), ("pad2", C.c_uint32 * 3), ("writer", writer_t), ("custom_ptr", C.c_void_p),
                    ("extra_info_type", C.c_int), ("extra_info", C.c_void_p), ("stats", C.c_void_p), ("error_code", C.c_int), ("progress_hook", C.c_void_p),
                    ("user_data", C.c_void_p), ("pad3", C.c_uint32 * 3), ("pad4", C.c_void_p), ("pad5", C.c_void_p), ("pad6", C.c_uint32 * 8),
                    ("memory_", C.c_void_p), ("memory_argb_", C.c_void_p), ("pad7", C.c_void_p * 2), ("slack", C.c_uint8 * 256)]

    abi = 0x020F  # (the library checks the major half only)
    cfg = Config()
    if not L.WebPConfigInitInternal(C.byref(cfg), 0, C.c_float(quality), abi):
        return None
    if (cfg.segments, cfg.filter_strength, cfg.filter_type, cfg.near_lossless, cfg.qmax) != (4, 60, 1, 100, 100):
        return None  # not the struct layout this helper was written against
    for k, v in settings.items():
        setattr(cfg, k, v)
    assert L.WebPValidateConfig(C.byref(cfg)), settings
    pic = Picture()
    assert L.WebPPictureInitInternal(C.byref(pic), abi)
    rgb = np.ascontiguousarray(rgb, np.uint8)
    pic.width, pic.height = rgb.shape[1], rgb.shape[0]
    assert L.WebPPictureImportRGB(C.byref(pic), rgb.ctypes.data_as(C.c_void_p), rgb.shape[1] * 3)
    out = bytearray()

    def sink(data, size, _pic):
        out.extend(C.string_at(data, size))
        return 1

    cb = writer_t(sink)
    pic.writer = cb
    ok = L.WebPEncode(C.byref(cfg), C.byref(pic))
    L.WebPPictureFree(C.byref(pic))
    assert ok, settings
    return bytes(out)


def _vp8_header_flags(data):
    """(simple filter?, sharpness, token partitions) read from the first partition of a simple lossy file — to make sure the
    files below really carry the features they are meant to."""
    body = data[data.find(b"VP8 ") + 8:]
    d = body[10:]
    st = dict(p=2, v=(d[0] << 8) | d[1], r=255, bc=0)

    def bit(prob):
        split = 1 + (((st["r"] - 1) * prob) >> 8)
        big = split << 8
        if st["v"] >= big:
            r, st["r"], st["v"] = 1, st["r"] - split, st["v"] - big
        else:
            r, st["r"] = 0, split
        while st["r"] < 128:
            st["v"] <<= 1
            st["r"] <<= 1
            st["bc"] += 1
            if st["bc"] == 8:
                st["bc"] = 0
                st["v"] |= d[st["p"]] if st["p"] < len(d) else 0
                st["p"] += 1
        return r

    def lit(n):
        v = 0
        for _ in range(n):
            v = (v << 1) | bit(128)
        return v

    lit(2)
    if lit(1):  # segments
        update_map = lit(1)
        if lit(1):
            lit(1)
            for n in (7, 6):
                for _ in range(4):
                    if lit(1):
                        lit(n + 1)
        if update_map:
            for _ in range(3):
                if lit(1):
                    lit(8)
    simple, _level, sharp = lit(1), lit(6), lit(3)
    if lit(1) and lit(1):
        for _ in range(8):
            if lit(1):
                lit(7)
    return simple, sharp, 1 << lit(2)


SETTINGS = [dict(filter_type=0, filter_strength=60), dict(filter_type=0, filter_strength=20, filter_sharpness=3),
            dict(filter_type=0, filter_strength=100, filter_sharpness=7, quality=20.0), dict(filter_type=0, filter_strength=100, quality=5.0),
            dict(method=0, partitions=1), dict(method=1, partitions=2, quality=30.0), dict(method=2, partitions=3, quality=90.0),
            dict(method=0, partitions=3, filter_type=0, filter_strength=80), dict(filter_sharpness=5, filter_strength=80),
            dict(filter_sharpness=2, filter_strength=40, quality=10.0), dict(filter_sharpness=7, filter_strength=100, quality=40.0),
            dict(segments=1), dict(segments=2, sns_strength=100), dict(segments=3, sns_strength=80, quality=35.0), dict(autofilter=1, quality=40.0),
            dict(filter_strength=0), dict(preprocessing=1), dict(use_sharp_yuv=1), dict(quality=55.0, pass_=6, method=6, partition_limit=50)]


def test_lossy_features_pillow_cannot_write(tmp_path):
    """Simple loop filter, sharpness, 2 / 4 / 8 token partitions, 1-4 segments: files from libwebp's encoder driven through its
    C API, decoded by libwebp (Pillow) and by this library — equal pixels."""
    smooth, noise, photo = images()
    edges = np.zeros((120, 160, 3), np.uint8)
    edges[::7] = 255
    edges[:, ::5, 1] = 200
    edges[30:60, 40:100] = (255, 0, 0)
    if _libwebp_encode(noise[:16, :16]) is None:
        pytest.skip("Pillow's bundled libwebp not found (or another struct layout)")
    seen_simple, seen_sharp, seen_parts = False, set(), set()
    for img in (photo, edges, noise):
        for s in SETTINGS:
            data = _libwebp_encode(img, **s)
            simple, sharp, parts = _vp8_header_flags(data)
            seen_simple |= bool(simple)
            seen_sharp.add(sharp)
            seen_parts.add(parts)
            f = tmp_path / "x.webp"
            f.write_bytes(data)
            want = np.asarray(PIL.open(f).convert("RGB"))
            got = mosaic_b200.webp_decode(f)
            assert got.shape == want.shape and (got == want).all(), s
    assert seen_simple and {0, 2, 3, 5, 7} <= seen_sharp and {1, 2, 4, 8} <= seen_parts


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
