"""JPEG in on the host side (include/mosaic_jpeg.hpp through the C ABI): what
`image::open(path)?.to_rgb8()` (main.rs:263-301) does for the format photo libraries are made of.

CPU only.  The decoder restates the arithmetic of libjpeg-turbo's defaults (islow inverse DCT, fancy upsampling, 16-bit
colour tables); Pillow decodes through libjpeg-turbo, so every pixel is expected to be EQUAL — baseline and progressive,
4:4:4 / 4:2:2 / 4:2:0 / 4:1:1, grey, RGB (Adobe transform 0), optimised Huffman tables, restart intervals, sizes that are
not multiples of the MCU.  (Against the reference's own decoder — image 0.25.6 / zune-jpeg — two conforming decoders
agree to about +-1 per channel: parity unpinned, see the header.)"""
import numpy as np
import pytest

import mosaic_b200

PIL = pytest.importorskip("PIL.Image")


def images():
    rng = np.random.default_rng(20261018)
    yy, xx = np.mgrid[0:70, 0:101]
    smooth = np.stack([(xx * 2) % 256, (yy * 3 + xx) % 256, (xx * yy // 7) % 256], -1).astype(np.uint8)
    noise = rng.integers(0, 256, (53, 67, 3), dtype=np.uint8)
    return smooth, noise


def roundtrip(tmp_path, arr, name, **kw):
    f = tmp_path / f"{name}.jpg"
    PIL.fromarray(arr).save(f, **kw)
    want = np.asarray(PIL.open(f).convert("RGB"))
    got = mosaic_b200.jpeg_decode(f)
    assert got.shape == want.shape, name
    assert (got == want).all(), f"{name}: {np.abs(got.astype(int) - want.astype(int)).max()} max difference to libjpeg-turbo"
    assert (mosaic_b200.image_decode(f) == want).all()


@pytest.mark.parametrize("progressive", [False, True])
@pytest.mark.parametrize("subsampling", [0, 1, 2, "4:1:1"])
def test_decode_equals_libjpeg_turbo(tmp_path, progressive, subsampling):
    smooth, noise = images()
    roundtrip(tmp_path, smooth, "smooth", quality=85, subsampling=subsampling, progressive=progressive)
    roundtrip(tmp_path, noise, "noise", quality=60, subsampling=subsampling, progressive=progressive)
    roundtrip(tmp_path, np.tile(smooth, (3, 3, 1))[:197, :283], "odd", quality=35, subsampling=subsampling, progressive=progressive,
              optimize=True)


@pytest.mark.parametrize("kw", [dict(quality=1), dict(quality=100, subsampling=0), dict(quality=90, optimize=True),
                                dict(quality=85, subsampling=2, restart_marker_blocks=3),
                                dict(quality=85, subsampling=1, restart_marker_rows=1, progressive=True),
                                dict(quality=85, keep_rgb=True), dict(quality=85, keep_rgb=True, progressive=True)])
def test_decode_variants(tmp_path, kw):
    smooth, noise = images()
    try:
        roundtrip(tmp_path, smooth, "a", **kw)
        roundtrip(tmp_path, noise, "b", **kw)
    except TypeError:
        pytest.skip("this Pillow cannot write that variant")


@pytest.mark.parametrize("shape", [(1, 1), (1, 9), (9, 1), (3, 16), (16, 3), (8, 8), (17, 33)])
@pytest.mark.parametrize("subsampling", [0, 2])
def test_decode_small_and_grey(tmp_path, shape, subsampling):
    smooth, _ = images()
    roundtrip(tmp_path, smooth[:shape[0], :shape[1]], "rgb", quality=80, subsampling=subsampling)
    roundtrip(tmp_path, np.ascontiguousarray(smooth[:shape[0], :shape[1], 1]), "grey", quality=80)
    roundtrip(tmp_path, np.ascontiguousarray(smooth[:shape[0], :shape[1], 1]), "greyp", quality=80, progressive=True)


def test_decode_errors_and_dispatch(tmp_path):
    _, noise = images()
    f = tmp_path / "cmyk.jpg"
    PIL.fromarray(noise).convert("CMYK").save(f)
    with pytest.raises(mosaic_b200.MosaicError, match="1- and 3-component"):
        mosaic_b200.jpeg_decode(f)
    bad = tmp_path / "bad.jpg"
    bad.write_bytes(b"\xff\xd8\xff\xe0 not a jpeg")
    with pytest.raises(mosaic_b200.MosaicError):
        mosaic_b200.jpeg_decode(bad)
    good = tmp_path / "good.jpg"
    PIL.fromarray(noise).save(good, quality=80)
    trunc = tmp_path / "trunc.jpg"
    trunc.write_bytes(good.read_bytes()[:200])
    with pytest.raises(mosaic_b200.MosaicError):
        mosaic_b200.jpeg_decode(trunc)
    # hostile headers: an over-subscribed Huffman table (more codes of a length than bit patterns) and dimensions no file of
    # this size can hold are refused before anything is built or allocated from them
    raw = bytearray(good.read_bytes())
    dht = raw.find(b"\xff\xc4")
    bad_dht = bytearray(raw)
    bad_dht[dht + 5] = 7  # seven codes of length 1
    (tmp_path / "dht.jpg").write_bytes(bytes(bad_dht))
    with pytest.raises(mosaic_b200.MosaicError, match="Huffman"):
        mosaic_b200.jpeg_decode(tmp_path / "dht.jpg")
    sof = raw.find(b"\xff\xc0")
    huge = bytearray(raw)
    huge[sof + 5:sof + 9] = b"\xff\xff\xff\xff"  # 65535 x 65535
    (tmp_path / "huge.jpg").write_bytes(bytes(huge))
    with pytest.raises(mosaic_b200.MosaicError, match="too short for the dimensions"):
        mosaic_b200.jpeg_decode(tmp_path / "huge.jpg")
    # image_decode tells the formats apart by content, whatever the extension says
    png = tmp_path / "really_a_png.jpg"
    mosaic_b200.png_encode(tmp_path / "x.png", noise)
    png.write_bytes((tmp_path / "x.png").read_bytes())
    assert (mosaic_b200.image_decode(png) == noise).all()
    other = tmp_path / "other.gif"
    other.write_bytes(b"GIF89a\x01\x00\x01\x00\x00\x00\x00;")
    with pytest.raises(mosaic_b200.MosaicError, match="not a PNG, JPEG or WebP"):
        mosaic_b200.image_decode(other)
    empty_webp = tmp_path / "empty.webp"  # a WebP container without an image chunk: handed to the WebP decoder, which refuses it
    empty_webp.write_bytes(b"RIFF\x04\x00\x00\x00WEBP")
    with pytest.raises(mosaic_b200.MosaicError, match="webp"):
        mosaic_b200.image_decode(empty_webp)


def test_cli_selects_jpeg_materials(tmp_path):
    """The host CLI's load_tiles (main.rs:166-261) over a directory of JPEG files: extension matched case-insensitively,
    dimensions from the frame header for the aspect filter.  (Without a GPU the run then stops with the library's loud
    error; with one it completes.)"""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "mosaic-art-rust_b200")
    subprocess.check_call(["make", "-C", pkg, "host/mosaic-b200"], stdout=subprocess.DEVNULL)
    cli = os.path.join(pkg, "host", "mosaic-b200")
    rng = np.random.default_rng(3)
    PIL.fromarray(rng.integers(0, 256, (60, 100, 3), dtype=np.uint8)).save(tmp_path / "target.JPG", quality=90)
    mdir = tmp_path / "m"
    mdir.mkdir()
    sizes = [(50, 30), (64, 64), (25, 15), (30, 70), (100, 60)]  # 3 of 5 match 5:3
    for i in range(10):
        w, h = sizes[i % 5]
        PIL.fromarray(rng.integers(0, 256, (h, w, 3), dtype=np.uint8)).save(
            mdir / f"f{i:02d}.{'jpeg' if i % 2 else 'jpg'}", quality=85, progressive=bool(i % 3 == 0))
    (mdir / "broken.jpg").write_bytes(b"\xff\xd8 nothing else")
    r = subprocess.run([cli, "-t", str(tmp_path / "target.JPG"), "-m", str(mdir), "-o", str(tmp_path / "o.png"), "--grid-w", "5",
                        "--grid-h", "3", "--similarity-db", str(tmp_path / "db.json")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 or "no usable CUDA device" in r.stderr, r.stderr[-1500:]
    assert "Target aspect ratio: 1.667" in r.stdout
    assert "Loaded 6 tiles" in r.stdout          # 2 rounds x 3 matching sizes
    assert "broken.jpg" in r.stderr
    if r.returncode == 0:
        out = mosaic_b200.png_decode(tmp_path / "o.png")
        assert out.shape == (60, 100, 3)


@pytest.mark.parametrize("quality", [30, 75, 95])
@pytest.mark.parametrize("shape", [(1, 1), (5, 3), (70, 101), (120, 200)])
def test_encode_is_what_libjpeg_writes(tmp_path, quality, shape):
    """output_img.save("x.jpg") (main.rs:668-672): baseline 4:4:4 with the IJG tables and arithmetic — the file decodes to
    exactly the pixels of the file libjpeg-turbo (Pillow) writes from the same image at the same quality, and is as large."""
    rng = np.random.default_rng(7)
    yy, xx = np.mgrid[0:shape[0], 0:shape[1]]
    arr = (np.stack([127 + 100 * np.sin(xx / 17.0) * np.cos(yy / 11.0), 127 + 90 * np.sin((xx + yy) / 23.0), 127 + 80 * np.cos(xx / 9.0)], -1)
           + rng.normal(0, 4, shape + (3,))).clip(0, 255).astype(np.uint8)
    ours, theirs = tmp_path / "ours.jpg", tmp_path / "theirs.jpg"
    mosaic_b200.jpeg_encode(ours, arr, quality)
    PIL.fromarray(arr).save(theirs, quality=quality, subsampling=0)
    dec = np.asarray(PIL.open(ours).convert("RGB"))
    assert (dec == np.asarray(PIL.open(theirs).convert("RGB"))).all()
    assert (mosaic_b200.jpeg_decode(ours) == dec).all()
    assert abs(ours.stat().st_size - theirs.stat().st_size) <= 64   # (headers may differ by a comment or a density field)
    with PIL.open(ours) as im:
        assert im.size == (shape[1], shape[0]) and im.mode == "RGB"


def test_image_open_goes_by_the_extension_like_image_open(tmp_path):
    """`image::open`: "the image's format is determined from the path's file extension" — no content sniffing.  A JPEG file
    named .png fails in the PNG decoder (the reference then drops the tile with a message, main.rs:198-206)."""
    rng = np.random.default_rng(4)
    img = rng.integers(0, 256, (21, 34, 3), dtype=np.uint8)
    mosaic_b200.png_encode(tmp_path / "a.png", img)
    mosaic_b200.jpeg_encode(tmp_path / "a.jpg", img)
    mosaic_b200.webp_encode(tmp_path / "a.webp", img)
    assert (mosaic_b200.image_open(tmp_path / "a.png") == img).all()
    assert (mosaic_b200.image_open(tmp_path / "a.jpg") == mosaic_b200.jpeg_decode(tmp_path / "a.jpg")).all()
    assert (mosaic_b200.image_open(tmp_path / "a.webp") == img).all()
    for name, src in (("B.PNG", "a.png"), ("c.JpEg", "a.jpg"), ("d.tar.WEBP", "a.webp")):     # lower-cased, last extension
        (tmp_path / name).write_bytes((tmp_path / src).read_bytes())
        assert mosaic_b200.image_open(tmp_path / name).shape == img.shape
    for name, src in (("jpeg_inside.png", "a.jpg"), ("png_inside.jpg", "a.png"), ("png_inside.webp", "a.png"),
                      ("no_extension", "a.png"), (".png", "a.png"), ("a.gif", "a.png"), ("a.png.bak", "a.png")):
        (tmp_path / name).write_bytes((tmp_path / src).read_bytes())
        with pytest.raises(mosaic_b200.MosaicError):
            mosaic_b200.image_open(tmp_path / name)
    d = tmp_path / "dir.png"
    d.mkdir()
    (d / "plain").write_bytes((tmp_path / "a.png").read_bytes())     # the dot belongs to the directory, not to the file
    with pytest.raises(mosaic_b200.MosaicError, match="extension"):
        mosaic_b200.image_open(d / "plain")
    assert (mosaic_b200.image_decode(tmp_path / "jpeg_inside.png") == mosaic_b200.jpeg_decode(tmp_path / "a.jpg")).all()  # the sniffing call
