"""Mutation fuzz of the host-side decoders under AddressSanitizer + UBSan (tests/cpp/fuzz_codecs.cpp): material directories
are untrusted input (main.rs:166-261 walks whatever the user points it at).  CPU only; a few seconds."""
import os
import shutil
import subprocess

import numpy as np
import pytest

PIL = pytest.importorskip("PIL.Image")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_decoders_survive_mutated_files(tmp_path):
    if not shutil.which("g++"):
        pytest.skip("no g++")
    exe = tmp_path / "fuzz_codecs"
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "fuzz_codecs.cpp"), "-o", str(exe), "-lz"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and "asan" in (r.stderr or "").lower():
        pytest.skip("this g++ has no sanitizer runtime")
    assert r.returncode == 0, r.stderr[-2000:]
    rng = np.random.default_rng(2)
    ph = rng.integers(0, 256, (90, 130, 3)).astype(np.float64)
    for _ in range(3):
        ph = (ph + np.roll(ph, 1, 0) + np.roll(ph, 1, 1)) / 3
    ph = ph.astype(np.uint8)
    files = []

    def save(name, img, **kw):
        try:
            img.save(tmp_path / name, **kw)
            files.append(str(tmp_path / name))
        except (OSError, KeyError, ValueError):  # (a Pillow without that format)
            pass

    im = PIL.fromarray(ph)
    save("a.jpg", im, quality=80)
    save("b.jpg", im, quality=60, progressive=True, subsampling=2)
    save("c.jpg", im, quality=90, subsampling=1, optimize=True)
    save("a.png", im)
    save("b.png", im.convert("P"))
    save("c.png", PIL.fromarray(np.dstack([ph, ph[:, :, 0]])))
    save("a.webp", im, quality=70)
    save("b.webp", im, lossless=True)
    save("c.webp", im.convert("P").convert("RGB"), lossless=True, quality=100, method=6)
    save("d.webp", PIL.fromarray(np.dstack([ph, ph[:, :, 0]])), quality=50)
    assert len(files) >= 6
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:allocator_may_return_null=0", UBSAN_OPTIONS="print_stacktrace=1")
    r = subprocess.run([str(exe), "400"] + files, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    assert "fuzz_codecs:" in r.stdout


def test_random_valid_files_decode_like_pillow(tmp_path):
    """The other side of robustness: 700 random VALID files — sizes 1..200, noise / smooth / blocky / flat content, JPEG at every
    quality with random progressive / optimise / subsampling / restart settings, WebP lossy and lossless at random quality and
    effort with and without alpha, PNG in seven modes at every compression level — decode to exactly Pillow's pixels."""
    import mosaic_b200
    rng = np.random.default_rng(123)

    def img(h, w, kind):
        if kind == 0:
            a = rng.integers(0, 256, (h, w, 3))
        elif kind == 1:
            yy, xx = np.mgrid[0:h, 0:w]
            a = np.stack([127 + 120 * np.sin(xx / 7.0 + yy / 11.0), 127 + 120 * np.cos(xx / 5.0), (xx * 3 + yy * 5) % 256], -1)
        elif kind == 2:
            a = np.kron(rng.integers(0, 256, ((h + 7) // 8, (w + 7) // 8, 3)), np.ones((8, 8, 1)))[:h, :w]
        else:
            a = np.full((h, w, 3), rng.integers(0, 256))
        return np.clip(a, 0, 255).astype(np.uint8)

    def save(im, f, **kw):
        try:
            im.save(f, **kw)
            return True
        except (OSError, KeyError, ValueError):      # a Pillow without that format
            return False

    done = 0
    for trial in range(300):
        a = img(int(rng.integers(1, 200)), int(rng.integers(1, 200)), int(rng.integers(0, 4)))
        im = PIL.fromarray(a)
        if rng.integers(0, 5) == 0:
            im = im.convert("L")
        kw = dict(quality=int(rng.integers(1, 101)), progressive=bool(rng.integers(0, 2)), optimize=bool(rng.integers(0, 2)))
        if im.mode == "RGB":
            kw["subsampling"] = int(rng.integers(0, 3))
        if rng.integers(0, 3) == 0:
            kw["restart_marker_blocks"] = int(rng.integers(1, 9))
        f = tmp_path / "t.jpg"
        if save(im, f, **kw):
            assert (mosaic_b200.jpeg_decode(f) == np.asarray(PIL.open(f).convert("RGB"))).all(), (a.shape, im.mode, kw)
            done += 1
    for trial in range(200):
        h, w = int(rng.integers(1, 160)), int(rng.integers(1, 160))
        a = img(h, w, int(rng.integers(0, 4)))
        if rng.integers(0, 4) == 0:
            a = np.dstack([a, rng.integers(0, 256, (h, w, 1)).astype(np.uint8)])
        kw = dict(quality=int(rng.integers(0, 101)), method=int(rng.integers(0, 7)), lossless=bool(rng.integers(0, 2)))
        f = tmp_path / "t.webp"
        if save(PIL.fromarray(a), f, **kw):
            assert (mosaic_b200.webp_decode(f) == np.asarray(PIL.open(f).convert("RGB"))).all(), (a.shape, kw)
            done += 1
    for trial in range(200):
        h, w = int(rng.integers(1, 120)), int(rng.integers(1, 120))
        a = img(h, w, int(rng.integers(0, 4)))
        mode = ["RGB", "L", "P", "RGBA", "LA", "1", "I;16"][int(rng.integers(0, 7))]
        if mode == "I;16":
            im = PIL.fromarray(a[:, :, 0].astype(np.uint16) * int(rng.integers(1, 258)))
        elif mode == "RGBA":
            im = PIL.fromarray(np.dstack([a, a[:, :, :1]]))
        else:
            im = PIL.fromarray(a).convert(mode)
        f = tmp_path / "t.png"
        assert save(im, f, optimize=bool(rng.integers(0, 2)), compress_level=int(rng.integers(0, 10)))
        ref = PIL.open(f)
        if mode == "I;16":                            # image's to_rgb8(): (v + 128) / 257 (Pillow's own 16 -> 8 conversion differs)
            g = ((np.asarray(ref).astype(np.uint32) + 128) // 257).astype(np.uint8)
            want = np.dstack([g, g, g])
        else:
            want = np.asarray(ref.convert("RGB"))
        assert (mosaic_b200.png_decode(f) == want).all(), (h, w, mode)
        done += 1
    assert done >= 500


def test_random_images_through_the_encoders(tmp_path):
    """output_img.save(path) (main.rs:668-672) over 200 random images (1..180 pixels a side; noise, smooth, blocky, flat, four-level
    content): Pillow reads back exactly what the PNG and WebP writers were given, and the JPEG writer's file decodes to the pixels
    of the file libjpeg-turbo writes from the same image at the same quality (4:4:4)."""
    import mosaic_b200
    rng = np.random.default_rng(321)
    for trial in range(200):
        h, w, kind = int(rng.integers(1, 180)), int(rng.integers(1, 180)), int(rng.integers(0, 5))
        if kind == 0:
            a = rng.integers(0, 256, (h, w, 3))
        elif kind == 1:
            yy, xx = np.mgrid[0:h, 0:w]
            a = np.stack([127 + 120 * np.sin(xx / 7.0 + yy / 11.0), 127 + 120 * np.cos(xx / 5.0), (xx * 3 + yy * 5) % 256], -1)
        elif kind == 2:
            a = np.kron(rng.integers(0, 256, ((h + 7) // 8, (w + 7) // 8, 3)), np.ones((8, 8, 1)))[:h, :w]
        elif kind == 3:
            a = np.full((h, w, 3), rng.integers(0, 256))
        else:
            a = rng.integers(0, 4, (h, w, 3)) * 80
        a = np.clip(a, 0, 255).astype(np.uint8)
        mosaic_b200.png_encode(tmp_path / "e.png", a)
        assert (np.asarray(PIL.open(tmp_path / "e.png").convert("RGB")) == a).all(), (h, w, kind)
        try:
            mosaic_b200.webp_encode(tmp_path / "e.webp", a)
            im = PIL.open(tmp_path / "e.webp")
            assert im.format == "WEBP" and (np.asarray(im.convert("RGB")) == a).all(), (h, w, kind)
        except OSError:                               # a Pillow without WebP
            pass
        q = int(rng.integers(1, 101))
        mosaic_b200.jpeg_encode(tmp_path / "e.jpg", a, q)
        PIL.fromarray(a).save(tmp_path / "r.jpg", quality=q, subsampling=0)
        got, ref = (np.asarray(PIL.open(tmp_path / n).convert("RGB")) for n in ("e.jpg", "r.jpg"))
        assert got.shape == ref.shape and (got == ref).all(), (h, w, kind, q)
