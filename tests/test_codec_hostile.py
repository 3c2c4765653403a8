"""Structure-aware hostile files (tests/hostile_files.py) through the host-side decoders under AddressSanitizer + UBSan
(tests/cpp/decode_files.cpp): every field of a JPEG / PNG / WebP container that sizes, indexes or selects something, at its
edge values.  CPU only.  (The byte-flip fuzzer is tests/test_codec_fuzz.py.)"""
import os
import re
import shutil
import subprocess
import sys

import numpy as np
import pytest

PIL = pytest.importorskip("PIL.Image")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _seeds():
    import io
    rng = np.random.default_rng(9)
    ph = rng.integers(0, 256, (61, 83, 3)).astype(np.float64)
    for _ in range(3):
        ph = (ph + np.roll(ph, 1, 0) + np.roll(ph, 1, 1)) / 3
    ph = ph.astype(np.uint8)
    im = PIL.fromarray(ph)
    g = im.convert("L")
    rgba = PIL.fromarray(np.dstack([ph, ph[:, :, 0]]))

    def enc(img, fmt, **kw):
        try:
            b = io.BytesIO()
            img.save(b, format=fmt, **kw)
            return b.getvalue()
        except (OSError, KeyError, ValueError):      # a Pillow without that format
            return None

    jpegs = [enc(im, "JPEG", quality=85), enc(im, "JPEG", quality=40, progressive=True, subsampling=2),
             enc(im, "JPEG", quality=95, subsampling=1, optimize=True), enc(g, "JPEG", quality=70),
             enc(g, "JPEG", quality=70, progressive=True), enc(im, "JPEG", quality=75, restart_marker_blocks=2)]
    pngs = [enc(im, "PNG"), enc(im.convert("P"), "PNG"), enc(rgba, "PNG"), enc(g, "PNG"), enc(g.convert("1"), "PNG"),
            enc(PIL.fromarray(ph[:, :, 0].astype(np.uint16) * 257), "PNG"), enc(g.convert("LA"), "PNG")]
    webps = [enc(im, "WEBP", quality=60), enc(im, "WEBP", lossless=True),
             enc(im.convert("P").convert("RGB"), "WEBP", lossless=True, quality=100, method=6), enc(rgba, "WEBP", quality=30),
             enc(rgba, "WEBP", lossless=True), enc(im, "WEBP", quality=5, method=0)]
    return [s for s in jpegs if s], [s for s in pngs if s], [s for s in webps if s]


def test_decoders_survive_structured_hostile_files(tmp_path):
    import hostile_files as hf
    if not shutil.which("g++"):
        pytest.skip("no g++")
    exe = tmp_path / "decode_files"
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "decode_files.cpp"), "-o", str(exe), "-lz"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and "asan" in (r.stderr or "").lower():
        pytest.skip("this g++ has no sanitizer runtime")
    assert r.returncode == 0, r.stderr[-2000:]
    jpegs, pngs, webps = _seeds()
    assert jpegs and pngs
    files = []
    for ext, seeds, gen in ((".jpg", jpegs, hf.hostile_jpegs), (".png", pngs, hf.hostile_pngs), (".webp", webps, hf.hostile_webps)):
        for s, seed in enumerate(seeds):
            files.append(tmp_path / f"seed{s}{ext}")
            files[-1].write_bytes(seed)
            for i, (tag, data) in enumerate(gen(seed)):
                f = tmp_path / f"{tag}_{s}_{i}{ext}"
                f.write_bytes(data)
                files.append(f)
    assert len(files) > 8000
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:allocator_may_return_null=0:max_allocation_size_mb=2048",
               UBSAN_OPTIONS="print_stacktrace=1")
    decoded = refused = 0
    for i in range(0, len(files), 1000):
        r = subprocess.run([str(exe)] + [str(f) for f in files[i:i + 1000]], capture_output=True, text=True, timeout=900, env=env)
        assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
        m = re.search(r"decode_files: (\d+) decoded, (\d+) refused", r.stdout)
        decoded += int(m.group(1))
        refused += int(m.group(2))
    assert decoded >= len(jpegs) + len(pngs) + len(webps) and refused > 4000, (decoded, refused)
