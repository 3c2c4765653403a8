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
