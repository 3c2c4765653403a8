"""PNG in / out of the host side (include/mosaic_png.hpp through the C ABI): what
`image::open(path)?.to_rgb8()` (main.rs:263-301) and `output_img.save(path)` (main.rs:668-672) do for PNG files.

CPU only.  The decoder is checked against files written by Pillow in every colour type / bit depth / interlace mode
(an independent encoder) and against hand-assembled files for what Pillow cannot write (16-bit RGB(A), 1/2/4-bit grey
with Adam7); the encoder is checked by decoding its files with Pillow (an independent decoder) and with itself."""
import io
import os
import struct
import zlib

import numpy as np
import pytest

import mosaic_b200

PIL = pytest.importorskip("PIL.Image")


def rgb_of_16(v):  # image 0.25: FromPrimitive<u16> for u8 = round(v * 255 / 65535)
    return ((v.astype(np.uint32) + 128) // 257).astype(np.uint8)


def hand_png(width, height, depth, ctype, rows, interlace=0, palette=None):
    """Assemble a PNG from raw (already packed, unfiltered) scanlines `rows` (bytes each); filter 0 everywhere."""
    def chunk(t, body):
        return struct.pack(">I", len(body)) + t + body + struct.pack(">I", zlib.crc32(t + body) & 0xFFFFFFFF)
    raw = b"".join(b"\x00" + r for r in rows)
    out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", width, height, depth, ctype, 0, 0, interlace))
    if palette is not None:
        out += chunk(b"PLTE", bytes(palette))
    z = zlib.compress(raw, 6)
    out += chunk(b"IDAT", z[:len(z) // 2]) + chunk(b"IDAT", z[len(z) // 2:])  # split on purpose
    return out + chunk(b"tEXt", b"k\x00v") + chunk(b"IEND", b"")


@pytest.fixture
def rng():
    return np.random.default_rng(20261018)


@pytest.mark.parametrize("mode", ["RGB", "RGBA", "L", "LA", "P", "1"])
@pytest.mark.parametrize("size", [(1, 1), (7, 5), (33, 17), (64, 64)])
def test_decode_matches_pillow(tmp_path, rng, mode, size):
    w, h = size
    if mode == "P":
        img = PIL.fromarray(rng.integers(0, 37, (h, w), dtype=np.uint8), "P")
        img.putpalette(rng.integers(0, 256, 37 * 3, dtype=np.uint8).tobytes())
    elif mode == "1":
        img = PIL.fromarray((rng.integers(0, 2, (h, w), dtype=np.uint8) * 255), "L").convert("1")
    else:
        ch = {"RGB": 3, "RGBA": 4, "L": 1, "LA": 2}[mode]
        arr = rng.integers(0, 256, (h, w, ch), dtype=np.uint8)
        img = PIL.fromarray(arr[:, :, 0] if ch == 1 else arr, mode)
    f = tmp_path / f"{mode}.png"
    img.save(f, optimize=bool(w % 2))
    got = mosaic_b200.png_decode(f)
    # to_rgb8: alpha is dropped (not blended), grey is replicated, the palette is expanded
    want = np.asarray(img.convert("RGBA").convert("RGB") if mode in ("RGBA", "LA") else img.convert("RGB"))
    if mode in ("RGBA", "LA"):
        want = np.asarray(img)[:, :, :3] if mode == "RGBA" else np.repeat(np.asarray(img)[:, :, :1], 3, axis=2)
    assert got.shape == (h, w, 3) and (got == want).all()


def test_decode_every_filter_type(tmp_path, rng):
    # Pillow's encoder picks filters adaptively on smooth content: a gradient + noise image exercises Sub/Up/Average/Paeth
    yy, xx = np.mgrid[0:96, 0:128]
    arr = np.stack([(xx * 2) % 256, (yy * 3 + xx) % 256, (xx * yy // 7) % 256], -1).astype(np.uint8)
    arr[::7, ::5] ^= rng.integers(0, 256, arr[::7, ::5].shape, dtype=np.uint8)
    f = tmp_path / "grad.png"
    PIL.fromarray(arr, "RGB").save(f)
    data = f.read_bytes()
    raw = zlib.decompress(b"".join(data[i + 8:i + 8 + struct.unpack(">I", data[i:i + 4])[0]]
                                   for i in range(8, len(data) - 12) if data[i + 4:i + 8] == b"IDAT"))
    used = {raw[r * (1 + 128 * 3)] for r in range(96)}
    assert len(used) >= 3, used  # several filter types really occur in the file
    assert (mosaic_b200.png_decode(f) == arr).all()


@pytest.mark.parametrize("ctype,depth", [(0, 1), (0, 2), (0, 4), (0, 8), (0, 16), (2, 8), (2, 16), (3, 1), (3, 2), (3, 4), (3, 8),
                                         (4, 8), (4, 16), (6, 8), (6, 16)])
@pytest.mark.parametrize("interlace", [0, 1])
def test_decode_hand_assembled(tmp_path, rng, ctype, depth, interlace):
    w, h = 13, 11
    ch = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]
    samples = rng.integers(0, 1 << depth if ctype != 3 else min(1 << depth, 5), (h, w, ch), dtype=np.uint32)
    palette = rng.integers(0, 256, (5, 3), dtype=np.uint8) if ctype == 3 else None

    def pack(block):  # (rows, cols, ch) samples -> packed scanlines
        lines = []
        for r in range(block.shape[0]):
            flat = block[r].reshape(-1)
            if depth == 16:
                lines.append(flat.astype(">u2").tobytes())
            elif depth == 8:
                lines.append(flat.astype(np.uint8).tobytes())
            else:
                bits = "".join(format(int(v), f"0{depth}b") for v in flat)
                bits += "0" * (-len(bits) % 8)
                lines.append(bytes(int(bits[i:i + 8], 2) for i in range(0, len(bits), 8)))
        return lines

    if interlace:
        rows = []
        for x0, y0, dx, dy in ((0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)):
            sub = samples[y0::dy, x0::dx]
            if sub.shape[0] and sub.shape[1]:
                rows += pack(sub)
    else:
        rows = pack(samples)
    f = tmp_path / "hand.png"
    f.write_bytes(hand_png(w, h, depth, ctype, rows, interlace, None if palette is None else palette.reshape(-1)))
    got = mosaic_b200.png_decode(f)
    if ctype == 3:
        want = palette[samples[:, :, 0]]
    else:
        s8 = rgb_of_16(samples) if depth == 16 else (samples * (255 // ((1 << depth) - 1))).astype(np.uint8)
        want = np.repeat(s8[:, :, :1], 3, axis=2) if ch <= 2 else s8[:, :, :3]
    assert (got == want).all()
    if depth == 8 or ctype == 3 or (ctype == 0 and depth < 16):  # Pillow agrees wherever it has the same semantics
        assert (np.asarray(PIL.open(f).convert("RGBA" if ch in (2, 4) else "RGB"))[:, :, :3] == want).all()


@pytest.mark.parametrize("size", [(1, 1), (5, 3), (200, 120), (257, 63)])
def test_encode_round_trips_and_pillow_reads_it(tmp_path, rng, size):
    w, h = size
    yy, xx = np.mgrid[0:h, 0:w]
    arr = np.stack([(xx * 3) % 256, (yy * 5) % 256, (xx + yy) % 256], -1).astype(np.uint8)
    arr[rng.random((h, w)) < 0.2] = rng.integers(0, 256, 3, dtype=np.uint8)
    f = tmp_path / "out.png"
    mosaic_b200.png_encode(f, arr)
    assert (np.asarray(PIL.open(f).convert("RGB")) == arr).all()
    with PIL.open(f) as im:
        assert im.mode == "RGB" and im.size == (w, h)
    assert (mosaic_b200.png_decode(f) == arr).all()
    assert os.path.getsize(f) < arr.nbytes + 1024  # (smooth content compresses)


def test_decode_errors_are_reported_not_thrown(tmp_path):
    bad = tmp_path / "bad.png"
    bad.write_bytes(b"not a png at all")
    with pytest.raises(mosaic_b200.MosaicError, match="signature"):
        mosaic_b200.png_decode(bad)
    with pytest.raises(mosaic_b200.MosaicError, match="cannot open"):
        mosaic_b200.png_decode(tmp_path / "missing.png")
    good = tmp_path / "good.png"
    PIL.fromarray(np.zeros((4, 4, 3), np.uint8), "RGB").save(good)
    data = bytearray(good.read_bytes())
    data[-20] ^= 0xFF  # inside the IDAT body: CRC mismatch
    (tmp_path / "crc.png").write_bytes(bytes(data))
    with pytest.raises(mosaic_b200.MosaicError, match="CRC"):
        mosaic_b200.png_decode(tmp_path / "crc.png")
    trunc = tmp_path / "trunc.png"
    trunc.write_bytes(good.read_bytes()[:-30])
    with pytest.raises(mosaic_b200.MosaicError):
        mosaic_b200.png_decode(trunc)


def test_header_larger_than_the_data_can_hold(tmp_path):
    """A header that claims 16M x 16M pixels in front of a 4x4 image's IDAT: refused before anything of the claimed size is
    allocated (deflate expands by at most ~1032 : 1, so the compressed stream bounds the image)."""
    import time
    rows = [bytes(12)] * 4
    data = bytearray(hand_png(4, 4, 8, 2, rows))
    ihdr = data.find(b"IHDR")
    data[ihdr + 4:ihdr + 12] = struct.pack(">II", 1 << 24, 1 << 24)
    data[ihdr + 17:ihdr + 21] = struct.pack(">I", zlib.crc32(bytes(data[ihdr:ihdr + 17])) & 0xFFFFFFFF)
    f = tmp_path / "huge.png"
    f.write_bytes(bytes(data))
    t0 = time.time()
    with pytest.raises(mosaic_b200.MosaicError, match="too short for the dimensions"):
        mosaic_b200.png_decode(f)
    assert time.time() - t0 < 2.0
