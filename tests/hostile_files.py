"""Structure-aware hostile image files for the host-side decoders (tests/test_codec_hostile.py): instead of flipping bytes
(tests/cpp/fuzz_codecs.cpp) these walk the container structure of a valid file and put every field that sizes, indexes or
selects something to its edge values — JPEG marker segments (SOF precision / dimensions / sampling factors / table ids, DQT and
DHT ids and code counts, SOS component lists and spectral / approximation parameters of every scan, restart intervals,
segment lengths), PNG chunks (IHDR dimensions and every bit-depth x colour-type x interlace combination, PLTE / tRNS lengths,
filter bytes, short / long / missing / garbage IDAT streams, chunk order and length fields) and WebP (RIFF and chunk sizes,
VP8X flags and canvas, VP8 frame tags / start codes / dimensions / first-partition bytes, VP8L header and bitstream bytes,
ALPH headers).  Each generator yields (tag, bytes)."""
import itertools
import struct
import zlib


# ------------------------------------------------------------------------------------------------------------ JPEG
def _jpeg_segments(data):
    segs, i = [], 2
    while i + 4 <= len(data) and data[i] == 0xFF:
        m = data[i + 1]
        ln = struct.unpack(">H", data[i + 2:i + 4])[0]
        segs.append((m, i, i + 4, ln - 2))
        if m == 0xDA:
            break
        i += 2 + ln
    return segs


def hostile_jpegs(data):
    def put(off, vals):
        b = bytearray(data)
        b[off:off + len(vals)] = bytes(vals)
        return bytes(b)

    for m, ms, ps, pl in _jpeg_segments(data):
        if m in (0xC0, 0xC1, 0xC2):
            for off, vals in ((0, (0, 7, 12, 16, 255)), (1, (0, 255)), (2, (0, 1, 255)), (3, (0, 255)), (4, (0, 1, 255)),
                              (5, (0, 1, 2, 4, 5, 255))):
                for v in vals:
                    yield f"sof{off}", put(ps + off, [v])
            nf = data[ps + 5]
            for c in range(nf):
                for v in (0x00, 0x01, 0x10, 0x44, 0x55, 0xFF, 0x41, 0x14, 0x33):
                    yield "hv", put(ps + 6 + 3 * c + 1, [v])
                for v in (3, 4, 255):
                    yield "tq", put(ps + 6 + 3 * c + 2, [v])
                for v in (0, 255, data[ps + 6]):
                    yield "cid", put(ps + 6 + 3 * c, [v])
            yield "sofdup", data[:ms] + data[ms:ps + pl] + data[ms:]
            yield "nosof", data[:ms] + data[ps + pl:]
        if m == 0xDB:
            for v in (0x04, 0x0F, 0x10, 0x13, 0x20, 0xFF):
                yield "dqt", put(ps, [v])
            yield "nodqt", data[:ms] + data[ps + pl:]
        if m == 0xC4:
            for v in (0x04, 0x0F, 0x14, 0x20, 0xFF):
                yield "dht", put(ps, [v])
            for k in range(16):
                for v in (255, 200, 0):
                    yield "dhtc", put(ps + 1 + k, [v])
            yield "nodht", data[:ms] + data[ps + pl:]
            for k in range(17, min(pl, 60)):
                yield "dhts", put(ps + k, [0xFF])
        if m == 0xDD:
            for v in (0, 1, 65535):
                yield "dri", put(ps, struct.pack(">H", v))
        if m == 0xDA:
            ns = data[ps]
            for v in (0, 5, 255):
                yield "sosn", put(ps, [v])
            for c in range(ns):
                for v in (0, 99, 255):
                    yield "soscid", put(ps + 1 + 2 * c, [v])
                for v in (0x44, 0x33, 0xFF, 0x40, 0x04):
                    yield "sostab", put(ps + 2 + 2 * c, [v])
            base = ps + 1 + 2 * ns
            for ss, se, ah in itertools.product((0, 1, 63, 64, 255), (0, 5, 63, 64, 255), (0x00, 0x10, 0x01, 0xDD, 0xFF, 0x0E, 0xE0)):
                yield "sosp", put(base, [ss, se, ah])
        for v in (0, 1, 2, 3, pl + 1, pl + 3, 65535):
            yield "len", put(ms + 2, struct.pack(">H", v))
    i = 0
    while True:                                        # every scan of a progressive file
        i = data.find(b"\xff\xda", i + 1)
        if i < 0 or i + 5 > len(data):
            break
        ns = data[i + 4]
        base = i + 5 + 2 * ns
        if base + 3 > len(data):
            break
        for ss, se, ah in itertools.product((0, 1, 6, 63, 64), (0, 5, 63, 64, 200), (0x00, 0x10, 0x01, 0x21, 0xFF, 0x0F)):
            yield "scan", put(base, [ss, se, ah])
        for v in (0, 4, 5):
            yield "scann", put(i + 4, [v])


# ------------------------------------------------------------------------------------------------------------ PNG
def _png_chunks(data):
    i, res = 8, []
    while i + 12 <= len(data):
        ln = struct.unpack(">I", data[i:i + 4])[0]
        res.append((data[i + 4:i + 8], i, ln))
        i += 12 + ln
    return res


def _png_build(chs):
    b = b"\x89PNG\r\n\x1a\n"
    for t, payload in chs:
        b += struct.pack(">I", len(payload)) + t + payload + struct.pack(">I", zlib.crc32(t + payload) & 0xFFFFFFFF)
    return b


def hostile_pngs(data):
    chs = [(t, data[i + 8:i + 8 + ln]) for t, i, ln in _png_chunks(data)]
    ihdr = chs[0][1]
    w, h, bd, ct, cm, fm, im = struct.unpack(">IIBBBBB", ihdr)

    def with_ihdr(**kw):
        f = dict(w=w, h=h, bd=bd, ct=ct, cm=cm, fm=fm, im=im)
        f.update(kw)
        return _png_build([(b"IHDR", struct.pack(">IIBBBBB", f["w"], f["h"], f["bd"], f["ct"], f["cm"], f["fm"], f["im"]))] + chs[1:])

    for nw in (0, 1, w - 1, w + 1, w * 2, 2 ** 31 - 1, 2 ** 31, 2 ** 32 - 1, 65536):
        yield "w", with_ihdr(w=nw)
    for nh in (0, 1, h - 1, h + 1, h * 2, 2 ** 31 - 1, 2 ** 32 - 1):
        yield "h", with_ihdr(h=nh)
    for nbd, nct in itertools.product((0, 1, 2, 3, 4, 8, 16, 32, 255), (0, 1, 2, 3, 4, 5, 6, 7, 255)):
        yield "bdct", with_ihdr(bd=nbd, ct=nct)
    for v in (1, 255):
        yield "cm", with_ihdr(cm=v)
        yield "fm", with_ihdr(fm=v)
    for v in (1, 2, 255):
        yield "im", with_ihdr(im=v)
    for nbd, nct, nim in itertools.product((1, 2, 4, 8, 16), (0, 2, 3, 4, 6), (0, 1)):
        yield "combo", with_ihdr(bd=nbd, ct=nct, im=nim)
        for nw2, nh2 in ((1, 1), (3, 5), (9, 2)):
            yield "combo_small", with_ihdr(bd=nbd, ct=nct, im=nim, w=nw2, h=nh2)
    idat = b"".join(p for t, p in chs if t == b"IDAT")
    raw = zlib.decompress(idat)
    others = [(t, p) for t, p in chs if t not in (b"IDAT", b"IEND")]

    def with_raw(r):
        return _png_build(others + [(b"IDAT", zlib.compress(bytes(r))), (b"IEND", b"")])

    yield "short", with_raw(raw[:len(raw) // 2])
    yield "long", with_raw(raw + raw)
    yield "empty", with_raw(b"")
    yield "noidat", _png_build(others + [(b"IEND", b"")])
    yield "idat0", _png_build(others + [(b"IDAT", b""), (b"IEND", b"")])
    yield "idatcut", _png_build(others + [(b"IDAT", idat[:10]), (b"IEND", b"")])
    yield "idatgarbage", _png_build(others + [(b"IDAT", b"\x78\x9c" + bytes(range(64))), (b"IEND", b"")])
    if im == 0:
        stride = len(raw) // h
        for ft in (5, 6, 255, 4, 3, 2, 1):
            r = bytearray(raw)
            for y in range(h):
                r[y * stride] = ft
            yield "filter", with_raw(r)
    for plen in (0, 1, 2, 3, 4, 767, 768, 769, 771, 1024):
        c2 = [(t, p) for t, p in others if t != b"PLTE"]
        c2.insert(1, (b"PLTE", (bytes(range(256)) * 5)[:plen]))
        yield "plte", _png_build(c2 + [(b"IDAT", idat), (b"IEND", b"")])
    for tlen in (0, 1, 2, 6, 7, 256, 257, 1000):
        yield "trns", _png_build(others + [(b"tRNS", bytes(tlen)), (b"IDAT", idat), (b"IEND", b"")])
    yield "ihdr2", _png_build([chs[0], chs[0]] + chs[1:])
    yield "noihdr", _png_build(chs[1:])
    yield "noiend", _png_build(chs[:-1])
    yield "idatfirst", _png_build([(b"IDAT", idat)] + chs)
    yield "ihdrshort", _png_build([(b"IHDR", ihdr[:12])] + chs[1:])
    yield "ihdrlong", _png_build([(b"IHDR", ihdr + b"\0\0")] + chs[1:])
    for t, i, ln in _png_chunks(data):
        for v in (0, ln - 1 if ln else 0, ln + 1, ln + 1000, 2 ** 31 - 1, 2 ** 31, 2 ** 32 - 1):
            b = bytearray(data)
            b[i:i + 4] = struct.pack(">I", v)
            yield "len", bytes(b)


# ------------------------------------------------------------------------------------------------------------ WebP
def _riff_chunks(data):
    i, res = 12, []
    while i + 8 <= len(data):
        ln = struct.unpack("<I", data[i + 4:i + 8])[0]
        res.append((data[i:i + 4], i, ln))
        i += 8 + ln + (ln & 1)
    return res


def _riff_build(chs):
    body = b"WEBP"
    for t, p in chs:
        body += t + struct.pack("<I", len(p)) + p + (b"\0" if len(p) & 1 else b"")
    return b"RIFF" + struct.pack("<I", len(body)) + body


def hostile_webps(data):
    chs = [(t, data[i + 8:i + 8 + ln]) for t, i, ln in _riff_chunks(data)]

    def swap(tag, payload):
        return _riff_build([(t, bytes(payload)) if t == tag else (t, p) for t, p in chs])

    for v in (0, 4, 11, len(data) - 9, len(data), len(data) * 2, 2 ** 31, 2 ** 32 - 1):
        b = bytearray(data)
        b[4:8] = struct.pack("<I", v)
        yield "riffsz", bytes(b)
    yield "fourcc", data[:8] + b"WEBQ" + data[12:]
    for t, i, ln in _riff_chunks(data):
        for v in (0, 1, ln - 1 if ln else 0, ln + 1, ln + 2, ln * 2, 2 ** 31 - 1, 2 ** 31, 2 ** 32 - 1, 2 ** 32 - 2):
            b = bytearray(data)
            b[i + 4:i + 8] = struct.pack("<I", v)
            yield "chsz", bytes(b)
    yield "dup", _riff_build(chs + chs)
    yield "rev", _riff_build(list(reversed(chs)))
    yield "noimage", _riff_build([c for c in chs if c[0] not in (b"VP8 ", b"VP8L")])
    yield "vp8x0", _riff_build([(b"VP8X", bytes(10))] + [c for c in chs if c[0] != b"VP8X"])
    for flags in (0x02, 0x10, 0x20, 0x08, 0x04, 0xFF):
        for cw, ch in ((0, 0), (1, 1), (0xFFFFFF, 0xFFFFFF), (82, 60), (83, 61), (84, 62)):
            vp8x = bytes([flags, 0, 0, 0]) + struct.pack("<I", cw)[:3] + struct.pack("<I", ch)[:3]
            yield "vp8x", _riff_build([(b"VP8X", vp8x)] + [c for c in chs if c[0] != b"VP8X"])
    yield "vp8xshort", _riff_build([(b"VP8X", bytes(3))] + chs)
    yield "alphfirst", _riff_build([(b"ALPH", bytes(range(50)))] + chs)
    yield "anim", _riff_build([(b"ANIM", bytes(6)), (b"ANMF", bytes(16) + chs[-1][1])])
    for t, p in chs:
        if t == b"VP8 ":
            tag = p[0] | (p[1] << 8) | (p[2] << 16)
            for kf, ver, show, psize in itertools.product((0, 1), (0, 1, 2, 3, 4, 7), (0, 1),
                                                          (0, 1, len(p), len(p) * 4, (1 << 19) - 1, tag >> 5)):
                nt = kf | (ver << 1) | (show << 4) | ((psize & 0x7FFFF) << 5)
                q = bytearray(p)
                q[0], q[1], q[2] = nt & 255, (nt >> 8) & 255, (nt >> 16) & 255
                yield "vp8tag", swap(t, q)
            for sc in (b"\x9d\x01\x2b", b"\0\0\0"):
                q = bytearray(p)
                q[3:6] = sc
                yield "vp8sc", swap(t, q)
            for wv, hv in itertools.product((0, 1, 83, 84, 0x3FFF, 0x4000 | 83, 0xC000 | 83, 0xFFFF),
                                            (0, 1, 61, 62, 0x3FFF, 0x8000 | 61, 0xFFFF)):
                q = bytearray(p)
                q[6:8] = struct.pack("<H", wv)
                q[8:10] = struct.pack("<H", hv)
                yield "vp8wh", swap(t, q)
            for cut in (3, 6, 9, 10, 11, 12, 20, 40, len(p) // 2, len(p) - 1):
                yield "vp8cut", swap(t, p[:cut])
            for off in range(10, 34):
                for v in (0, 0xFF, 0x55, 0xAA, 0x80, 0x01):
                    q = bytearray(p)
                    if off < len(q):
                        q[off] = v
                    yield "vp8hdr", swap(t, q)
        if t == b"VP8L":
            for sig in (0, 0x2E, 0xFF):
                q = bytearray(p)
                q[0] = sig
                yield "vp8lsig", swap(t, q)
            for wv, hv, al, ver in itertools.product((0, 82, 0x3FFF), (0, 60, 0x3FFF), (0, 1), (0, 1, 7)):
                q = bytearray(p)
                q[1:5] = struct.pack("<I", wv | (hv << 14) | (al << 28) | (ver << 29))
                yield "vp8lhdr", swap(t, q)
            for off in range(5, 45):
                for v in (0, 0xFF, 0x55, 0xAA, 0x0F, 0xF0, 0x01, 0x80):
                    q = bytearray(p)
                    if off < len(q):
                        q[off] = v
                    yield "vp8lbits", swap(t, q)
            for cut in (1, 4, 5, 6, 8, 16, len(p) // 2, len(p) - 1):
                yield "vp8lcut", swap(t, p[:cut])
        if t == b"ALPH":
            for hb in range(64):
                q = bytearray(p)
                q[0] = hb
                yield "alph", swap(t, q)
