"""The similarity-database reader against damaged files.

1. Mutation fuzz under AddressSanitizer + UBSan (tests/cpp/fuzz_simdb.cpp): `similarity_db.json` lies on disk between runs
   and main.rs:120-143 loads whatever is there — a damaged file must be an error code, never a crash.
2. What is ACCEPTED: `load_or_new` (similarity.rs:126-139) rebuilds the database from scratch on any error, so the set of
   files that load is behaviour.  A model of serde_json's rules for the struct (built on Python's json scanner with hooks
   that keep the number literals) decides accept / reject for ~6,000 mutated files and the reader has to agree, values
   included.

CPU only."""
import json
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

import mosaic_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "mosaic-art-rust_b200", "csrc")


def test_reader_survives_mutated_files(tmp_path):
    if not shutil.which("nvcc"):
        pytest.skip("no nvcc (simdb.cu includes the library's common header)")
    exe = tmp_path / "fuzz_simdb"
    san = "-fsanitize=address,-fsanitize=undefined,-fno-sanitize-recover=undefined"
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O1", "-g", "-Xcompiler", san,
           "-I", os.path.join(ROOT, "include"), "-I", CSRC, os.path.join(CSRC, "simdb.cu"),
           os.path.join(ROOT, "tests", "cpp", "fuzz_simdb.cpp"), "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and "asan" in (r.stderr or "").lower():
        pytest.skip("this host compiler has no sanitizer runtime")
    assert r.returncode == 0, r.stderr[-2000:]
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:allocator_may_return_null=0:max_allocation_size_mb=1024",
               UBSAN_OPTIONS="print_stacktrace=1")
    r = subprocess.run([str(exe), "1200", str(tmp_path / "scratch.json")], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    m = re.search(r"fuzz_simdb: (\d+) loaded, (\d+) refused", r.stdout)
    assert m and int(m.group(1)) > 100 and int(m.group(2)) > 1000, r.stdout


# ------------------------------------------------------------------------------------------------------------------
# a model of `serde_json::from_str::<SimilarityDatabase>` (serde_json 1.0.140 + serde derive, [3P] restated)
class _Num:
    def __init__(self, text, is_int):
        self.text, self.is_int = text, is_int


class _Reject(Exception):
    pass


def _no_constants(name):
    raise _Reject(f"{name} is not JSON")


def _f32(v):
    if not isinstance(v, _Num):
        raise _Reject("f32 wants a number")
    if v.is_int:
        i = int(v.text)
        if v.text != "-0" and -2 ** 63 <= i < 2 ** 64:
            if abs(i) > 2 ** 53:
                return None                          # integer -> f32 in one rounding: not modelled, value not compared
            return np.float32(i)
    d = float(v.text)
    if d in (float("inf"), float("-inf")):
        raise _Reject("number out of range")
    with np.errstate(over="ignore"):
        return np.float32(d)                         # Rust's `as f32`: beyond the range +-inf


def _usize(v):
    if not isinstance(v, _Num) or not v.is_int or v.text.startswith("-") or int(v.text) >= 2 ** 64:
        raise _Reject("usize wants a plain non-negative integer")
    return int(v.text)


def _string(v):
    if not isinstance(v, str):
        raise _Reject("string expected")
    try:
        v.encode("utf-8")                            # a lone surrogate from a \\u escape
    except UnicodeEncodeError:
        raise _Reject("lone surrogate")
    return v


def _struct(v, names):
    """field name -> value; a struct is an object (unknown fields skipped, duplicates refused) or a sequence"""
    if isinstance(v, list) and not isinstance(v, _Pairs):
        if len(v) != len(names):
            raise _Reject("invalid length")
        return dict(zip(names, v))
    if not isinstance(v, _Pairs):
        raise _Reject("struct expected")
    out = {}
    for k, val in v:
        _string(k)
        if k in names:
            if k in out:
                raise _Reject("duplicate field")
            out[k] = val
    if len(out) != len(names):
        raise _Reject("missing field")
    return out


class _Pairs(list):
    pass


def _model(data: bytes):
    """(lab f32 list with None for unmodelled entries, sim, {index: path} for rows, n path_to_index) or raises _Reject"""
    try:
        text = data.decode("utf-8")
    except UnicodeDecodeError:
        raise _Reject("not UTF-8")
    try:
        doc = json.loads(text, parse_float=lambda s: _Num(s, False), parse_int=lambda s: _Num(s, True),
                         parse_constant=_no_constants, object_pairs_hook=_Pairs)
    except (json.JSONDecodeError, RecursionError) as e:
        raise _Reject(str(e))
    top = _struct(doc, ["path_to_index", "index_to_path", "lab_colors", "similarities"])
    p2i, i2p = {}, {}
    if not isinstance(top["path_to_index"], _Pairs):
        raise _Reject("map expected")
    for k, v in top["path_to_index"]:
        p2i[_string(k)] = _usize(v)
    if not isinstance(top["index_to_path"], _Pairs):
        raise _Reject("map expected")
    for k, v in top["index_to_path"]:
        _string(k)
        if not re.fullmatch(r"0|[1-9][0-9]*", k) or int(k) >= 2 ** 64:
            raise _Reject("numeric key expected")
        i2p[int(k)] = _string(v)
    if not isinstance(top["lab_colors"], list) or isinstance(top["lab_colors"], _Pairs):
        raise _Reject("sequence expected")
    lab = []
    for c in top["lab_colors"]:
        f = _struct(c, ["l", "a", "b"])
        lab += [_f32(f["l"]), _f32(f["a"]), _f32(f["b"])]
    if not isinstance(top["similarities"], list) or isinstance(top["similarities"], _Pairs):
        raise _Reject("sequence expected")
    sim = [_f32(v) for v in top["similarities"]]
    return lab, sim, i2p, len(p2i)


def _seed_files(tmp_path):
    rng = np.random.default_rng(11)
    seeds = []
    for n in (0, 1, 4, 7):
        paths = [f"materials/img_{i}.png" for i in range(n)]
        if n >= 4:
            paths[1] = 'dir "q"\\b/\t\x01.jpg'
            paths[2] = "写真/🖼.webp"
        lab = np.stack([rng.uniform(0, 100, n), rng.uniform(-100, 100, n), rng.uniform(-100, 100, n)], 1).astype(np.float32)
        sim = rng.uniform(0, 250, n * (n - 1) // 2 if n > 1 else 0).astype(np.float32)
        if n >= 4:
            lab[1] = (1e-7, -0.0, 3.4e38)
            sim[:3] = (1.17549435e-38, 16777216.0, 0.1)
        f = tmp_path / f"seed{n}.json"
        mosaic_b200.SimilarityDb.save(str(f), paths, lab, sim)
        seeds.append(f.read_bytes())
    # hand-written: sequence forms, unknown fields, integers and exponents where floats are expected
    seeds.append(b'[{"a.png": 0, "b.png": 1}, {"0": "a.png", "1": "b.png"}, [[1, 2.5, -3], {"l": 1e1, "b": 7, "extra": [1, {"x": null}],'
                 b' "a": -0}], [12]]')
    seeds.append(b'{"version": {"major": 2, "tags": ["x", true, false, null, 1e999]}, "similarities": [0.5, 1E+2, 1e-60, 123456789012],'
                 b' "lab_colors": [{"l": 50, "a": 0.0, "b": 1}, {"l": 0, "a": 0, "b": 0}],'
                 b' "index_to_path": {"1": "y", "0": "x", "18446744073709551615": "far"}, "path_to_index": {"x": 0, "y": 1, "x": 7}}')
    return seeds


_TOKENS = [b'"', b"\\", b"\\u0041", b"\\ud800", b"\\udc00", b"\\ud83d\\uddbc", b"{", b"}", b"[", b"]", b",", b":", b"-", b"+", b".",
           b"e", b"E+", b"e-999", b"1e999", b"0", b"00", b"9", b"99999999999999999999", b"18446744073709551615",
           b"18446744073709551616", b"-1", b"-0", b"1.0", b"null", b"true", b"NaN", b"Infinity", b"inf", b"0x10", b" ", b"\n",
           b"\t", b"\x0b", b"\x00", b"\x1f", b"\xff", b"\xc3\xa9", b"\xed\xa0\x80", b"\xc0\xaf", b'"l"', b'"path_to_index"',
           b'"similarities": []', b'"index_to_path"', b'"lab_colors"', b'"other": {"a": [1, 2, {"b": "c"}]},']


def _mutate(rng, seed: bytes) -> bytes:
    bad = bytearray(seed)
    for _ in range(1 + int(rng.integers(0, 3))):
        if len(bad) < 3:
            break
        pos = int(rng.integers(0, len(bad)))
        kind = int(rng.integers(0, 6))
        if kind == 0:
            bad[pos] ^= 1 << int(rng.integers(0, 8))
        elif kind == 1:
            bad[pos] = int(rng.integers(0, 256))
        elif kind == 2:
            tok = _TOKENS[int(rng.integers(0, len(_TOKENS)))]
            if rng.integers(0, 2):
                bad[pos:pos + 1] = tok
            else:
                bad[pos:pos] = tok
        elif kind == 3:
            del bad[pos:pos + 1 + int(rng.integers(0, 12))]
        elif kind == 4:
            src = int(rng.integers(0, len(bad)))
            bad[pos:pos] = bad[src:src + 1 + int(rng.integers(0, 64))]
        else:
            d = pos
            while d < len(bad) and not (48 <= bad[d] <= 57):
                d += 1
            if d < len(bad):
                bad[d:d] = bytes([49 + int(rng.integers(0, 9))]) * (1 + int(rng.integers(0, 18)))
    if rng.integers(0, 12) == 0 and len(bad) > 4:
        del bad[1 + int(rng.integers(0, len(bad) - 1)):]
    return bytes(bad)


def _bits(values):
    return np.array(values, np.float32).view(np.uint32)


def test_reader_accepts_what_serde_json_accepts(tmp_path):
    seeds = _seed_files(tmp_path)
    rng = np.random.default_rng(2026)
    f = tmp_path / "mutated.json"
    accepted = refused = 0
    cases = [(s, s) for s in seeds]                  # the seeds themselves first
    for s in seeds:
        cases += [(s, _mutate(rng, s)) for _ in range(1000)]
    for seed, data in cases:
        f.write_bytes(data)
        try:
            want = _model(data)
        except _Reject as why:
            want = None
            reason = str(why)
        try:
            db = mosaic_b200.SimilarityDb.load(str(f))
        except mosaic_b200.MosaicError as e:
            assert want is None, f"the reader refuses a file serde_json loads ({e}): {data[:400]!r}"
            refused += 1
            continue
        assert want is not None, f"the reader loads a file serde_json refuses ({reason}): {data[:400]!r}"
        accepted += 1
        lab, sim, i2p, n_paths = want
        info = db.info()
        assert info == dict(rows=len(lab) // 3, similarities=len(sim), paths=n_paths), data[:400]
        glab, gsim = db.arrays()
        for got, exp in ((glab.ravel(), lab), (gsim.ravel(), sim)):
            keep = np.array([v is not None for v in exp], bool)
            assert (got.view(np.uint32)[keep] == _bits([v for v in exp if v is not None])).all(), data[:400]
        for i in range(info["rows"]):
            assert db.path(i) == i2p.get(i, ""), (i, data[:400])
        assert db.path(info["rows"]) is None and db.path(-1) is None
        db.close()
    assert accepted > 300 and refused > 2000, (accepted, refused)


def test_reader_follows_serde_in_the_named_cases(tmp_path):
    """The rules one by one (each line: file, loads?)."""
    ok = '"path_to_index": {}, "index_to_path": {}, "lab_colors": [], "similarities": []'
    cases = [
        ("{" + ok + "}", True),
        (" \n\t\r{" + ok + "}\n", True),
        ("﻿{" + ok + "}", False),                                        # a BOM is not whitespace
        ("{" + ok + "} x", False),                                            # trailing characters
        ("{" + ok + ', "future_field": {"a": [1, 2, {"b": null}]}}', True),   # unknown fields are skipped
        ("{" + ok + ', "similarities": []}', False),                          # duplicate field
        ("{" + ok.replace(', "similarities": []', "") + "}", False),          # missing field
        ("[{}, {}, [], []]", True),                                           # the struct as a sequence
        ("[{}, {}, []]", False),
        ("[{}, {}, [], [], []]", False),
        ("null", False),
        ("{" + ok.replace('"similarities": []', '"similarities": [1, 2.5, -3e2, 0.0, -0, 1E+2]') + "}", True),
        ("{" + ok.replace('"similarities": []', '"similarities": [01]') + "}", False),     # leading zero
        ("{" + ok.replace('"similarities": []', '"similarities": [+1]') + "}", False),
        ("{" + ok.replace('"similarities": []', '"similarities": [.5]') + "}", False),
        ("{" + ok.replace('"similarities": []', '"similarities": [5.]') + "}", False),
        ("{" + ok.replace('"similarities": []', '"similarities": [1e]') + "}", False),
        ("{" + ok.replace('"similarities": []', '"similarities": [NaN]') + "}", False),
        ("{" + ok.replace('"similarities": []', '"similarities": [inf]') + "}", False),
        ("{" + ok.replace('"similarities": []', '"similarities": [0x10]') + "}", False),
        ("{" + ok.replace('"similarities": []', '"similarities": [null]') + "}", False),   # what the writer emits for NaN
        ("{" + ok.replace('"similarities": []', '"similarities": [1e999]') + "}", False),  # number out of range
        ("{" + ok.replace('"similarities": []', '"similarities": [1e-999]') + "}", True),  # underflows to zero
        ("{" + ok.replace('"similarities": []', '"similarities": [1e39]') + "}", True),    # f64 -> f32: +inf
        ("{" + ok.replace('"similarities": []', '"similarities": [1,]') + "}", False),     # trailing comma
        ("{" + ok.replace('"path_to_index": {}', '"path_to_index": {"a": 1}') + "}", True),
        ("{" + ok.replace('"path_to_index": {}', '"path_to_index": {"a": 1.0}') + "}", False),
        ("{" + ok.replace('"path_to_index": {}', '"path_to_index": {"a": -1}') + "}", False),
        ("{" + ok.replace('"path_to_index": {}', '"path_to_index": {"a": -0}') + "}", False),
        ("{" + ok.replace('"path_to_index": {}', '"path_to_index": {"a": 1e2}') + "}", False),
        ("{" + ok.replace('"path_to_index": {}', '"path_to_index": {"a": 18446744073709551615}') + "}", True),
        ("{" + ok.replace('"path_to_index": {}', '"path_to_index": {"a": 18446744073709551616}') + "}", False),
        ("{" + ok.replace('"path_to_index": {}', '"path_to_index": {"a": "1"}') + "}", False),
        ("{" + ok.replace('"index_to_path": {}', '"index_to_path": {"0": "a"}') + "}", True),
        ("{" + ok.replace('"index_to_path": {}', '"index_to_path": {"00": "a"}') + "}", False),
        ("{" + ok.replace('"index_to_path": {}', '"index_to_path": {" 0": "a"}') + "}", False),
        ("{" + ok.replace('"index_to_path": {}', '"index_to_path": {"+0": "a"}') + "}", False),
        ("{" + ok.replace('"index_to_path": {}', '"index_to_path": {"-1": "a"}') + "}", False),
        ("{" + ok.replace('"index_to_path": {}', '"index_to_path": {"1.0": "a"}') + "}", False),
        ("{" + ok.replace('"index_to_path": {}', '"index_to_path": {"\\u0030": "a"}') + "}", False),
        ("{" + ok.replace('"index_to_path": {}', '"index_to_path": {"": "a"}') + "}", False),
        ("{" + ok.replace('"index_to_path": {}', '"index_to_path": {0: "a"}') + "}", False),
        ("{" + ok.replace('"index_to_path": {}', '"index_to_path": {"0": "tab\there"}') + "}", False),   # raw control character
        ("{" + ok.replace('"index_to_path": {}', '"index_to_path": {"0": "tab\\there \\u00e9 \\ud83d\\uddbc \\/"}') + "}", True),
        ("{" + ok.replace('"index_to_path": {}', '"index_to_path": {"0": "\\ud83d"}') + "}", False),     # lone leading surrogate
        ("{" + ok.replace('"index_to_path": {}', '"index_to_path": {"0": "\\ud83d\\u0041"}') + "}", False),
        ("{" + ok.replace('"index_to_path": {}', '"index_to_path": {"0": "\\uddbc"}') + "}", False),     # lone trailing surrogate
        ("{" + ok.replace('"index_to_path": {}', '"index_to_path": {"0": "\\x41"}') + "}", False),
        ("{" + ok.replace('"lab_colors": []', '"lab_colors": [{"l": 1, "a": 2, "b": 3}]') + "}", True),
        ("{" + ok.replace('"lab_colors": []', '"lab_colors": [{"l": 1, "a": 2, "b": 3, "alpha": 1.0}]') + "}", True),
        ("{" + ok.replace('"lab_colors": []', '"lab_colors": [{"l": 1, "a": 2, "b": 3, "l": 1}]') + "}", False),
        ("{" + ok.replace('"lab_colors": []', '"lab_colors": [{"l": 1, "a": 2}]') + "}", False),
        ("{" + ok.replace('"lab_colors": []', '"lab_colors": [[1, 2, 3]]') + "}", True),
        ("{" + ok.replace('"lab_colors": []', '"lab_colors": [[1, 2]]') + "}", False),
        ("{" + ok.replace('"lab_colors": []', '"lab_colors": [[1, 2, 3, 4]]') + "}", False),
        ("{" + ok.replace('"lab_colors": []', '"lab_colors": {}') + "}", False),
        ("{" + ok + ', "deep": ' + "[" * 127 + "]" * 127 + "}", True),
        ("{" + ok + ', "deep": ' + "[" * 200 + "]" * 200 + "}", False),       # recursion limit exceeded
    ]
    f = tmp_path / "case.json"
    for text, loads in cases:
        f.write_bytes(text.encode("utf-8"))
        try:
            mosaic_b200.SimilarityDb.load(str(f)).close()
            got = True
        except mosaic_b200.MosaicError:
            got = False
        assert got == loads, text[:300]
    for raw in (b'{"path_to_index": {"\xff": 0}, "index_to_path": {}, "lab_colors": [], "similarities": []}',      # not UTF-8
                b'{"path_to_index": {"\xc0\xaf": 0}, "index_to_path": {}, "lab_colors": [], "similarities": []}',  # overlong
                b'{"path_to_index": {"\xed\xa0\x80": 0}, "index_to_path": {}, "lab_colors": [], "similarities": []}'):  # surrogate
        f.write_bytes(raw)
        with pytest.raises(mosaic_b200.MosaicError):
            mosaic_b200.SimilarityDb.load(str(f))


def test_values_of_unusual_number_literals(tmp_path):
    f = tmp_path / "n.json"
    lits = ["16777217", "-16777217", "9007199254740993", "18446744073709551615", "-9223372036854775808", "1e39", "-1e39",
            "1e-46", "-1e-999", "3.4028235e38", "3.4028236e38", "0.1", "1E2", "-0", "-0.0", "123456789012345678901234567890",
            "1152921573326323713", "-1152921573326323713"]
    f.write_text('{"path_to_index": {}, "index_to_path": {}, "lab_colors": [], "similarities": [' + ", ".join(lits) + "]}")
    _, sim = mosaic_b200.SimilarityDb.load(str(f)).arrays()
    with np.errstate(over="ignore"):
        want = np.array([np.float32(16777216.0), np.float32(-16777216.0),
                         np.float32(9007199254740992.0),          # 2^53 + 1: ties to even as an integer -> f32 in one step
                         np.float32(18446744073709551616.0), np.float32(-9223372036854775808.0),
                         np.float32(np.inf), np.float32(-np.inf), np.float32(0.0), np.float32(-0.0),
                         np.float32(3.4028235e38), np.float32(np.inf),     # past the midpoint between f32::MAX and 2^128
                         np.float32(0.1), np.float32(100.0), np.float32(-0.0), np.float32(-0.0),
                         np.float32(1.2345678901234568e29),                # beyond 64 bits: f64, then f32
                         # 2^60 + 2^36 + 1: ONE rounding (u64 / i64 `as f32`) goes up to 2^60 + 2^37; through f64 the +1 would
                         # be lost and the tie would go down to 2^60
                         np.float32(2.0 ** 60 + 2.0 ** 37), np.float32(-(2.0 ** 60 + 2.0 ** 37))], np.float32)
    assert (sim.view(np.uint32) == want.view(np.uint32)).all(), (sim, want)


def test_writer_refuses_paths_that_are_not_utf8(tmp_path):
    """serde's `Serialize for Path` fails on a non-UTF-8 path and save_to_file writes nothing (similarity.rs:112-116)."""
    import ctypes as C
    L = mosaic_b200.load_library()
    f = tmp_path / "x.json"
    paths = (C.c_char_p * 2)(b"fine.png", b"bad\xff.png")
    lab = np.zeros((2, 3), np.float32)
    sim = np.zeros(1, np.float32)
    rc = L.mosaic_simdb_save_json(os.fsencode(str(f)), paths, 2, lab.ctypes.data_as(C.POINTER(C.c_float)),
                                  sim.ctypes.data_as(C.POINTER(C.c_float)))
    assert rc != 0 and not f.exists()


def test_number_text_does_not_follow_the_locale(tmp_path):
    """A host that called setlocale() to a comma-decimal locale still writes and reads JSON numbers (no printf / strtod)."""
    import ctypes as C
    import locale
    chosen = None
    for name in ("de_DE.UTF-8", "fr_FR.UTF-8", "de_DE", "fr_FR", "ru_RU.UTF-8", "nl_NL.UTF-8"):
        try:
            locale.setlocale(locale.LC_NUMERIC, name)
            chosen = name
            break
        except locale.Error:
            continue
    if chosen is None or locale.localeconv()["decimal_point"] != ",":
        locale.setlocale(locale.LC_NUMERIC, "C")
        pytest.skip("no comma-decimal locale installed on this box")
    try:
        f = tmp_path / "l.json"
        lab = np.array([[50.5, -1.25, 0.1]], np.float32)
        mosaic_b200.SimilarityDb.save(str(f), ["a"], lab, np.zeros(0, np.float32))
        assert '"l": 50.5' in f.read_text()
        got, _ = mosaic_b200.SimilarityDb.load(str(f)).arrays()
        assert (got.view(np.uint32) == lab.view(np.uint32)).all()
    finally:
        locale.setlocale(locale.LC_NUMERIC, "C")


def test_float_text_over_random_bit_patterns(tmp_path):
    """ryu prints the shortest digits that read back as the same f32; so does numpy's `unique` mode.  30,000 random f32 BIT
    PATTERNS (denormals, the extremes): the writer's digits are numpy's, and the reader returns the bits."""
    rng = np.random.default_rng(5)
    vals = rng.integers(0, 2 ** 32, 40000, dtype=np.uint64).astype(np.uint32).view(np.float32)
    vals = vals[np.isfinite(vals)][:30000]
    n = len(vals) // 3
    lab = vals[:3 * n].reshape(n, 3)
    f = tmp_path / "bits.json"
    import ctypes as C
    L = mosaic_b200.load_library()                                          # (no matrix: the stale-list form of the writer)
    paths = (C.c_char_p * n)(*[b"p%d" % i for i in range(n)])
    L.mosaic_simdb_save_json_n.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                           C.c_uint64]
    assert L.mosaic_simdb_save_json_n(os.fsencode(str(f)), paths, n, lab.ctypes.data_as(C.POINTER(C.c_float)), None, 0) == 0
    nums = re.findall(r'"[lab]": (-?[0-9.e+-]+)', f.read_text())
    assert len(nums) == 3 * n
    for s, v in zip(nums, lab.ravel()):
        if v == 0:
            assert s in ("0.0", "-0.0")
            continue
        mine = re.sub(r"e.*$", "", s).replace("-", "").replace(".", "").lstrip("0").rstrip("0")
        ref = np.format_float_scientific(v, unique=True, trim="-").split("e")[0].replace("-", "").replace(".", "").rstrip("0")
        assert mine == ref and np.float32(float(s)) == v, (s, v)
    got, _ = mosaic_b200.SimilarityDb.load(str(f)).arrays()
    assert (got.view(np.uint32) == lab.view(np.uint32)).all()


def test_reader_loads_any_layout_of_a_valid_document(tmp_path):
    """The same database written in layouts the library's writer never produces — any key order, any whitespace, compact or
    indented, ASCII-escaped or raw UTF-8 strings, integers / exponents for floats, unknown fields at both levels, structs as
    sequences: serde_json reads them all (similarity.rs:121), so the reader must, with the same values."""
    rng = np.random.default_rng(77)
    f = tmp_path / "layout.json"
    for trial in range(300):
        n = int(rng.integers(0, 7))
        paths = [f"m/{i}-é🖼\"\\\t.png" if i % 3 == 0 else f"dir {i}/tile.JPG" for i in range(n)]
        lab = np.round(rng.uniform(-100, 100, (n, 3)), int(rng.integers(0, 5))).astype(np.float32)
        sim = np.round(rng.uniform(0, 250, n * (n - 1) // 2), int(rng.integers(0, 5))).astype(np.float32)

        def num(v):
            v = float(v)
            if v != 0 and v == int(v) and abs(v) < 1e15 and rng.integers(0, 2):   # (not -0.0: "0" is +0.0)
                return int(v)                                            # an integer literal where an f32 is expected
            return v

        def lab_struct(c):
            if rng.integers(0, 4) == 0:
                return [num(c[0]), num(c[1]), num(c[2])]                 # the struct as a sequence
            items = [("l", num(c[0])), ("a", num(c[1])), ("b", num(c[2]))]
            if rng.integers(0, 3) == 0:
                items.append(("alpha", [1, {"x": None}, "y"]))           # unknown field: skipped
            order = rng.permutation(len(items))
            return {items[i][0]: items[i][1] for i in order}

        fields = [("path_to_index", {p: i for i, p in enumerate(paths)}), ("index_to_path", {str(i): p for i, p in enumerate(paths)}),
                  ("lab_colors", [lab_struct(c) for c in lab]), ("similarities", [num(v) for v in sim])]
        if trial % 5 == 4:
            doc = [v for _, v in fields]                                  # the struct as a sequence of its fields
        else:
            if rng.integers(0, 2):
                fields.append(("version", {"major": 2, "notes": ["a", 1.5e300, True]}))
            order = rng.permutation(len(fields))
            doc = {fields[i][0]: fields[i][1] for i in order}
        indent = [None, 0, 1, 4, "\t"][int(rng.integers(0, 5))]
        seps = [(",", ":"), (", ", ": "), (" ,\n", " :\t")][int(rng.integers(0, 3))] if indent is None else None
        text = json.dumps(doc, indent=indent, separators=seps, ensure_ascii=bool(rng.integers(0, 2)))
        text = " \n\t" * int(rng.integers(0, 2)) + text + "\r\n " * int(rng.integers(0, 2))
        f.write_bytes(text.encode("utf-8"))
        db = mosaic_b200.SimilarityDb.load(str(f))
        assert db.info() == dict(rows=n, similarities=len(sim), paths=n), text[:300]
        glab, gsim = db.arrays()
        assert (glab.reshape(-1, 3).view(np.uint32) == lab.view(np.uint32)).all() and (gsim.view(np.uint32) == sim.view(np.uint32)).all(), text[:300]
        assert [db.path(i) for i in range(n)] == paths and [db.index_of(p) for p in paths] == list(range(n))
        db.close()


def test_utf8_boundaries(tmp_path):
    """`read_to_string` (similarity.rs:120) takes exactly the well-formed UTF-8 sequences: the first and last code point of every
    length load, overlong forms, surrogates, code points past U+10FFFF, stray continuation bytes and truncated sequences do not
    (Python's strict decoder is the model)."""
    f = tmp_path / "u.json"
    seqs = [b"\x7f", b"\xc2\x80", b"\xdf\xbf", b"\xe0\xa0\x80", b"\xed\x9f\xbf", b"\xee\x80\x80", b"\xef\xbf\xbf", b"\xf0\x90\x80\x80",
            b"\xf4\x8f\xbf\xbf", b"\xc0\x80", b"\xc1\xbf", b"\xe0\x9f\xbf", b"\xed\xa0\x80", b"\xed\xbf\xbf", b"\xf0\x8f\xbf\xbf",
            b"\xf4\x90\x80\x80", b"\xf5\x80\x80\x80", b"\xf8\x88\x80\x80\x80", b"\x80", b"\xbf", b"\xe2\x82", b"\xf0\x9f\x96", b"\xc2",
            b"\xc2\x41", b"\xe2\x41\x80", b"\xfe", b"\xff"]
    for seq in seqs:
        data = b'{"path_to_index": {"a' + seq + b'z": 0}, "index_to_path": {"0": "a' + seq + b'z"}, "lab_colors": [[1, 2, 3]], "similarities": []}'
        f.write_bytes(data)
        try:
            want = data.decode("utf-8")
        except UnicodeDecodeError:
            want = None
        try:
            db = mosaic_b200.SimilarityDb.load(str(f))
            got = db.path(0)
            db.close()
        except mosaic_b200.MosaicError:
            got = None
        assert (got is None) == (want is None), seq
        if want is not None:
            assert got == "a" + seq.decode("utf-8") + "z"
