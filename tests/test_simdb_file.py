"""The similarity-database FILE (reference src/similarity.rs:31-138): writer/reader of the
serde_json format behind the C ABI.  Host-only: runs without a GPU."""
from __future__ import annotations

import json
import os

import numpy as np
import pytest

import mosaic_b200
from mosaic_b200 import synth
from oracle import oracle as orc

# a file as the reference writes it: HashMap order is arbitrary, index_to_path keys are strings, 2-space pretty
# layout (similarity.rs:112-116); fixture values = the reference's own test tiles (similarity.rs:186-192)
REFERENCE_STYLE = """{
  "path_to_index": {
    "tiles/tile3.png": 2,
    "tiles/tile1.png": 0,
    "tiles/tile2.png": 1
  },
  "index_to_path": {
    "1": "tiles/tile2.png",
    "2": "tiles/tile3.png",
    "0": "tiles/tile1.png"
  },
  "lab_colors": [
    {
      "l": 50.0,
      "a": 0.0,
      "b": 0.0
    },
    {
      "l": 60.0,
      "a": 10.0,
      "b": 10.0
    },
    {
      "l": 40.0,
      "a": -10.0,
      "b": -10.0
    }
  ],
  "similarities": [
    17.320509,
    17.320509,
    34.641018
  ]
}"""


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__ as g
    g.build()


def test_load_reference_style_file(tmp_path):
    f = tmp_path / "similarity_db.json"
    f.write_text(REFERENCE_STYLE)
    db = mosaic_b200.SimilarityDb.load(str(f))
    assert db.info() == dict(rows=3, similarities=3, paths=3)
    assert [db.path(i) for i in range(3)] == ["tiles/tile1.png", "tiles/tile2.png", "tiles/tile3.png"]
    assert db.index_of("tiles/tile3.png") == 2 and db.index_of("nope.png") == -1
    lab, sim = db.arrays()
    assert (lab == np.array([[50, 0, 0], [60, 10, 10], [40, -10, -10]], np.float32)).all()
    # similarity.rs:185-237: symmetric access, zero on the diagonal, None for unknown paths
    assert db.get_similarity("tiles/tile1.png", "tiles/tile2.png") == np.float32(17.320509)
    assert db.get_similarity("tiles/tile2.png", "tiles/tile1.png") == np.float32(17.320509)
    assert db.get_similarity("tiles/tile3.png", "tiles/tile2.png") == np.float32(34.641018)
    assert db.get_similarity("tiles/tile2.png", "tiles/tile2.png") == 0.0
    assert db.get_similarity("tiles/tile2.png", "other.png") is None
    # the stored matrix is what build_similarities computes from the stored Lab values (similarity.rs:69-82)
    assert (sim.view(np.uint32) == orc.similarity_tri(lab).view(np.uint32)).all()


def test_writer_matches_serde_pretty_layout(tmp_path):
    f = tmp_path / "db.json"
    lab = np.array([[50, 0, 0], [60, 10, 10], [40, -10, -10]], np.float32)
    mosaic_b200.SimilarityDb.save(str(f), ["tiles/tile1.png", "tiles/tile2.png", "tiles/tile3.png"], lab,
                                  orc.similarity_tri(lab))
    text = f.read_text()
    # same bytes as the reference-style file once its (arbitrary) HashMap order is put in index order
    ref = json.loads(REFERENCE_STYLE)
    doc = json.loads(text)
    assert doc == ref                                       # same JSON value
    lines = text.split("\n")
    assert lines[0] == "{" and lines[1] == '  "path_to_index": {' and lines[2] == '    "tiles/tile1.png": 0,'
    assert '    "0": "tiles/tile1.png",' in lines and "    17.320509," in lines and lines[-1] == "}"
    assert '      "l": 50.0,' in lines and '      "b": -10.0' in lines   # serde prints integral floats with ".0"
    assert not text.endswith("\n")                          # to_string_pretty has no trailing newline


def test_round_trip_is_bit_exact(tmp_path):
    n = 300
    lab = synth.lab_points(n)
    sim = orc.similarity_tri(lab)
    paths = [f"materials/ü {i:04d} \"q\"\\x.png" for i in range(n)]   # UTF-8 and characters JSON must escape
    f = tmp_path / "db.json"
    mosaic_b200.SimilarityDb.save(str(f), paths, lab, sim)
    db = mosaic_b200.SimilarityDb.load(str(f))
    lab2, sim2 = db.arrays()
    assert (lab2.view(np.uint32) == lab.view(np.uint32)).all()
    assert (sim2.view(np.uint32) == sim.view(np.uint32)).all()
    assert [db.path(i) for i in (0, 7, n - 1)] == [paths[0], paths[7], paths[n - 1]]
    # python's json reads the same document (valid JSON, valid escapes)
    doc = json.loads(f.read_text())
    assert doc["path_to_index"][paths[5]] == 5 and doc["index_to_path"]["5"] == paths[5]
    assert np.float32(doc["similarities"][10]) == sim[10]


def test_f32_text_is_shortest_round_trip(tmp_path):
    # ryu prints the shortest digit string that parses back to the same f32; numpy's repr does the same
    rng = np.random.default_rng(0)
    vals = np.concatenate([rng.uniform(-200, 200, 500), 10.0 ** rng.uniform(-12, 12, 300),
                           [0.0, 1.0, -1.0, 0.1, 100.0, 1e-7, 1.5e10, 12345678.0, 1e16, 1e-6, 123456.7, 0.000123]]).astype(np.float32)
    n = len(vals) // 3
    lab = vals[:3 * n].reshape(n, 3)
    f = tmp_path / "db.json"
    mosaic_b200.SimilarityDb.save(str(f), [f"p{i}" for i in range(n)], lab, np.zeros(n * (n - 1) // 2, np.float32))
    doc = json.loads(f.read_text())
    text = f.read_text()
    got = np.array([[c["l"], c["a"], c["b"]] for c in doc["lab_colors"]], np.float64).astype(np.float32)
    assert (got.view(np.uint32) == lab.view(np.uint32)).all()
    # digit count: never more significant digits than numpy's shortest repr
    import re
    nums = re.findall(r'"[lab]": (-?[0-9.e+-]+)', text)
    assert len(nums) == 3 * n
    for s, v in zip(nums, lab.ravel()):
        digits = re.sub(r"e.*$", "", s).replace("-", "").replace(".", "").lstrip("0").rstrip("0")
        ref = np.format_float_scientific(v, unique=True, trim="-").split("e")[0].replace("-", "").replace(".", "").rstrip("0")
        assert len(digits) <= max(len(ref), 1), (s, v)
    # ryu's layout rules
    fmt = {1.0: "1.0", 0.1: "0.1", 100.0: "100.0", 1e-7: "1e-7", 1.5e10: "15000000000.0", 12345678.0: "12345678.0",
           1e16: "1e16", 1e-6: "0.000001", 0.0: "0.0", -1.0: "-1.0"}
    for v, s in fmt.items():
        assert re.search(r'": ' + re.escape(s) + r"[,\n]", text), (v, s)


def test_load_errors_are_loud(tmp_path):
    f = tmp_path / "bad.json"
    f.write_text('{"path_to_index": {}, "index_to_path": {}, "lab_colors": []}')   # serde: missing field `similarities`
    with pytest.raises(mosaic_b200.MosaicError):
        mosaic_b200.SimilarityDb.load(str(f))
    with pytest.raises(mosaic_b200.MosaicError):
        mosaic_b200.SimilarityDb.load(str(tmp_path / "does_not_exist.json"))
    f.write_text('{"path_to_index": {}, "index_to_path": {}, "lab_colors": [], "similarities": []}')
    db = mosaic_b200.SimilarityDb.load(str(f))     # SimilarityDatabase::new() round trip (similarity.rs:239-247)
    assert db.info() == dict(rows=0, similarities=0, paths=0)


@pytest.mark.gpu
def test_context_writes_the_database_file(tmp_path):
    n = 64
    mats = synth.materials(n)
    with mosaic_b200.Context(0) as ctx:
        ctx.set_materials(mats)
        mlab = ctx.material_lab()
        paths = [f"m/{i}.png" for i in range(n)]
        f = tmp_path / "similarity_db.json"
        ctx.save_similarity_db(str(f), paths)
    db = mosaic_b200.SimilarityDb.load(str(f))
    lab, sim = db.arrays()
    assert (lab.view(np.uint32) == mlab.view(np.uint32)).all()
    assert (sim.view(np.uint32) == orc.similarity_tri(mlab).view(np.uint32)).all()


# ---------------------------------------------------------------- fuzz (hypothesis): any paths, any finite floats
def test_fuzz_round_trip_and_json_validity(tmp_path):
    """Random databases — paths with quotes, backslashes, control characters, non-ASCII; floats over the whole finite
    f32 range incl. denormals and negative zero: the file must be valid JSON that Python's parser reads to the same
    values, and the library's own reader must return every bit."""
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")
    f32 = st.floats(width=32, allow_nan=False, allow_infinity=False)
    path = st.text(st.characters(blacklist_categories=("Cs",), blacklist_characters="\x00"), min_size=1, max_size=24)

    @hyp.settings(max_examples=60, deadline=None, suppress_health_check=list(hyp.HealthCheck))
    @hyp.given(st.lists(path, min_size=0, max_size=7, unique=True), st.data())
    def run(paths, data):
        n = len(paths)
        lab = np.array(data.draw(st.lists(f32, min_size=3 * n, max_size=3 * n)), np.float32).reshape(n, 3)
        m = n * (n - 1) // 2
        sim = np.array(data.draw(st.lists(f32, min_size=m, max_size=m)), np.float32)
        file = str(tmp_path / "fuzz.json")
        mosaic_b200.SimilarityDb.save(file, paths, lab, sim)
        doc = json.load(open(file, encoding="utf-8"))           # valid JSON, UTF-8
        assert doc["path_to_index"] == {p: i for i, p in enumerate(paths)}
        assert doc["index_to_path"] == {str(i): p for i, p in enumerate(paths)}
        got_lab = np.array([[c["l"], c["a"], c["b"]] for c in doc["lab_colors"]], np.float64).reshape(n, 3)
        assert (got_lab.astype(np.float32).view(np.uint32) == lab.view(np.uint32)).all()   # shortest text round-trips
        assert (np.array(doc["similarities"], np.float64).astype(np.float32).view(np.uint32) == sim.view(np.uint32)).all()
        db = mosaic_b200.SimilarityDb.load(file)
        try:
            info = db.info()
            assert info["rows"] == n and info["similarities"] == m and info["paths"] == n
            l2, s2 = db.arrays()
            assert (l2.view(np.uint32) == lab.view(np.uint32)).all() and (s2.view(np.uint32) == sim.view(np.uint32)).all()
            for i, p in enumerate(paths):
                assert db.path(i) == p and db.index_of(p) == i
            if n >= 2:
                assert np.float32(db.get_similarity(paths[0], paths[1])).view(np.uint32) == sim[:1].view(np.uint32)[0]
                assert db.get_similarity(paths[0], paths[0]) == 0.0
        finally:
            db.close()

    run()
