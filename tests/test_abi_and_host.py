"""CPU tests (no GPU): the C-ABI library loads and exports every symbol include/mosaic_b200.h declares,
fails loudly without a device (no CPU fallback), and the host-side sharding logic of the multi-GPU path is
correct — including a world_size-2 gloo run that re-assembles sharded stages exactly like the NCCL path does."""
import ctypes
import os
import re
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mosaic_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mosaic_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(lib_built):
    import mosaic_b200
    lib = ctypes.CDLL(mosaic_b200.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in include/mosaic_b200.h but not exported"
    # the Python binding covers the same surface
    assert set(mosaic_b200.EXPORTED_SYMBOLS) == set(names)


def test_library_contains_sm100a_code_only(lib_built):
    import mosaic_b200
    out = subprocess.run(["cuobjdump", "-lelf", mosaic_b200.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_a_device(lib_built):
    import mosaic_b200
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(mosaic_b200.MosaicError) as e:
        mosaic_b200.Context(0)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_does_not_touch_the_oracle(lib_built):
    # the oracle is test infrastructure: nothing under the package or include/ may include, import, link or call it
    import mosaic_b200
    pats = [r'#\s*include\s*[<"][^">]*oracle', r"\bimport\s+oracle", r"\bfrom\s+oracle", r"liboracle", r"\borc_[a-z]",
            r"oracle/"]
    roots = [os.path.join(ROOT, "mosaic-art-rust_b200"), os.path.join(ROOT, "include")]
    for root in roots:
        for base, _, files in os.walk(root):
            if os.sep + "build" in base:
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".cc")) or f == "Makefile":
                    text = open(os.path.join(base, f), errors="ignore").read()
                    for pat in pats:
                        assert not re.search(pat, text), (os.path.join(base, f), pat)
    # and the shared library neither links nor exports oracle symbols
    nm = subprocess.run(["nm", "-D", mosaic_b200.LIB_PATH], capture_output=True, text=True).stdout
    assert "orc_" not in nm
    ldd = subprocess.run(["ldd", mosaic_b200.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in ldd


def test_shard_plan_tiles_every_range(lib_built):
    import mosaic_b200
    for world in (1, 2, 3, 4, 8):
        for gh, n in ((120, 10000), (28, 500), (7, 3), (1000, 100000), (1, 1), (5, 50000)):
            shards = [mosaic_b200.shard_plan(world, r, gh, n) for r in range(world)]
            total_tri = n * (n - 1) // 2
            for key, total in (("rows", gh), ("materials", n), ("tri", total_tri)):
                assert shards[0][key][0] == 0 and shards[-1][key][1] == total
                for a, b in zip(shards, shards[1:]):
                    assert a[key][1] == b[key][0]           # contiguous, no gap, no overlap
                assert all(p[key][0] <= p[key][1] for p in shards)
            sizes = [p["rows"][1] - p["rows"][0] for p in shards]
            assert max(sizes) - min(sizes) <= 1
            tri = [p["tri"][1] - p["tri"][0] for p in shards]
            if total_tri > 512 * world * 8:                  # balanced by entry count (not by row count)
                assert max(tri) - min(tri) <= 2 * 512 + world
            assert all(p["tri"][0] % 512 == 0 for p in shards)  # 16-byte aligned vector stores on every rank
    with pytest.raises(mosaic_b200.MosaicError):
        mosaic_b200.shard_plan(2, 2, 10, 10)


WORKER = r'''
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "mosaic-art-rust_b200"))
import numpy as np, torch, torch.distributed as dist
import mosaic_b200
from mosaic_b200 import synth
from oracle import oracle as orc
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
gw, gh, n, k = 12, 9, 150, 100
mats = synth.materials(n, w=16, h=16)
tgt = synth.target(gw * 10, gh * 8)
shards = [mosaic_b200.shard_plan(world, r, gh, n) for r in range(world)]
mine = shards[rank]
mb, me = mine["materials"]; rb, re_ = mine["rows"]; tb, te = mine["tri"]

def gather_slices(part, shape_of_rank):
    """allgather of unequal contiguous slices: one byte broadcast per source rank (what dist.cu does with NCCL)"""
    part = np.ascontiguousarray(part)
    full = []
    for r in range(world):
        shape = tuple(shape_of_rank(r))
        nbytes = int(np.prod(shape)) * part.dtype.itemsize
        buf = torch.from_numpy(part.view(np.uint8).reshape(-1).copy()) if r == rank else torch.empty(nbytes, dtype=torch.uint8)
        assert buf.numel() == nbytes
        dist.broadcast(buf, src=r)
        full.append(buf.numpy().view(part.dtype).reshape(shape))
    return np.concatenate(full, axis=0)

rows_of = lambda r: shards[r]["rows"][1] - shards[r]["rows"][0]
# each rank computes only ITS slices (what its GPU would compute) ...
mlab = gather_slices(orc.material_lab(mats[mb:me]), lambda r: (shards[r]["materials"][1] - shards[r]["materials"][0], 3))
th = tgt.shape[0] // gh
clab_part = orc.cell_lab(np.ascontiguousarray(tgt[rb * th:re_ * th]), gw, re_ - rb)
idx_part, d2_part = orc.topk(mlab, clab_part, k)
idx = gather_slices(idx_part, lambda r: (rows_of(r) * gw, k))
d2 = gather_slices(d2_part, lambda r: (rows_of(r) * gw, k))
tri_full = orc.similarity_tri(mlab)
tri = gather_slices(tri_full[tb:te], lambda r: (shards[r]["tri"][1] - shards[r]["tri"][0],))
# ... rank 0 runs the serial chain and broadcasts the placement
placed = torch.zeros(gw * gh, dtype=torch.int32)
if rank == 0:
    p, nfb, _, _ = orc.place(idx, d2, mlab, gw, gh, 3, 0.3)
    g, _ = orc.anneal(mlab, p.reshape(gh, gw), 300, 42)
    placed = torch.from_numpy(g.ravel().copy())
dist.broadcast(placed, src=0)
out_part = orc.render(tgt, gw, gh, mats, placed.numpy(), 0.3)[rb * th:re_ * th]
out = gather_slices(out_part, lambda r: (rows_of(r) * th, out_part.shape[1], 3))
# reference: the unsharded pipeline
rm = orc.material_lab(mats); rc = orc.cell_lab(tgt, gw, gh)
ridx, rd2 = orc.topk(rm, rc, k)
rp, _, _, _ = orc.place(ridx, rd2, rm, gw, gh, 3, 0.3)
rg, _ = orc.anneal(rm, rp.reshape(gh, gw), 300, 42)
rout = orc.render(tgt, gw, gh, mats, rg.ravel(), 0.3)
assert (mlab == rm).all() and (idx == ridx).all() and (d2.view(np.uint32) == rd2.view(np.uint32)).all()
assert (tri == tri_full).all()
assert (placed.numpy() == rg.ravel()).all() and (out == rout).all()
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_sharded_pipeline_gloo_world2(lib_built, tmp_path):
    """world_size 2 on CPU (gloo): the partition + 'allgather of unequal slices' + 'rank 0 runs the chain,
    broadcasts' logic of the multi-GPU path reproduces the unsharded result exactly."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o}"
        assert f"rank {r} ok" in o


def test_reference_arm_of_the_bench_runs_on_cpu(tmp_path):
    """`bench.py --impl reference` (the CPU arm the driver times next to the GPU arm) needs no GPU: it prints one JSON line
    with the contract's keys, the same `config` object as the GPU arm and a `cpu_baseline` that describes the run."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--config", "C1", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "cells_placed_per_sec" and line["unit"] == "cells/s"
    assert line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("C1:")


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/mosaic_b200.h must compile as C99 (what cgo, bindgen or a C host would feed it to),
    with no C++ or torch types in the signatures."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "use_header.c"
    src.write_text('#include "mosaic_b200.h"\n'
                   "int main(void) { mosaic_params p; mosaic_report r; (void)p; (void)r; return MOSAIC_OK; }\n")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"),
                           "-fsyntax-only", str(src)])
    text = open(os.path.join(root, "include", "mosaic_b200.h")).read()
    assert "torch::" not in text and "std::" not in text and "at::" not in text and "#include <torch" not in text


# ---------------------------------------------------------------- the Rust side (unverified source) against the header
_RUST_TO_C = {
    "c_int": "int", "i32": "int32_t", "u32": "uint32_t", "u64": "uint64_t", "i64": "int64_t", "c_float": "float", "f32": "float",
    "u8": "uint8_t", "c_char": "char", "c_void": "void", "usize": "size_t", "mosaic_ctx": "mosaic_ctx", "mosaic_simdb": "mosaic_simdb",
    "mosaic_opt_result": "mosaic_opt_result", "mosaic_params": "mosaic_params", "mosaic_report": "mosaic_report",
}


def _rust_type_to_c(t):
    t = t.strip()
    stars = []
    while t.startswith("*"):
        m = re.match(r"\*(const|mut)\s+", t)
        stars.append(m.group(1))
        t = t[m.end():]
    base = _RUST_TO_C[t]
    # `*const T` -> `const T*`; `*mut *mut T` -> `T**`; `*const *const c_char` -> `const char* const*`
    out = base
    for kind in reversed(stars):
        out = ("const " + out if kind == "const" and "*" not in out else out + (" const" if kind == "const" else "")) + "*"
    return re.sub(r"\s+", " ", out).strip()


def _norm_c(t):
    t = re.sub(r"\s+", " ", t.strip())
    t = re.sub(r"\s*\*\s*", "*", t)
    t = re.sub(r"\bint32_t\b", "int32_t", t)
    return t


def _c_params(decl):
    out = []
    for p in decl.split(","):
        p = p.strip()
        if p in ("", "void"):
            continue
        p = re.sub(r"\[[^\]]*\]", "", p)
        m = re.match(r"(.*?[\s\*])([A-Za-z_][A-Za-z0-9_]*)$", p)         # strip the parameter name
        out.append(_norm_c(m.group(1)))
    return out


def test_rust_ffi_declarations_match_the_header():
    """integration/rust/ cannot be compiled here (no rustc): at least its `extern "C"` blocks are held against the header —
    every function exists, with the same number of parameters, the same C types in the same order and the same return type;
    the #[repr(C)] structs list the header's fields in the header's order."""
    hdr = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    c_decls = {}
    for m in re.finditer(r"^([A-Za-z_][A-Za-z0-9_ \*]*?)\b(mosaic_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.M | re.S):
        c_decls[m.group(2)] = (_norm_c(m.group(1)), _c_params(m.group(3)))
    checked = 0
    sources = [(name, open(os.path.join(ROOT, "integration", "rust", "src", name)).read()) for name in ("b200.rs", "lib_compat.rs")]
    # the same declarations as INTEGRATION.md prints them
    sources.append(("INTEGRATION.md", "\n".join(re.findall(r"```rust\n(.*?)```", open(os.path.join(ROOT, "INTEGRATION.md")).read(), flags=re.S))))
    for name, src in sources:
        src = re.sub(r"//[^\n]*", "", src)
        for block in re.findall(r'extern\s+"C"\s*\{(.*?)\n\}', src, flags=re.S):
            for m in re.finditer(r"fn\s+(mosaic_[a-z0-9_]+)\s*\((.*?)\)\s*(?:->\s*([^;]+))?;", block, flags=re.S):
                fn, params, ret = m.group(1), m.group(2), (m.group(3) or "").strip()
                assert fn in c_decls, f"{name}: {fn} is not declared in include/mosaic_b200.h"
                c_ret, c_params = c_decls[fn]
                r_params = [_norm_c(_rust_type_to_c(p.split(":", 1)[1])) for p in re.split(r",\s*(?=[a-z_0-9]+\s*:)", params.strip()) if p.strip()]
                assert r_params == c_params, f"{name}: {fn}: {r_params} != {c_params}"
                r_ret = _norm_c(_rust_type_to_c(ret)) if ret else "void"
                assert r_ret == c_ret, f"{name}: {fn}: returns {r_ret}, the header says {c_ret}"
                checked += 1
    assert checked >= 30
    # struct layouts: field names and order
    b200 = open(os.path.join(ROOT, "integration", "rust", "src", "b200.rs")).read()
    for struct in ("mosaic_opt_result", "mosaic_params", "mosaic_report"):
        c_body = re.search(r"typedef struct \{([^}]*)\}\s*" + struct + r"\s*;", hdr, flags=re.S).group(1)
        c_fields = [(re.sub(r"\s+", " ", t.strip()), n) for t, n in re.findall(r"([A-Za-z_][A-Za-z0-9_ ]*?)\s+([a-z_0-9]+(?:\s*,\s*[a-z_0-9]+)*)\s*;", c_body)]
        c_flat = [(t, n.strip()) for t, names in c_fields for n in names.split(",")]
        r_body = re.search(r"pub struct " + struct + r"\s*\{(.*?)\n\}", b200, flags=re.S).group(1)
        r_body = re.sub(r"//[^\n]*", "", r_body)
        r_flat = [(_RUST_TO_C[t.strip()], n) for n, t in re.findall(r"pub\s+([a-z_0-9]+)\s*:\s*([A-Za-z_0-9]+)\s*,", r_body)]
        assert [(t.replace("int", "int") if t != "int" else "int", n) for t, n in r_flat] == c_flat, (struct, r_flat, c_flat)


def test_context_less_errors_are_per_thread(lib_built, tmp_path):
    """The codecs and the database reader have no context and may run on many threads at once (the reference decodes its
    tiles on rayon workers, main.rs:193-208): every thread reads the message of ITS failed call from mosaic_last_error(NULL)."""
    import threading
    import mosaic_b200
    L = ctypes.CDLL(mosaic_b200.LIB_PATH)
    L.mosaic_last_error.restype = ctypes.c_char_p
    L.mosaic_last_error.argtypes = [ctypes.c_void_p]
    L.mosaic_simdb_load_json.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]
    n_threads, rounds = 8, 300
    bad = []
    barrier = threading.Barrier(n_threads)

    def worker(t):
        name = f"no_such_database_{t}.json"
        path = os.fsencode(str(tmp_path / name))
        barrier.wait()
        for _ in range(rounds):
            h = ctypes.c_void_p()
            rc = L.mosaic_simdb_load_json(path, ctypes.byref(h))
            msg = (L.mosaic_last_error(None) or b"").decode()
            if rc == 0 or name not in msg:
                bad.append((t, rc, msg))
                return

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(n_threads)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not bad, bad[:3]
