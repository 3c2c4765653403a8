"""mosaic_b200 — thin ctypes binding of libmosaic_b200.so (include/mosaic_b200.h).

This is the Python face of the C ABI used by tests/, bench.py and __graft_entry__.py.  It contains no
compute: every method forwards to a C entry point that launches hand-written sm_100a kernels, and it
raises MosaicError when the shared library or a CUDA device is missing (there is no CPU fallback).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_PKG_DIR), "libmosaic_b200.so")

MAX_K = 100
STAGES = ("material_lab", "cell_lab", "topk", "place", "similarity", "anneal", "render")


class MosaicError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"mosaic_b200 error {code}: {msg}")
        self.code = code


class OptResult(C.Structure):
    """OptimizationResult (reference src/optimizer.rs:203-210)."""
    _fields_ = [("initial_cost", C.c_float), ("final_cost", C.c_float), ("best_cost", C.c_float),
                ("improved_count", C.c_uint64), ("accepted_count", C.c_uint64), ("iterations", C.c_uint64)]

    def improvement_percentage(self):  # optimizer.rs:212-220
        if self.initial_cost > 0.0:
            return float(np.float32((np.float32(self.initial_cost) - np.float32(self.final_cost))
                                    / np.float32(self.initial_cost)) * np.float32(100.0))
        return 0.0


class Params(C.Structure):
    _fields_ = [("max_usage", C.c_int), ("adjacency_weight", C.c_float), ("color_strength", C.c_float),
                ("enable_optimization", C.c_int), ("optimization_iterations", C.c_uint64), ("seed", C.c_uint64),
                ("build_similarity", C.c_int)]


class Report(C.Structure):
    _fields_ = [("n_fallbacks", C.c_uint32), ("opt", OptResult), ("annealed", C.c_int)]


_lib = None


PROGRESS_FN = C.CFUNCTYPE(None, C.c_float, C.c_int, C.c_char_p, C.c_void_p)  # mosaic_progress_fn


def load_library():
    """Load libmosaic_b200.so; fails loudly when it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MosaicError(-100, f"{LIB_PATH} is missing: build it with `make -C {os.path.dirname(LIB_PATH)}` "
                                "(python -c 'import __graft_entry__ as g; g.build()')")
    L = C.CDLL(LIB_PATH)
    vp, u8p, f32p, i32p, u32p, u64p = (C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_float),
                                       C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64))
    sig = {
        "mosaic_ctx_create": ([C.POINTER(vp), C.c_int], C.c_int),
        "mosaic_ctx_create_sharded": ([C.POINTER(vp), C.c_int, C.c_int, C.c_int, vp, C.c_size_t], C.c_int),
        "mosaic_nccl_unique_id": ([vp, C.c_size_t], C.c_int),
        "mosaic_ctx_destroy": ([vp], None),
        "mosaic_last_error": ([vp], C.c_char_p),
        "mosaic_ctx_set_stream": ([vp, vp], C.c_int),
        "mosaic_ctx_sync": ([vp], C.c_int),
        "mosaic_stage_ms": ([vp, C.c_int, f32p], C.c_int),
        "mosaic_kernel_launches": ([vp], C.c_uint64),
        "mosaic_collective_ms": ([vp, f32p], C.c_int),
        "mosaic_sm_count": ([vp], C.c_int),
        "mosaic_exchange_path": ([vp], C.c_char_p),
        "mosaic_similarity_local": ([vp, vp], C.c_int),
        "mosaic_set_materials_rgb8": ([vp, vp, C.c_int, C.c_int, C.c_int], C.c_int),
        "mosaic_set_materials_ragged_rgb8": ([vp, u8p, C.POINTER(C.c_uint64), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int], C.c_int),
        "mosaic_set_materials_device": ([vp, vp, C.c_int, C.c_int, C.c_int], C.c_int),
        "mosaic_set_target_rgb8": ([vp, vp, C.c_int, C.c_int, C.c_int, C.c_int], C.c_int),
        "mosaic_set_target_device": ([vp, vp, C.c_int, C.c_int, C.c_int, C.c_int], C.c_int),
        "mosaic_set_material_lab": ([vp, f32p, C.c_int], C.c_int),
        "mosaic_set_similarity_lab": ([vp, f32p, C.c_int], C.c_int),
        "mosaic_set_cell_lab": ([vp, f32p, C.c_int, C.c_int], C.c_int),
        "mosaic_material_lab": ([vp, vp], C.c_int),
        "mosaic_cell_lab": ([vp, vp], C.c_int),
        "mosaic_similarity_tri": ([vp, vp], C.c_int),
        "mosaic_candidates": ([vp, C.c_int, vp, vp], C.c_int),
        "mosaic_place": ([vp, C.c_int, C.c_float, vp, u32p], C.c_int),
        "mosaic_fallback_cells": ([vp, u32p, C.c_uint32, u32p], C.c_int),
        "mosaic_set_placement": ([vp, i32p], C.c_int),
        "mosaic_anneal": ([vp, C.c_uint64, C.c_float, C.c_float, C.c_uint64, vp, C.POINTER(OptResult)], C.c_int),
        "mosaic_optimize_greedy": ([vp, C.c_uint64, C.c_uint64, vp, C.POINTER(OptResult)], C.c_int),
        "mosaic_render": ([vp, C.c_float, vp], C.c_int),
        "mosaic_generate": ([vp, C.POINTER(Params), vp, vp, C.POINTER(Report)], C.c_int),
        "mosaic_selftest": ([vp, u64p], C.c_int),
        "mosaic_debug_chain_profile": ([vp, C.c_int, u64p], C.c_int),
        "mosaic_debug_set_chain": ([vp, C.c_int], C.c_int),
        "mosaic_dims": ([vp] + [C.POINTER(C.c_int)] * 5, C.c_int),
        "mosaic_device_output": ([vp], vp),
        "mosaic_device_placement": ([vp], vp),
        "mosaic_device_similarity": ([vp], vp),
        "mosaic_shard": ([vp] + [C.POINTER(C.c_int)] * 4 + [u64p, u64p], C.c_int),
        "mosaic_shard_plan": ([C.c_int] * 4 + [C.POINTER(C.c_int)] * 4 + [u64p, u64p], C.c_int),
        "mosaic_simdb_save_json": ([C.c_char_p, C.POINTER(C.c_char_p), C.c_int, f32p, f32p], C.c_int),
        "mosaic_simdb_save_json_n": ([C.c_char_p, C.POINTER(C.c_char_p), C.c_int, f32p, f32p, C.c_uint64], C.c_int),
        "mosaic_simdb_save_ctx": ([vp, C.c_char_p, C.POINTER(C.c_char_p)], C.c_int),
        "mosaic_simdb_load_json": ([C.c_char_p, C.POINTER(vp)], C.c_int),
        "mosaic_simdb_free": ([vp], None),
        "mosaic_simdb_info": ([vp, C.POINTER(C.c_int), u64p, C.POINTER(C.c_int)], C.c_int),
        "mosaic_simdb_path": ([vp, C.c_int], C.c_char_p),
        "mosaic_simdb_index_of": ([vp, C.c_char_p], C.c_int64),
        "mosaic_simdb_copy": ([vp, f32p, f32p], C.c_int),
        "mosaic_simdb_get_similarity": ([vp, C.c_char_p, C.c_char_p, f32p], C.c_int),
        "mosaic_set_progress_callback": ([vp, PROGRESS_FN, vp], C.c_int),
        "mosaic_png_decode_rgb8": ([C.c_char_p, C.POINTER(u8p), C.POINTER(C.c_int), C.POINTER(C.c_int)], C.c_int),
        "mosaic_png_encode_rgb8": ([C.c_char_p, vp, C.c_int, C.c_int], C.c_int),
        "mosaic_jpeg_encode_rgb8": ([C.c_char_p, vp, C.c_int, C.c_int, C.c_int], C.c_int),
        "mosaic_jpeg_decode_rgb8": ([C.c_char_p, C.POINTER(u8p), C.POINTER(C.c_int), C.POINTER(C.c_int)], C.c_int),
        "mosaic_image_decode_rgb8": ([C.c_char_p, C.POINTER(u8p), C.POINTER(C.c_int), C.POINTER(C.c_int)], C.c_int),
        "mosaic_image_open_rgb8": ([C.c_char_p, C.POINTER(u8p), C.POINTER(C.c_int), C.POINTER(C.c_int)], C.c_int),
        "mosaic_webp_decode_rgb8": ([C.c_char_p, C.POINTER(u8p), C.POINTER(C.c_int), C.POINTER(C.c_int)], C.c_int),
        "mosaic_webp_encode_rgb8": ([C.c_char_p, vp, C.c_int, C.c_int], C.c_int),
        "mosaic_host_free": ([vp], None),
        "mosaic_host_alloc": ([C.c_size_t, C.POINTER(vp)], C.c_int),
        "mosaic_host_release": ([vp], C.c_int),
        "mosaic_host_register": ([vp, C.c_size_t], C.c_int),
        "mosaic_host_unregister": ([vp], C.c_int),
    }
    for name, (args, res) in sig.items():
        fn = getattr(L, name)  # AttributeError here = the library does not export what the header declares
        fn.argtypes = args
        fn.restype = res
    _lib = L
    return L


EXPORTED_SYMBOLS = (
    "mosaic_ctx_create", "mosaic_ctx_create_sharded", "mosaic_nccl_unique_id", "mosaic_ctx_destroy",
    "mosaic_last_error", "mosaic_ctx_set_stream", "mosaic_ctx_sync", "mosaic_stage_ms", "mosaic_kernel_launches",
    "mosaic_collective_ms", "mosaic_sm_count", "mosaic_similarity_local", "mosaic_exchange_path",
    "mosaic_set_materials_rgb8", "mosaic_set_materials_ragged_rgb8", "mosaic_set_materials_device", "mosaic_set_target_rgb8",
    "mosaic_set_target_device", "mosaic_set_material_lab", "mosaic_set_similarity_lab", "mosaic_set_cell_lab", "mosaic_material_lab",
    "mosaic_cell_lab", "mosaic_similarity_tri", "mosaic_candidates", "mosaic_place", "mosaic_fallback_cells",
    "mosaic_set_placement", "mosaic_anneal", "mosaic_optimize_greedy", "mosaic_render", "mosaic_generate", "mosaic_selftest", "mosaic_dims",
    "mosaic_debug_chain_profile", "mosaic_debug_set_chain",
    "mosaic_device_output", "mosaic_device_placement", "mosaic_device_similarity", "mosaic_shard",
    "mosaic_shard_plan",
    "mosaic_simdb_save_json", "mosaic_simdb_save_json_n", "mosaic_simdb_save_ctx", "mosaic_simdb_load_json", "mosaic_simdb_free",
    "mosaic_simdb_info", "mosaic_simdb_path", "mosaic_simdb_index_of", "mosaic_simdb_copy",
    "mosaic_simdb_get_similarity",
    "mosaic_set_progress_callback", "mosaic_png_decode_rgb8", "mosaic_png_encode_rgb8", "mosaic_jpeg_decode_rgb8", "mosaic_jpeg_encode_rgb8", "mosaic_image_decode_rgb8", "mosaic_image_open_rgb8", "mosaic_webp_decode_rgb8", "mosaic_webp_encode_rgb8", "mosaic_host_free",
    "mosaic_host_alloc", "mosaic_host_release", "mosaic_host_register", "mosaic_host_unregister",
)


def shard_plan(world, rank, grid_h, n_materials):
    """Partition of rows / materials / packed similarity entries for one rank (pure host function)."""
    L = load_library()
    v = [C.c_int() for _ in range(4)]
    t = [C.c_uint64(), C.c_uint64()]
    rc = L.mosaic_shard_plan(world, rank, grid_h, n_materials, *[C.byref(x) for x in v], C.byref(t[0]), C.byref(t[1]))
    if rc != 0:
        raise MosaicError(rc, "mosaic_shard_plan: invalid arguments")
    return dict(rows=(v[0].value, v[1].value), materials=(v[2].value, v[3].value), tri=(t[0].value, t[1].value))


class SimilarityDb:
    """The reference's similarity-database FILE (SimilarityDatabase, src/similarity.rs:31-138): JSON with
    path_to_index / index_to_path / lab_colors / similarities.  Host-only, no GPU needed."""

    def __init__(self, handle):
        self._L = load_library()
        self._h = handle

    @staticmethod
    def _paths(paths):
        arr = (C.c_char_p * len(paths))(*[p.encode("utf-8") for p in paths])
        return arr

    @classmethod
    def save(cls, file, paths, lab, sim):
        """SimilarityDatabase::save_to_file (similarity.rs:112-116)."""
        L = load_library()
        lab = np.ascontiguousarray(lab, np.float32).reshape(-1, 3)
        sim = np.ascontiguousarray(sim, np.float32).ravel()
        n = len(paths)
        assert lab.shape[0] == n and sim.shape[0] == (n * (n - 1) // 2 if n > 1 else 0)
        rc = L.mosaic_simdb_save_json(os.fsencode(file), cls._paths(paths), n,
                                      lab.ctypes.data_as(C.POINTER(C.c_float)), sim.ctypes.data_as(C.POINTER(C.c_float)))
        if rc != 0:
            raise MosaicError(rc, (L.mosaic_last_error(None) or b"").decode())

    @classmethod
    def load(cls, file):
        """SimilarityDatabase::load_from_file (similarity.rs:119-123)."""
        L = load_library()
        h = C.c_void_p()
        rc = L.mosaic_simdb_load_json(os.fsencode(file), C.byref(h))
        if rc != 0:
            raise MosaicError(rc, (L.mosaic_last_error(None) or b"").decode())
        return cls(h)

    def close(self):
        if getattr(self, "_h", None):
            self._L.mosaic_simdb_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        n, ns, npth = C.c_int(), C.c_uint64(), C.c_int()
        self._L.mosaic_simdb_info(self._h, C.byref(n), C.byref(ns), C.byref(npth))
        return dict(rows=n.value, similarities=ns.value, paths=npth.value)

    def path(self, index):
        p = self._L.mosaic_simdb_path(self._h, index)
        return None if p is None else p.decode("utf-8")

    def index_of(self, path):
        return int(self._L.mosaic_simdb_index_of(self._h, path.encode("utf-8")))

    def arrays(self):
        i = self.info()
        lab = np.empty((i["rows"], 3), np.float32)
        sim = np.empty(i["similarities"], np.float32)
        self._L.mosaic_simdb_copy(self._h, lab.ctypes.data_as(C.POINTER(C.c_float)), sim.ctypes.data_as(C.POINTER(C.c_float)))
        return lab, sim

    def get_similarity(self, path1, path2):
        """SimilarityDatabase::get_similarity (similarity.rs:85-103): None where the reference returns None."""
        out = C.c_float()
        rc = self._L.mosaic_simdb_get_similarity(self._h, path1.encode("utf-8"), path2.encode("utf-8"), C.byref(out))
        return out.value if rc == 0 else None


def _ptr(a):
    """numpy array / torch tensor / int address / None -> c_void_p."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):  # torch tensor (pinned host or device memory)
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


class Context:
    """One GPU's mosaic context (mosaic_ctx).  Mirrors the call sequence of the reference's
    MosaicGenerator::new + generate_mosaic (src/main.rs:105-164, 477-685)."""

    def __init__(self, device=0, rank=0, world=1, nccl_id: bytes | None = None):
        self._L = load_library()
        h = C.c_void_p()
        if world > 1:
            buf = C.create_string_buffer(nccl_id, 128)
            rc = self._L.mosaic_ctx_create_sharded(C.byref(h), device, rank, world, buf, 128)
        else:
            rc = self._L.mosaic_ctx_create(C.byref(h), device)
        if rc != 0:
            raise MosaicError(rc, (self._L.mosaic_last_error(None) or b"").decode())
        self._h = h
        self.rank, self.world = rank, world
        self._keep = []

    @staticmethod
    def nccl_unique_id() -> bytes:
        L = load_library()
        buf = C.create_string_buffer(128)
        rc = L.mosaic_nccl_unique_id(buf, 128)
        if rc != 0:
            raise MosaicError(rc, (L.mosaic_last_error(None) or b"").decode())
        return buf.raw

    def close(self):
        if getattr(self, "_h", None):
            self._L.mosaic_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise MosaicError(rc, (self._L.mosaic_last_error(self._h) or b"").decode())

    # ---- plumbing
    def set_stream(self, cuda_stream: int | None):
        self._check(self._L.mosaic_ctx_set_stream(self._h, C.c_void_p(cuda_stream) if cuda_stream else None))

    def sync(self):
        self._check(self._L.mosaic_ctx_sync(self._h))

    def stage_ms(self, stage) -> float:
        idx = STAGES.index(stage) if isinstance(stage, str) else int(stage)
        ms = C.c_float()
        self._check(self._L.mosaic_stage_ms(self._h, idx, C.byref(ms)))
        return ms.value

    def kernel_launches(self) -> int:
        return int(self._L.mosaic_kernel_launches(self._h))

    def collective_ms(self) -> float:
        """Device time spent in exchange steps (collectives, peer hand-off waits) since the previous call."""
        ms = C.c_float()
        self._check(self._L.mosaic_collective_ms(self._h, C.byref(ms)))
        return ms.value

    def sm_count(self) -> int:
        return int(self._L.mosaic_sm_count(self._h))

    def exchange_path(self) -> str:
        """"single", "peer" (NVLink peer windows) or "nccl: <reason>" — how a sharded mosaic_generate exchanges data."""
        return (self._L.mosaic_exchange_path(self._h) or b"").decode()

    def dims(self):
        v = [C.c_int() for _ in range(5)]
        self._check(self._L.mosaic_dims(self._h, *[C.byref(x) for x in v]))
        return dict(zip(("n", "grid_w", "grid_h", "tile_w", "tile_h"), (x.value for x in v)))

    def shard(self):
        v = [C.c_int() for _ in range(4)]
        t = [C.c_uint64(), C.c_uint64()]
        self._check(self._L.mosaic_shard(self._h, *[C.byref(x) for x in v], C.byref(t[0]), C.byref(t[1])))
        return dict(rows=(v[0].value, v[1].value), materials=(v[2].value, v[3].value), tri=(t[0].value, t[1].value))

    # ---- inputs
    def set_materials(self, mats):
        """mats: (n, h, w, 3) uint8 numpy array or pinned torch tensor (host)."""
        n, h, w, c = mats.shape
        assert c == 3
        self._check(self._L.mosaic_set_materials_rgb8(self._h, _ptr(mats), n, w, h))

    def set_materials_ragged(self, images):
        """Materials of individual sizes: a list of (h_i, w_i, 3) uint8 arrays (the reference opens files of any size)."""
        imgs = [np.ascontiguousarray(im, np.uint8) for im in images]
        assert all(im.ndim == 3 and im.shape[2] == 3 for im in imgs)
        sizes = np.array([im.size for im in imgs], np.uint64)
        offs = np.zeros(len(imgs), np.uint64)
        offs[1:] = np.cumsum(sizes)[:-1]
        blob = np.concatenate([im.ravel() for im in imgs])
        w = np.array([im.shape[1] for im in imgs], np.int32)
        h = np.array([im.shape[0] for im in imgs], np.int32)
        self._check(self._L.mosaic_set_materials_ragged_rgb8(
            self._h, blob.ctypes.data_as(C.POINTER(C.c_uint8)), offs.ctypes.data_as(C.POINTER(C.c_uint64)),
            w.ctypes.data_as(C.POINTER(C.c_int32)), h.ctypes.data_as(C.POINTER(C.c_int32)), len(imgs)))

    def set_materials_device(self, dev_tensor):
        n, h, w, c = dev_tensor.shape
        self._keep.append(dev_tensor)
        self._check(self._L.mosaic_set_materials_device(self._h, _ptr(dev_tensor), n, w, h))

    def set_target(self, img, grid_w, grid_h):
        h, w, c = img.shape
        assert c == 3
        self._check(self._L.mosaic_set_target_rgb8(self._h, _ptr(img), w, h, grid_w, grid_h))

    def set_target_device(self, dev_tensor, grid_w, grid_h):
        h, w, c = dev_tensor.shape
        self._keep.append(dev_tensor)
        self._check(self._L.mosaic_set_target_device(self._h, _ptr(dev_tensor), w, h, grid_w, grid_h))

    def set_material_lab(self, lab):
        lab = np.ascontiguousarray(lab, np.float32)
        self._check(self._L.mosaic_set_material_lab(self._h, lab.ctypes.data_as(C.POINTER(C.c_float)), lab.shape[0]))

    def set_similarity_lab(self, lab):
        """Lab colours of the similarity database (main.rs:120-143) when they differ from the tiles'; None clears."""
        if lab is None:
            self._check(self._L.mosaic_set_similarity_lab(self._h, None, 0))
            return
        lab = np.ascontiguousarray(lab, np.float32)
        self._check(self._L.mosaic_set_similarity_lab(self._h, lab.ctypes.data_as(C.POINTER(C.c_float)), lab.shape[0]))

    def set_cell_lab(self, lab, grid_w, grid_h):
        lab = np.ascontiguousarray(lab, np.float32)
        assert lab.shape[0] == grid_w * grid_h
        self._check(self._L.mosaic_set_cell_lab(self._h, lab.ctypes.data_as(C.POINTER(C.c_float)), grid_w, grid_h))

    def set_placement(self, placed):
        placed = np.ascontiguousarray(placed, np.int32).ravel()
        self._check(self._L.mosaic_set_placement(self._h, placed.ctypes.data_as(C.POINTER(C.c_int32))))

    # ---- stages
    def material_lab(self, fetch=True):
        out = np.empty((self.dims()["n"], 3), np.float32) if fetch else None
        self._check(self._L.mosaic_material_lab(self._h, _ptr(out)))
        return out

    def cell_lab(self, fetch=True):
        d = self.dims()
        out = np.empty((d["grid_w"] * d["grid_h"], 3), np.float32) if fetch else None
        self._check(self._L.mosaic_cell_lab(self._h, _ptr(out)))
        return out

    def similarity_tri(self, fetch=True):
        n = self.dims()["n"]
        out = np.empty(n * (n - 1) // 2 if n > 1 else 0, np.float32) if fetch else None
        self._check(self._L.mosaic_similarity_tri(self._h, _ptr(out)))
        return out

    def similarity_local(self, out=None):
        """This rank's range of the packed matrix, without the all-gather; `out` (host, tri_end - tri_begin floats) or None."""
        self._check(self._L.mosaic_similarity_local(self._h, _ptr(out)))
        return out

    def candidates(self, k=MAX_K, fetch=True):
        d = self.dims()
        kk = min(k if 0 < k <= MAX_K else MAX_K, d["n"])
        cells = d["grid_w"] * d["grid_h"]
        idx = np.empty((cells, kk), np.uint32) if fetch else None
        d2 = np.empty((cells, kk), np.float32) if fetch else None
        self._check(self._L.mosaic_candidates(self._h, k, _ptr(idx), _ptr(d2)))
        return idx, d2

    def place(self, max_usage=3, adjacency_weight=0.3, fetch=True):
        d = self.dims()
        placed = np.empty(d["grid_w"] * d["grid_h"], np.int32) if fetch else None
        nfb = C.c_uint32()
        self._check(self._L.mosaic_place(self._h, max_usage, adjacency_weight, _ptr(placed), C.byref(nfb)))
        return placed, nfb.value

    def fallback_cells(self, cap=4096):
        buf = np.zeros(cap, np.uint32)
        n = C.c_uint32()
        self._check(self._L.mosaic_fallback_cells(self._h, buf.ctypes.data_as(C.POINTER(C.c_uint32)), cap, C.byref(n)))
        return buf[:min(n.value, cap)].copy()

    def anneal(self, iterations=1000, seed=0, t0=100.0, decay=0.99995, fetch=True):
        d = self.dims()
        placed = np.empty(d["grid_w"] * d["grid_h"], np.int32) if fetch else None
        res = OptResult()
        self._check(self._L.mosaic_anneal(self._h, iterations, t0, decay, seed, _ptr(placed), C.byref(res)))
        return placed, res

    def optimize_greedy(self, max_iterations=1000, seed=0, fetch=True):
        """MosaicOptimizer::optimize_greedy (optimizer.rs:132-198) on the context's placement."""
        d = self.dims()
        placed = np.empty(d["grid_w"] * d["grid_h"], np.int32) if fetch else None
        res = OptResult()
        self._check(self._L.mosaic_optimize_greedy(self._h, max_iterations, seed, _ptr(placed), C.byref(res)))
        return placed, res

    def render(self, color_strength=0.3, out=None, fetch=True):
        d = self.dims()
        shape = (d["grid_h"] * d["tile_h"], d["grid_w"] * d["tile_w"], 3)
        if out is None and fetch:
            out = np.zeros(shape, np.uint8)
        self._check(self._L.mosaic_render(self._h, color_strength, _ptr(out)))
        return out

    def generate(self, max_usage=3, adjacency_weight=0.3, color_strength=0.3, enable_optimization=True,
                 optimization_iterations=1000, seed=0, build_similarity=False, placed_out=None, image_out=None):
        """Whole path in one call (generate_mosaic, main.rs:477-685).  placed_out / image_out: host buffers
        (numpy or pinned torch) or None to keep results on the device."""
        p = Params(max_usage, adjacency_weight, color_strength, 1 if enable_optimization else 0,
                   optimization_iterations, seed, 1 if build_similarity else 0)
        rep = Report()
        self._check(self._L.mosaic_generate(self._h, C.byref(p), _ptr(placed_out), _ptr(image_out), C.byref(rep)))
        return rep

    def save_similarity_db(self, file, paths):
        """Lab means + freshly built packed matrix -> the reference's JSON file (main.rs:122-131)."""
        assert len(paths) == self.dims()["n"]
        self._check(self._L.mosaic_simdb_save_ctx(self._h, os.fsencode(file), SimilarityDb._paths(paths)))

    def selftest(self):
        m = (C.c_uint64 * 2)()
        self._check(self._L.mosaic_selftest(self._h, m))
        return int(m[0]), int(m[1])

    def set_progress_callback(self, fn):
        """fn(fraction, stage_name, message) is called as the DEVICE passes the stages of generate() (on a thread of the
        CUDA runtime); None removes it."""
        if fn is None:
            self._check(self._L.mosaic_set_progress_callback(self._h, C.cast(None, PROGRESS_FN), None))
            self._progress_cb = None
            return

        def tramp(fraction, stage, message, _user):
            fn(float(fraction), STAGES[stage] if 0 <= stage < len(STAGES) else str(stage), message.decode("utf-8", "replace"))

        cb = PROGRESS_FN(tramp)
        self._check(self._L.mosaic_set_progress_callback(self._h, cb, None))
        self._progress_cb = cb  # keep the trampoline alive

    def set_chain(self, version):
        """0 = default placement chain, 1 = greedy.cu, 2 = the windowed chain of greedy_window.cu, 3 = the wave chain of greedy_wave.cu."""
        self._check(self._L.mosaic_debug_set_chain(self._h, int(version)))
        self._chain = int(version)

    def chain_profile(self, enable=True):
        """Enable/disable the instrumented placement kernel and return the counters of its last launch."""
        m = (C.c_uint64 * 16)()
        self._check(self._L.mosaic_debug_chain_profile(self._h, 1 if enable else 0, m))
        env = os.environ.get("MOSAIC_CHAIN", "")
        chain = getattr(self, "_chain", 0) or (3 if env.startswith("wa") else 2 if env.startswith("w") else 1)
        window = chain == 2
        if chain == 3:  # wave chain: [0..5] from thread 0, [8..11] summed over the other warps
            keys = ("sweeps", "n_serial_w0", "n_exact_w0", "cyc_phase1_w0", "cyc_barrier_a_w0", "cyc_phase2_w0", "n_eval_w0",
                    "cyc_barrier_b_w0", "n_eval", "n_exact", "n_serial", "cyc_phase1_others", "cyc_total", "cyc_phase2_others")
        elif not window:  # first-generation chain kernel
            keys = ("cyc_fast", "cyc_exact", "cyc_generic", "cyc_fallback", "n_fast", "n_exact", "n_generic",
                    "n_fallback", "cyc_wait_producer", "cyc_wait_committer", "cyc_total", "cyc_exit_to_full_step", "n_block_exits",
                    "cyc_full_step_uncontended", "n_full_step_uncontended", "n_hot_blocks")
        else:  # windowed chain: cycle counters are summed over its two warps
            keys = ("cyc_sets", "n_sets", "cyc_slow", "cyc_reset", "cyc_scan", "cyc_listen", "cyc_ring_wait", "n_fast",
                    "n_slow", "n_reset", "n_rescans", "cyc_total_both_warps")
        return dict(zip(keys, (int(x) for x in m)))

    def output_shape(self):
        d = self.dims()
        return (d["grid_h"] * d["tile_h"], d["grid_w"] * d["tile_w"], 3)


def _decode_with(fn_name, path) -> np.ndarray:
    L = load_library()
    buf = C.POINTER(C.c_uint8)()
    w, h = C.c_int(), C.c_int()
    rc = getattr(L, fn_name)(os.fsencode(path), C.byref(buf), C.byref(w), C.byref(h))
    if rc != 0:
        raise MosaicError(rc, (L.mosaic_last_error(None) or b"").decode())
    try:
        return np.ctypeslib.as_array(buf, shape=(h.value, w.value, 3)).copy()
    finally:
        L.mosaic_host_free(buf)


def jpeg_decode(path) -> np.ndarray:
    """image::open(path)?.to_rgb8() for a JPEG file: (h, w, 3) uint8, libjpeg-turbo's default arithmetic.  Host only."""
    return _decode_with("mosaic_jpeg_decode_rgb8", path)


def webp_decode(path) -> np.ndarray:
    """image::open(path)?.to_rgb8() for a WebP file (VP8 lossy with libwebp's default conversion to RGB, VP8L lossless).  Host only."""
    return _decode_with("mosaic_webp_decode_rgb8", path)


def webp_encode(path, img) -> None:
    """RgbImage::save(path) for a .webp path: lossless VP8L.  Host only."""
    L = load_library()
    img = np.ascontiguousarray(img, np.uint8)
    if img.ndim != 3 or img.shape[2] != 3:
        raise ValueError("webp_encode: (h, w, 3) uint8 expected")
    rc = L.mosaic_webp_encode_rgb8(os.fsencode(path), img.ctypes.data_as(C.c_void_p), img.shape[1], img.shape[0])
    if rc != 0:
        raise MosaicError(rc, (L.mosaic_last_error(None) or b"").decode())


def image_decode(path) -> np.ndarray:
    """PNG, JPEG or WebP by the file's first bytes.  Host only."""
    return _decode_with("mosaic_image_decode_rgb8", path)


def image_open(path) -> np.ndarray:
    """image::open(path)?.to_rgb8(): the format follows the path's EXTENSION (lower-cased), as in the `image` crate; the
    content is not sniffed.  Host only."""
    return _decode_with("mosaic_image_open_rgb8", path)


def png_decode(path) -> np.ndarray:
    """image::open(path)?.to_rgb8() for a PNG file (main.rs:263-301): (h, w, 3) uint8.  Host only."""
    L = load_library()
    buf = C.POINTER(C.c_uint8)()
    w, h = C.c_int(), C.c_int()
    rc = L.mosaic_png_decode_rgb8(os.fsencode(path), C.byref(buf), C.byref(w), C.byref(h))
    if rc != 0:
        raise MosaicError(rc, (L.mosaic_last_error(None) or b"").decode())
    try:
        return np.ctypeslib.as_array(buf, shape=(h.value, w.value, 3)).copy()
    finally:
        L.mosaic_host_free(buf)


def png_encode(path, img) -> None:
    """RgbImage::save(path) for a ".png" path (main.rs:668-672).  img: (h, w, 3) uint8.  Host only."""
    L = load_library()
    img = np.ascontiguousarray(img, np.uint8)
    assert img.ndim == 3 and img.shape[2] == 3
    rc = L.mosaic_png_encode_rgb8(os.fsencode(path), img.ctypes.data_as(C.c_void_p), img.shape[1], img.shape[0])
    if rc != 0:
        raise MosaicError(rc, (L.mosaic_last_error(None) or b"").decode())


class PinnedArray:
    """A numpy array over page-locked host memory of the library (mosaic_host_alloc): uploads and downloads through it are
    asynchronous.  `register(array)` page-locks memory numpy already owns instead (mosaic_host_register)."""

    def __init__(self, shape, dtype=np.uint8):
        L = load_library()
        self._L = L
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        rc = L.mosaic_host_alloc(nbytes, C.byref(p))
        if rc != 0:
            raise MosaicError(rc, (L.mosaic_last_error(None) or b"").decode())
        self._p = p
        buf = (C.c_uint8 * max(nbytes, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def close(self):
        if self._p:
            self.array = None
            self._L.mosaic_host_release(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class register:
    """Context manager: page-lock an existing C-contiguous numpy array for the duration of the block."""

    def __init__(self, array):
        assert array.flags["C_CONTIGUOUS"]
        self.array = array

    def __enter__(self):
        L = load_library()
        rc = L.mosaic_host_register(self.array.ctypes.data_as(C.c_void_p), self.array.nbytes)
        if rc != 0:
            raise MosaicError(rc, (L.mosaic_last_error(None) or b"").decode())
        return self.array

    def __exit__(self, *a):
        load_library().mosaic_host_unregister(self.array.ctypes.data_as(C.c_void_p))


def jpeg_encode(path, img, quality=75) -> None:
    """RgbImage::save(path) for a ".jpg" path (main.rs:668-672): baseline 4:4:4 at the given quality.  Host only."""
    L = load_library()
    img = np.ascontiguousarray(img, np.uint8)
    assert img.ndim == 3 and img.shape[2] == 3
    rc = L.mosaic_jpeg_encode_rgb8(os.fsencode(path), img.ctypes.data_as(C.c_void_p), img.shape[1], img.shape[0], int(quality))
    if rc != 0:
        raise MosaicError(rc, (L.mosaic_last_error(None) or b"").decode())
