"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs as the checker / CPU baseline.  The product
(mosaic-art-rust_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
# Flags of the two builds (oracle/Makefile).  "strict" is the parity checker; "fast" is the same source compiled the way
# BASELINE.md section 3 promises for the timed CPU arm (-O3 -march=native, still no FMA contraction / fast-math, so the
# results are identical — tests/test_oracle_golden.py checks that).  -march=native binds the binary to the host it was
# built on, so the fast library is built on the machine that runs it and carries a tag of that host's CPU in its name.
FLAGS = {"strict": "-O2 -std=c++17 -ffp-contract=off -fno-fast-math -fPIC -Wall -Wextra -pthread",
         "fast": "-O3 -march=native -std=c++17 -ffp-contract=off -fno-fast-math -fPIC -Wall -Wextra -pthread"}


def _host_tag() -> str:
    import hashlib
    sig = ""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith(("model name", "flags")):
                    sig += line
                if line.strip() == "" and sig:
                    break
    except OSError:
        pass
    return hashlib.sha1(sig.encode()).hexdigest()[:10]


def fast_lib_path() -> str:
    return os.path.join(_HERE, f"liboracle_fast_{_host_tag()}.so")


def _stale(target) -> bool:
    src = os.path.join(_HERE, "mosaic_oracle.cpp")
    hdr = os.path.join(_HERE, "mosaic_oracle.h")
    return (not os.path.exists(target)) or any(os.path.getmtime(p) > os.path.getmtime(target) for p in (src, hdr))


def build(force: bool = False) -> str:
    if force or _stale(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def build_fast(force: bool = False) -> str:
    """The timed CPU arm's build: -O3 -march=native on THIS host (bench.py only)."""
    target = fast_lib_path()
    if force or _stale(target):
        subprocess.check_call(["make", "-C", _HERE, "-B", "fast", f"FAST_LIB={os.path.basename(target)}"],
                              stdout=subprocess.DEVNULL)
    return target


class OptResult(C.Structure):
    _fields_ = [("initial_cost", C.c_float), ("final_cost", C.c_float), ("best_cost", C.c_float),
                ("improved_count", C.c_uint64), ("accepted_count", C.c_uint64), ("iterations", C.c_uint64)]


_lib = None
_libs = {}
_variant = "strict"


def select(variant: str) -> str:
    """Switch every wrapper of this module to the "strict" (parity, default) or the "fast" (timed CPU arm) build.
    Returns the compiler flags of the selected build."""
    global _lib, _variant
    assert variant in FLAGS
    _variant = variant
    _lib = _libs.get(variant)
    lib()
    return FLAGS[variant]


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build_fast() if _variant == "fast" else build())
        u8p, f32p, i32p, u32p, i16p = (C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_int32),
                                       C.POINTER(C.c_uint32), C.POINTER(C.c_int16))
        L.orc_rgb8_to_lab.argtypes = [C.c_uint8, C.c_uint8, C.c_uint8, f32p]
        L.orc_average_lab_region.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, f32p]
        L.orc_material_lab.argtypes = [u8p, C.c_int, C.c_int, C.c_int, f32p, C.c_int]
        L.orc_cell_lab.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, f32p, C.c_int]
        L.orc_average_rgb_region.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p]
        L.orc_topk.argtypes = [f32p, C.c_int, f32p, C.c_int, C.c_int, u32p, f32p, C.c_int]
        L.orc_topk_kdtree.argtypes = [f32p, C.c_int, f32p, C.c_int, C.c_int, u32p, f32p, C.c_int]
        L.orc_topk_kdtree.restype = None
        L.orc_lab_distance.argtypes = [f32p, f32p]
        L.orc_lab_distance.restype = C.c_float
        L.orc_similarity_tri.argtypes = [f32p, C.c_int, f32p, C.c_int]
        L.orc_tri_index.argtypes = [C.c_uint64] * 3
        L.orc_tri_index.restype = C.c_uint64
        L.orc_penalty.argtypes = [f32p, i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int32, C.c_float]
        L.orc_penalty.restype = C.c_float
        L.orc_total_cost.argtypes = [f32p, i32p, C.c_int, C.c_int]
        L.orc_total_cost.restype = C.c_float
        L.orc_swap_delta.argtypes = [f32p, i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_swap_delta.restype = C.c_float
        L.orc_place.argtypes = [u32p, f32p, C.c_int, f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                i32p, u32p, u32p, C.c_uint32]
        L.orc_place.restype = C.c_uint32
        L.orc_rng_new.argtypes = [C.c_uint64]
        L.orc_rng_new.restype = C.c_void_p
        L.orc_rng_free.argtypes = [C.c_void_p]
        L.orc_rng_next_u32.argtypes = [C.c_void_p]
        L.orc_rng_next_u32.restype = C.c_uint32
        L.orc_rng_next_u64.argtypes = [C.c_void_p]
        L.orc_rng_next_u64.restype = C.c_uint64
        L.orc_rng_gen_range.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_rng_gen_range.restype = C.c_uint64
        L.orc_rng_gen_f32.argtypes = [C.c_void_p]
        L.orc_rng_gen_f32.restype = C.c_float
        L.orc_chacha_block.argtypes = [u32p, C.c_uint64, C.c_int, u32p]
        L.orc_expf.argtypes = [C.c_float]
        L.orc_expf.restype = C.c_float
        L.orc_anneal.argtypes = [f32p, i32p, C.c_int, C.c_int, C.c_uint64, C.c_float, C.c_float, C.c_uint64,
                                 C.POINTER(OptResult)]
        L.orc_anneal_greedy.argtypes = [f32p, i32p, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.POINTER(OptResult)]
        L.orc_resize_window.argtypes = [C.c_int, C.c_int]
        L.orc_resize_window.restype = C.c_int
        L.orc_resize_coeffs.argtypes = [C.c_int, C.c_int, i32p, i32p, i16p, C.POINTER(C.c_int)]
        L.orc_resize_coeffs.restype = C.c_int
        L.orc_resize_rgb8.argtypes = [u8p, C.c_int, C.c_int, u8p, C.c_int, C.c_int]
        L.orc_color_adjustment_new.argtypes = [C.c_float] * 4 + [f32p]
        L.orc_optimal_adjustment.argtypes = [u8p, u8p, C.c_float, f32p]
        L.orc_adjust_pixel.argtypes = [f32p, u8p, u8p]
        L.orc_rgb_to_hsv.argtypes = [C.c_float] * 3 + [f32p]
        L.orc_hsv_to_rgb.argtypes = [C.c_float] * 3 + [f32p]
        L.orc_render.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_int, C.c_int, C.c_int,
                                 i32p, C.c_float, u8p, C.c_int]
        _lib = L
        _libs[_variant] = L
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def rgb8_to_lab(r, g, b):
    out = np.zeros(3, np.float32)
    lib().orc_rgb8_to_lab(r, g, b, _p(out, C.c_float))
    return out


def average_lab(img):
    """img: (h, w, 3) uint8 -> 3 floats (lib.rs:94)."""
    img = _c(img, np.uint8)
    h, w, _ = img.shape
    out = np.zeros(3, np.float32)
    lib().orc_average_lab_region(_p(img, C.c_uint8), w, 0, 0, w, h, _p(out, C.c_float))
    return out


def material_lab(mats, threads=1):
    mats = _c(mats, np.uint8)
    n, h, w, _ = mats.shape
    out = np.zeros((n, 3), np.float32)
    lib().orc_material_lab(_p(mats, C.c_uint8), n, w, h, _p(out, C.c_float), threads)
    return out


def cell_lab(target, gw, gh, threads=1):
    target = _c(target, np.uint8)
    h, w, _ = target.shape
    out = np.zeros((gw * gh, 3), np.float32)
    lib().orc_cell_lab(_p(target, C.c_uint8), w, h, gw, gh, _p(out, C.c_float), threads)
    return out


def average_rgb(img, x0=0, y0=0, rw=None, rh=None):
    img = _c(img, np.uint8)
    h, w, _ = img.shape
    rw = w if rw is None else rw
    rh = h if rh is None else rh
    out = np.zeros(3, np.uint8)
    lib().orc_average_rgb_region(_p(img, C.c_uint8), w, x0, y0, rw, rh, _p(out, C.c_uint8))
    return out


def topk(mat_lab, cell_lab_, k=100, threads=1):
    mat_lab = _c(mat_lab, np.float32)
    cell_lab_ = _c(cell_lab_, np.float32)
    n, c = mat_lab.shape[0], cell_lab_.shape[0]
    k = min(k, n)
    idx = np.zeros((c, k), np.uint32)
    d2 = np.zeros((c, k), np.float32)
    lib().orc_topk(_p(mat_lab, C.c_float), n, _p(cell_lab_, C.c_float), c, k, _p(idx, C.c_uint32),
                   _p(d2, C.c_float), threads)
    return idx, d2


def topk_kdtree(mat_lab, cell_lab_, k=100, threads=1):
    """orc_topk's result through a bucketed k-d tree (O(log n + k) per query, like the reference's kiddo search)."""
    mat_lab = _c(mat_lab, np.float32)
    cell_lab_ = _c(cell_lab_, np.float32)
    n, c = mat_lab.shape[0], cell_lab_.shape[0]
    k = min(k, n)
    idx = np.zeros((c, k), np.uint32)
    d2 = np.zeros((c, k), np.float32)
    lib().orc_topk_kdtree(_p(mat_lab, C.c_float), n, _p(cell_lab_, C.c_float), c, k, _p(idx, C.c_uint32),
                          _p(d2, C.c_float), threads)
    return idx, d2


def lab_distance(a, b):
    a = _c(a, np.float32)
    b = _c(b, np.float32)
    return float(lib().orc_lab_distance(_p(a, C.c_float), _p(b, C.c_float)))


def tri_index(i, j, n):
    return int(lib().orc_tri_index(i, j, n))


def similarity_tri(lab, threads=1):
    lab = _c(lab, np.float32)
    n = lab.shape[0]
    out = np.zeros(n * (n - 1) // 2 if n > 0 else 0, np.float32)
    if n > 1:
        lib().orc_similarity_tri(_p(lab, C.c_float), n, _p(out, C.c_float), threads)
    return out


def penalty(mat_lab, grid, x, y, candidate, weight):
    mat_lab = _c(mat_lab, np.float32)
    grid = _c(grid, np.int32)
    gh, gw = grid.shape
    return float(lib().orc_penalty(_p(mat_lab, C.c_float), _p(grid, C.c_int32), gw, gh, x, y, candidate, weight))


def total_cost(mat_lab, grid):
    mat_lab = _c(mat_lab, np.float32)
    grid = _c(grid, np.int32)
    gh, gw = grid.shape
    return np.float32(lib().orc_total_cost(_p(mat_lab, C.c_float), _p(grid, C.c_int32), gw, gh))


def swap_delta(mat_lab, grid, p1, p2):
    mat_lab = _c(mat_lab, np.float32)
    grid = _c(grid, np.int32)
    gh, gw = grid.shape
    return np.float32(lib().orc_swap_delta(_p(mat_lab, C.c_float), _p(grid, C.c_int32), gw, gh,
                                           p1[0], p1[1], p2[0], p2[1]))


def place(cand_idx, cand_d2, mat_lab, gw, gh, max_usage, weight, fallback_cap=4096):
    cand_idx = _c(cand_idx, np.uint32)
    cand_d2 = _c(cand_d2, np.float32)
    mat_lab = _c(mat_lab, np.float32)
    c, k = cand_idx.shape
    assert c == gw * gh
    n = mat_lab.shape[0]
    placed = np.zeros(c, np.int32)
    usage = np.zeros(n, np.uint32)
    fb = np.zeros(fallback_cap, np.uint32)
    nfb = lib().orc_place(_p(cand_idx, C.c_uint32), _p(cand_d2, C.c_float), k, _p(mat_lab, C.c_float), n, gw, gh,
                          max_usage, weight, _p(placed, C.c_int32), _p(usage, C.c_uint32), _p(fb, C.c_uint32),
                          fallback_cap)
    return placed, int(nfb), fb[:min(nfb, fallback_cap)].copy(), usage


class Rng:
    def __init__(self, seed):
        self._h = lib().orc_rng_new(seed)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_rng_free(self._h)
            self._h = None

    def next_u32(self):
        return int(lib().orc_rng_next_u32(self._h))

    def next_u64(self):
        return int(lib().orc_rng_next_u64(self._h))

    def gen_range(self, n):
        return int(lib().orc_rng_gen_range(self._h, n))

    def gen_f32(self):
        return np.float32(lib().orc_rng_gen_f32(self._h))


def chacha_block(key, counter, rounds):
    key = _c(key, np.uint32)
    out = np.zeros(16, np.uint32)
    lib().orc_chacha_block(_p(key, C.c_uint32), counter, rounds, _p(out, C.c_uint32))
    return out


def expf(x):
    return np.float32(lib().orc_expf(C.c_float(float(x))))


def anneal(mat_lab, grid, iterations, seed, t0=100.0, decay=0.99995):
    """grid: (gh, gw) int32; returns (new grid, OptResult)."""
    mat_lab = _c(mat_lab, np.float32)
    g = _c(grid, np.int32).copy()
    gh, gw = g.shape if g.ndim == 2 else (0, 0)
    res = OptResult()
    lib().orc_anneal(_p(mat_lab, C.c_float), _p(g, C.c_int32), gw, gh, iterations, t0, decay, seed, C.byref(res))
    return g, res


def anneal_greedy(mat_lab, grid, iterations, seed):
    """MosaicOptimizer::optimize_greedy (optimizer.rs:132-198). grid: (gh, gw) int32; returns (new grid, OptResult)."""
    mat_lab = _c(mat_lab, np.float32)
    g = _c(grid, np.int32).copy()
    gh, gw = g.shape if g.ndim == 2 else (0, 0)
    res = OptResult()
    lib().orc_anneal_greedy(_p(mat_lab, C.c_float), _p(g, C.c_int32), gw, gh, iterations, seed, C.byref(res))
    return g, res


def resize_coeffs(in_size, out_size):
    win = lib().orc_resize_window(in_size, out_size)
    start = np.zeros(out_size, np.int32)
    size = np.zeros(out_size, np.int32)
    coeffs = np.zeros((out_size, win), np.int16)
    prec = C.c_int(0)
    lib().orc_resize_coeffs(in_size, out_size, _p(start, C.c_int32), _p(size, C.c_int32), _p(coeffs, C.c_int16),
                            C.byref(prec))
    return start, size, coeffs, prec.value


def resize_rgb8(src, dw, dh):
    src = _c(src, np.uint8)
    sh, sw, _ = src.shape
    dst = np.zeros((dh, dw, 3), np.uint8)
    lib().orc_resize_rgb8(_p(src, C.c_uint8), sw, sh, _p(dst, C.c_uint8), dw, dh)
    return dst


def color_adjustment_new(brightness, contrast, hue_shift, saturation):
    out = np.zeros(4, np.float32)
    lib().orc_color_adjustment_new(brightness, contrast, hue_shift, saturation, _p(out, C.c_float))
    return out


def optimal_adjustment(tile_avg, target_avg, strength):
    t = _c(tile_avg, np.uint8)
    g = _c(target_avg, np.uint8)
    out = np.zeros(4, np.float32)
    lib().orc_optimal_adjustment(_p(t, C.c_uint8), _p(g, C.c_uint8), strength, _p(out, C.c_float))
    return out


def adjust_pixel(adj, px):
    adj = _c(adj, np.float32)
    px = _c(px, np.uint8)
    out = np.zeros(3, np.uint8)
    lib().orc_adjust_pixel(_p(adj, C.c_float), _p(px, C.c_uint8), _p(out, C.c_uint8))
    return out


def rgb_to_hsv(r, g, b):
    out = np.zeros(3, np.float32)
    lib().orc_rgb_to_hsv(r, g, b, _p(out, C.c_float))
    return out


def hsv_to_rgb(h, s, v):
    out = np.zeros(3, np.float32)
    lib().orc_hsv_to_rgb(h, s, v, _p(out, C.c_float))
    return out


def render(target, gw, gh, materials, placed, strength, threads=1):
    target = _c(target, np.uint8)
    materials = _c(materials, np.uint8)
    placed = _c(placed, np.int32)
    H, W, _ = target.shape
    n, mh, mw, _ = materials.shape
    tw, th = W // gw, H // gh
    out = np.zeros((gh * th, gw * tw, 3), np.uint8)
    lib().orc_render(_p(target, C.c_uint8), W, H, gw, gh, _p(materials, C.c_uint8), n, mw, mh,
                     _p(placed, C.c_int32), strength, _p(out, C.c_uint8), threads)
    return out
