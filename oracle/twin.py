"""ctypes face of oracle/libggoracle.so -- TEST INFRASTRUCTURE ONLY.

The CPU twin oracle (oracle/twin.c) restates gogpu/gg internal/gpu/tilecompute. Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; nothing under gg_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

LINE = np.dtype([("path_ix", "<u4"), ("p0", "<f4", 2), ("p1", "<f4", 2)])           # types.go:13-19
PATH = np.dtype([("bbox", "<u4", 4), ("tiles", "<u4")])                              # types.go:21-25
TILE = np.dtype([("backdrop", "<i4"), ("seg_count_or_ix", "<u4")])                   # types.go:27-31
SEGCOUNT = np.dtype([("line_ix", "<u4"), ("counts", "<u4")])                         # types.go:33-38
SEGMENT = np.dtype([("p0", "<f4", 2), ("p1", "<f4", 2), ("y_edge", "<f4")])          # types.go:40-46
PATH_MONOID = np.dtype([(n, "<u4") for n in ("trans_ix", "path_seg_ix", "path_seg_offset", "style_ix", "path_ix")])
DRAW_MONOID = np.dtype([(n, "<u4") for n in ("path_ix", "clip_ix", "scene_offset", "info_offset")])
ELEMENT = np.dtype([("type", "<u4"), ("line_start", "<u4"), ("line_count", "<u4"), ("color", "u1", 4),
                    ("even_odd", "<u4"), ("blend", "<u4"), ("alpha", "<f4"), ("packed_rgba", "<u4")])
ELEM_DRAW, ELEM_BEGIN_CLIP, ELEM_END_CLIP = 0, 1, 2


class _Layout(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("n_draw_objects", "n_paths", "n_clips", "path_tag_base", "path_data_base",
                                          "draw_tag_base", "draw_data_base", "transform_base", "style_base")]


class _Coarse(C.Structure):
    _fields_ = [("width_in_tiles", C.c_int), ("height_in_tiles", C.c_int),
                ("n_paths", C.c_uint32), ("paths", C.c_void_p),
                ("n_tiles", C.c_uint32), ("tiles", C.c_void_p),
                ("n_segments", C.c_uint32), ("segments", C.c_void_p),
                ("path_seg_base", C.c_void_p), ("path_total_segs", C.c_void_p),
                ("ptcl_offsets", C.c_void_p), ("ptcl_words", C.c_void_p),
                ("n_scene_words", C.c_uint32), ("scene", C.c_void_p), ("layout", _Layout),
                ("n_tag_words", C.c_uint32), ("tag_monoids", C.c_void_p),
                ("draw_monoids", C.c_void_p),
                ("n_info", C.c_uint32), ("info", C.c_void_p),
                ("n_gtab_words", C.c_uint32), ("gtab", C.c_void_p)]


def build(force=False):
    so = os.path.join(_HERE, "libggoracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.ot_flatten_fill.restype = C.c_uint32
        L.ot_flatten_fill.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        L.ot_flatten_path.restype = C.c_uint32
        L.ot_flatten_path.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_int, C.c_void_p, C.c_uint32]
        L.ot_path_monoid_new.argtypes = [C.c_uint32, C.c_void_p]
        L.ot_draw_monoid_new.argtypes = [C.c_uint32, C.c_void_p]
        L.ot_rasterize.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.ot_rasterize_scene.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.ot_coarse_run.restype = C.POINTER(_Coarse)
        L.ot_coarse_run.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_int, C.c_int]
        L.ot_coarse_free.argtypes = [C.POINTER(_Coarse)]
        L.ot_fine_tile.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.ot_fine_frame.argtypes = [C.POINTER(_Coarse), C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.ot_path_count.restype = C.c_uint32
        L.ot_path_count.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ot_path_tiling.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ot_line_bbox.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_void_p]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def flatten_fill(cubics):
    """flatten.go:32 FlattenFill. cubics: (n, 8) float32 -> LINE array."""
    cubics = np.ascontiguousarray(cubics, dtype=np.float32).reshape(-1, 8)
    cap = max(64, 128 * len(cubics))
    while True:
        out = np.zeros(cap, dtype=LINE)
        n = lib().ot_flatten_fill(_p(cubics), len(cubics), _p(out), cap)
        if n <= cap:
            return out[:n].copy()
        cap = n


def flatten_path(verbs, coords, auto_close=False):
    """path_convert.go:29 convertPathToPathDef (geometry). verbs u8, coords f64."""
    verbs = np.ascontiguousarray(verbs, dtype=np.uint8)
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    cap = max(64, 64 * len(verbs))
    while True:
        out = np.zeros(cap, dtype=LINE)
        n = lib().ot_flatten_path(_p(verbs), len(verbs), _p(coords), int(auto_close), _p(out), cap)
        if n <= cap:
            return out[:n].copy()
        cap = n


def path_monoid(word):
    m = np.zeros(1, dtype=PATH_MONOID)
    lib().ot_path_monoid_new(int(word) & 0xFFFFFFFF, _p(m))
    return m[0]


def draw_monoid(tag):
    m = np.zeros(1, dtype=DRAW_MONOID)
    lib().ot_draw_monoid_new(int(tag) & 0xFFFFFFFF, _p(m))
    return m[0]


def rasterize(lines, even_odd, w, h):
    lines = np.ascontiguousarray(lines, dtype=LINE)
    out = np.zeros((h, w), dtype=np.float32)
    lib().ot_rasterize(_p(lines), len(lines), int(even_odd), w, h, _p(out))
    return out


def make_elements(elems):
    """elems: list of dicts {type, lines(LINE array)|None, color, even_odd, blend, alpha}."""
    e = np.zeros(len(elems), dtype=ELEMENT)
    chunks, k = [], 0
    for i, d in enumerate(elems):
        ln = d.get("lines")
        n = 0 if ln is None else len(ln)
        e[i]["type"] = d.get("type", ELEM_DRAW)
        e[i]["line_start"] = k
        e[i]["line_count"] = n
        e[i]["color"] = d.get("color", (0, 0, 0, 255))
        e[i]["even_odd"] = int(d.get("even_odd", 0))
        e[i]["blend"] = d.get("blend", 0)
        e[i]["alpha"] = d.get("alpha", 1.0)
        if n:
            chunks.append(np.ascontiguousarray(ln, dtype=LINE))
            k += n
    lines = np.concatenate(chunks) if chunks else np.zeros(0, dtype=LINE)
    return e, lines


def rasterize_scene(bg, elems, lines, w, h):
    """rasterizer.go:176 RasterizeScene -> (h, w, 4) straight RGBA8."""
    bg = np.asarray(bg, dtype=np.uint8)
    out = np.zeros((h, w, 4), dtype=np.uint8)
    lib().ot_rasterize_scene(_p(bg), _p(elems), len(elems), _p(lines), w, h, _p(out))
    return out


def _view(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n).copy()


class Coarse:
    """Copy-out of ot_coarse (coarse.go:17-43 CoarseOutput + scan results)."""

    def __init__(self, elems, lines, w, h, _handle=None):
        L = lib()
        if _handle is None:
            elems = np.ascontiguousarray(elems, dtype=ELEMENT)
            lines = np.ascontiguousarray(lines, dtype=LINE)
            _handle = L.ot_coarse_run(_p(elems), len(elems), _p(lines), w, h)
        self._c = _handle
        c = self._c.contents
        self.w, self.h = w, h
        self.wt, self.ht = c.width_in_tiles, c.height_in_tiles
        self.paths = _view(c.paths, c.n_paths, PATH)
        self.tiles = _view(c.tiles, c.n_tiles, TILE)
        self.segments = _view(c.segments, c.n_segments, SEGMENT)
        self.path_seg_base = _view(c.path_seg_base, c.n_paths, np.uint32)
        self.path_total_segs = _view(c.path_total_segs, c.n_paths, np.uint32)
        ng = self.wt * self.ht
        self.ptcl_offsets = _view(c.ptcl_offsets, ng + 1, np.uint32)
        self.ptcl_words = _view(c.ptcl_words, int(self.ptcl_offsets[-1]) if ng >= 0 else 0, np.uint32)
        self.scene = _view(c.scene, c.n_scene_words, np.uint32)
        self.layout = {n: getattr(c.layout, n) for n, _ in _Layout._fields_}
        self.tag_monoids = _view(c.tag_monoids, c.n_tag_words, PATH_MONOID)
        self.draw_monoids = _view(c.draw_monoids, c.layout.n_draw_objects, DRAW_MONOID)
        self.info = _view(c.info, c.n_info, np.uint32)

    @classmethod
    def from_packed(cls, scene_words, layout, w, h):
        """CPU pipeline up to PTCL from ggcuda's packed scene (oracle/packed.c)."""
        L = lib()
        L.ot_coarse_from_packed.restype = C.POINTER(_Coarse)
        L.ot_coarse_from_packed.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        sw = np.ascontiguousarray(scene_words, dtype=np.uint32)
        l13 = _layout13(layout)
        return cls(None, None, w, h, _handle=L.ot_coarse_from_packed(_p(sw), _p(l13), w, h))

    def ptcl(self, tile_ix):
        return self.ptcl_words[self.ptcl_offsets[tile_ix]:self.ptcl_offsets[tile_ix + 1]]

    def fine(self, bg, straight=True, premul=True):
        bg = np.asarray(bg, dtype=np.uint8)
        o_s = np.zeros((self.h, self.w, 4), dtype=np.uint8) if straight else None
        o_p = np.zeros((self.h, self.w, 4), dtype=np.uint8) if premul else None
        lib().ot_fine_frame(self._c, _p(bg), self.w, self.h,
                            _p(o_s) if straight else None, _p(o_p) if premul else None)
        return o_s, o_p

    def __del__(self):
        try:
            if self._c:
                lib().ot_coarse_free(self._c)
                self._c = None
        except Exception:
            pass


def fine_tile(ptcl_words, segments, bg_premul_f32):
    """fine.go:40 fineRasterizeTile -> (256, 4) premultiplied float32."""
    w = np.ascontiguousarray(ptcl_words, dtype=np.uint32)
    s = np.ascontiguousarray(segments, dtype=SEGMENT)
    bg = np.asarray(bg_premul_f32, dtype=np.float32)
    out = np.zeros((256, 4), dtype=np.float32)
    lib().ot_fine_tile(_p(w), len(w), _p(s) if len(s) else None, len(s), _p(bg), _p(out))
    return out


# ---- scene helpers shared by the reference's golden tests (rasterizer_test.go:139-171) ----
def circle_cubics(cx, cy, r):
    f = np.float32
    cx, cy, r = f(cx), f(cy), f(r)
    k = f(r * f(0.5522847498))
    return np.array([
        [cx + r, cy, cx + r, cy + k, cx + k, cy + r, cx, cy + r],
        [cx, cy + r, cx - k, cy + r, cx - r, cy + k, cx - r, cy],
        [cx - r, cy, cx - r, cy - k, cx - k, cy - r, cx, cy - r],
        [cx, cy - r, cx + k, cy - r, cx + r, cy - k, cx + r, cy],
    ], dtype=np.float32)


def polygon_lines(verts):
    verts = np.asarray(verts, dtype=np.float32)
    out = []
    n = len(verts)
    for i in range(n):
        p0, p1 = verts[i], verts[(i + 1) % n]
        if p0[0] == p1[0] and p0[1] == p1[1]:
            continue
        out.append((0, p0, p1))
    return np.array(out, dtype=LINE)


# ---- ggcuda packed-scene entry points (oracle/packed.c) ----
class _Timing(C.Structure):
    _fields_ = [("t_flatten", C.c_double), ("t_coarse", C.c_double), ("t_fine", C.c_double),
                ("n_lines", C.c_uint32), ("n_segments", C.c_uint32), ("n_ptcl_words", C.c_uint32)]


def _layout13(layout):
    """layout: numpy record with gg_b200._lib.LAYOUT fields, or a 13-sequence."""
    if hasattr(layout, "dtype") and layout.dtype.names:
        return np.array([int(layout[n]) for n in layout.dtype.names], dtype=np.uint32)
    return np.asarray(layout, dtype=np.uint32)


def scan_stages(scene_words, layout):
    """pathtag scan, draw scan + leaf, clip leaf (pathtag.go:76-121, draw_leaf.go:54-151, clip_leaf.go:27-56) on ggcuda's
    packed scene -> (tag_monoids, draw_monoids after the clip-leaf fix-up, info, clip_inps)."""
    L = lib()
    L.ot_scan_stages.restype = C.c_uint32
    L.ot_scan_stages.argtypes = [C.c_void_p] + [C.c_uint32] * 7 + [C.c_void_p] * 4
    sw = np.ascontiguousarray(scene_words, dtype=np.uint32)
    g = {n: int(layout[n]) for n in layout.dtype.names}
    tm = np.zeros(g["n_tag_words"], dtype=PATH_MONOID)
    dm = np.zeros(max(1, g["n_draws"]), dtype=DRAW_MONOID)
    info = np.zeros(max(1, g["n_draws"]), dtype=np.uint32)
    ci = np.zeros((max(1, g["n_clips"]), 2), dtype=np.int32)
    n_info = L.ot_scan_stages(_p(sw), g["n_scene_words"], g["path_tag_base"], g["n_tag_words"], g["draw_tag_base"], g["draw_data_base"],
                              g["n_draws"], g["n_clips"], _p(tm), _p(dm), _p(info), _p(ci))
    return tm, dm[:g["n_draws"]], info[:n_info], ci[:g["n_clips"]]


def flatten_packed(scene_words, layout):
    L = lib()
    L.ot_flatten_packed.restype = C.c_uint32
    L.ot_flatten_packed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    sw = np.ascontiguousarray(scene_words, dtype=np.uint32)
    l13 = _layout13(layout)
    cap = 1 << 16
    while True:
        out = np.zeros(cap, dtype=LINE)
        n = L.ot_flatten_packed(_p(sw), _p(l13), _p(out), cap)
        if n <= cap:
            return out[:n].copy()
        cap = n


def render_packed(scene_words, layout, w, h, bg_premul=(0, 0, 0, 0), threads=1):
    """Whole CPU pipeline from ggcuda's packed scene -> ((h, w, 4) premultiplied RGBA8, timing dict)."""
    L = lib()
    L.ot_render_packed.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(_Timing)]
    sw = np.ascontiguousarray(scene_words, dtype=np.uint32)
    l13 = _layout13(layout)
    bg = np.asarray(bg_premul, dtype=np.uint8)
    out = np.zeros((h, w, 4), dtype=np.uint8)
    tm = _Timing()
    L.ot_render_packed(_p(sw), _p(l13), w, h, _p(bg), int(threads), _p(out), C.byref(tm))
    return out, {n: getattr(tm, n) for n, _ in _Timing._fields_}
