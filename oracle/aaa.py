"""ctypes face of oracle/aaa.c -- TEST INFRASTRUCTURE ONLY: gg's CPU rasteriser (internal/raster AAA filler + SoftwareRenderer's
per-draw 8-bit source-over), the pixel oracle of SURVEY section 8 row a15. Only tests/ and bench.py's parity leg import this."""
import ctypes as C

import numpy as np

from . import twin


def _lib():
    L = twin.lib()
    L.oa_coverage.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]
    L.oa_coverage.restype = None
    L.oa_fill.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]
    L.oa_fill.restype = None
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def coverage(verbs, coords, w, h, even_odd=False, aa_shift=2, flatten=True, clip_margin=-1.0):
    """raster.FillToBuffer (analytic_filler.go:2587-2611): 8-bit coverage of one path. flatten / aa_shift as the reference's
    own tests pass them to NewEdgeBuilder / SetFlattenCurves; clip_margin >= 0 adds SoftwareRenderer.Fill's canvas clip."""
    v = np.ascontiguousarray(verbs, dtype=np.uint8)
    c = np.ascontiguousarray(coords, dtype=np.float64).ravel()
    out = np.zeros((h, w), dtype=np.uint8)
    _lib().oa_coverage(_p(v), v.size, _p(c), w, h, int(even_odd), aa_shift, int(flatten), float(clip_margin), _p(out))
    return out


class Pixmap:
    """gg.Pixmap + SoftwareRenderer.Fill for solid colours: premultiplied RGBA8, truncating source-over per draw."""

    def __init__(self, w, h):
        self.w, self.h = w, h
        self.data = np.zeros((h, w, 4), dtype=np.uint8)

    def fill(self, verbs, coords, rgba_straight, even_odd=False):
        v = np.ascontiguousarray(verbs, dtype=np.uint8)
        c = np.ascontiguousarray(coords, dtype=np.float64).ravel()
        col = np.ascontiguousarray(rgba_straight, dtype=np.float64)
        _lib().oa_fill(_p(self.data), self.w, self.h, _p(v), v.size, _p(c), _p(col), int(even_odd))
