"""ctypes binding of libggcuda.so (include/ggcuda.h). No CPU fallback: if the library is
missing or no B200 is present, every entry point raises."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libggcuda.so")

OK, ERR_CUDA, ERR_INVALID, ERR_UNSUPPORTED, ERR_NOMEM = 0, -1, -2, -3, -4
COMPOSITE_OVER, KEEP_SCENE, TARGET_F32, NO_WAIT = 1, 2, 4, 8
(BUF_SCENE, BUF_TAG_MONOIDS, BUF_DRAW_MONOIDS, BUF_INFO, BUF_CLIP_INPS, BUF_LINES, BUF_PATHS, BUF_TILES,
 BUF_SEG_START, BUF_SEGMENTS, BUF_PTCL_OFF, BUF_PTCL, BUF_HIT_CNT, BUF_LAYOUT, BUF_RESTART) = range(15)

# every symbol include/ggcuda.h declares (tests check the .so exports exactly these)
SYMBOLS = [
    "ggcuda_create", "ggcuda_destroy", "ggcuda_last_error", "ggcuda_set_stream", "ggcuda_begin", "ggcuda_set_background",
    "ggcuda_set_band", "ggcuda_fill_path", "ggcuda_stroke_path", "ggcuda_push_clip", "ggcuda_push_layer", "ggcuda_pop",
    "ggcuda_add_encoding", "ggcuda_flush", "ggcuda_upload", "ggcuda_render_device", "ggcuda_render_device_multi", "ggcuda_get_stats", "ggcuda_set_timing",
    "ggcuda_debug_read", "ggcuda_pack_host", "ggcuda_begin_keyed", "ggcuda_set_dirty_rect", "ggcuda_register_target",
    "ggcuda_unregister_target", "ggcuda_comm_unique_id", "ggcuda_comm_init", "ggcuda_comm_destroy", "ggcuda_all_gather_bands",
    "ggcuda_sync", "ggcuda_encoding_hash", "ggcuda_fill_path_gradient", "ggcuda_add_image", "ggcuda_broadcast_band",
]

LINE = np.dtype([("path_ix", "<u4"), ("p0", "<f4", 2), ("p1", "<f4", 2)])
PATH = np.dtype([("bbox", "<u4", 4), ("tiles", "<u4")])
TILE = np.dtype([("backdrop", "<i4"), ("seg_count", "<u4")])
SEGMENT = np.dtype([("p0", "<f4", 2), ("p1", "<f4", 2), ("y_edge", "<f4")])
PATH_MONOID = np.dtype([(n, "<u4") for n in ("trans_ix", "path_seg_ix", "path_seg_offset", "style_ix", "path_ix")])
DRAW_MONOID = np.dtype([(n, "<u4") for n in ("path_ix", "clip_ix", "scene_offset", "info_offset")])
CLIP_INP = np.dtype([("ix", "<u4"), ("path_ix", "<i4")])
LAYOUT = np.dtype([(n, "<u4") for n in ("n_tag_bytes", "n_tag_words", "n_draws", "n_paths", "n_clips", "path_tag_base",
                                         "path_data_base", "draw_tag_base", "draw_data_base", "transform_base", "style_base",
                                         "clip_aux_base", "n_scene_words")])


CREATE_HOST_STROKES = 1   # ggcuda_create flag (diagnostic host polyline stroker)


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("n_draws", "n_paths", "n_clips", "n_tag_bytes", "n_lines", "n_path_tiles",
                                          "n_seg_counts", "n_segments", "n_hits", "n_ptcl_words", "n_spill", "passes",
                                          "kernel_launches")] + \
               [("scene_bytes", C.c_uint64), ("device_bytes", C.c_uint64)] + \
               [(n, C.c_float) for n in ("ms_front", "ms_binning", "ms_coarse", "ms_fine")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class GGCudaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"ggcuda error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """dlopen libggcuda.so and declare prototypes. Raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -m gg_b200.build` (there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, u32, sz = C.c_void_p, C.c_uint32, C.c_size_t
    L.ggcuda_create.argtypes = [C.c_int, u32, C.POINTER(vp)]
    L.ggcuda_destroy.argtypes = [vp]
    L.ggcuda_destroy.restype = None
    L.ggcuda_last_error.argtypes = [vp]
    L.ggcuda_last_error.restype = C.c_char_p
    L.ggcuda_set_stream.argtypes = [vp, vp]
    L.ggcuda_begin.argtypes = [vp, u32, u32]
    L.ggcuda_set_background.argtypes = [vp, vp]
    L.ggcuda_set_band.argtypes = [vp, u32, u32]
    L.ggcuda_fill_path.argtypes = [vp, vp, u32, vp, u32, vp, C.c_int]
    L.ggcuda_stroke_path.argtypes = [vp, vp, u32, vp, u32, vp, C.c_double, C.c_int, C.c_int, C.c_double]
    L.ggcuda_fill_path_gradient.argtypes = [vp, vp, u32, vp, u32, C.c_int, vp, vp, u32, C.c_int, C.c_int]
    L.ggcuda_add_image.argtypes = [vp, u32, u32, vp, vp]
    L.ggcuda_broadcast_band.argtypes = [vp, vp, vp, u32, C.c_int, C.c_size_t, vp]
    L.ggcuda_push_clip.argtypes = [vp, vp, u32, vp, u32]
    L.ggcuda_push_layer.argtypes = [vp, u32, C.c_float]
    L.ggcuda_pop.argtypes = [vp]
    L.ggcuda_add_encoding.argtypes = [vp, vp, sz, vp, sz, vp, sz, vp, sz, vp, sz]
    L.ggcuda_flush.argtypes = [vp, vp, sz, u32]
    L.ggcuda_upload.argtypes = [vp]
    L.ggcuda_render_device.argtypes = [vp, vp, sz, u32]
    L.ggcuda_render_device_multi.argtypes = [vp, vp, vp, u32, C.c_int, sz, u32]
    L.ggcuda_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.ggcuda_set_timing.argtypes = [vp, C.c_int]
    L.ggcuda_debug_read.argtypes = [vp, C.c_int, vp, sz]
    L.ggcuda_debug_read.restype = C.c_longlong
    L.ggcuda_pack_host.argtypes = [vp, vp, sz, vp]
    L.ggcuda_pack_host.restype = C.c_longlong
    L.ggcuda_begin_keyed.argtypes = [vp, u32, u32, C.c_uint64, C.POINTER(C.c_int)]
    L.ggcuda_set_dirty_rect.argtypes = [vp, u32, u32, u32, u32]
    L.ggcuda_register_target.argtypes = [vp, vp, sz]
    L.ggcuda_unregister_target.argtypes = [vp, vp]
    L.ggcuda_comm_unique_id.argtypes = [vp]
    L.ggcuda_comm_init.argtypes = [vp, C.c_int, C.c_int, vp]
    L.ggcuda_comm_destroy.argtypes = [vp]
    L.ggcuda_all_gather_bands.argtypes = [vp, vp, sz]
    L.ggcuda_sync.argtypes = [vp]
    L.ggcuda_encoding_hash.argtypes = [vp, sz, vp, sz, vp, sz, vp, sz, vp, sz]
    L.ggcuda_encoding_hash.restype = C.c_uint64
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


def encoding_hash(tags, path_data, draw_data, transforms, brushes=None):
    """scene.Encoding.Hash (scene/encoding.go:752-802); with `brushes` the colours are folded in behind it."""
    t = np.ascontiguousarray(tags, dtype=np.uint8)
    pd = np.ascontiguousarray(path_data, dtype=np.float32)
    dd = np.ascontiguousarray(draw_data, dtype=np.uint32)
    tr = np.ascontiguousarray(transforms, dtype=np.float32).ravel()
    br = None if brushes is None else np.ascontiguousarray(brushes, dtype=np.float64).ravel()
    return int(load().ggcuda_encoding_hash(_p(t), t.size, _p(pd), pd.size, _p(dd), dd.size, _p(tr), tr.size,
                                           _p(br) if br is not None else None, 0 if br is None else br.size // 4))


def comm_unique_id():
    """128 bytes identifying a new NCCL communicator (one rank calls this and hands them to the others)."""
    u = np.zeros(128, dtype=np.uint8)
    rc = load().ggcuda_comm_unique_id(_p(u))
    if rc != 0:
        raise GGCudaError(rc, (load().ggcuda_last_error(None) or b"").decode())
    return u.tobytes()


class Context:
    """Thin object wrapper over a ggcuda_ctx."""

    def __init__(self, device=0, flags=0):
        self.L = load()
        h = C.c_void_p()
        rc = self.L.ggcuda_create(device, flags, C.byref(h))
        if rc != 0:
            raise GGCudaError(rc, (self.L.ggcuda_last_error(None) or b"").decode())
        self.h = h

    def _ck(self, rc):
        if rc != 0:
            raise GGCudaError(rc, (self.L.ggcuda_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.ggcuda_destroy(self.h)   # unregisters every page-locked target
            self.h = None
        self._pinned = {}

    __del__ = close

    def set_stream(self, stream_ptr):
        self._ck(self.L.ggcuda_set_stream(self.h, C.c_void_p(stream_ptr)))

    def begin(self, w, h):
        self._ck(self.L.ggcuda_begin(self.h, w, h))

    def begin_keyed(self, w, h, key):
        """True if the scene with this key is still resident on the device (nothing may be added then)."""
        res = C.c_int(0)
        self._ck(self.L.ggcuda_begin_keyed(self.h, w, h, C.c_uint64(key & 0xFFFFFFFFFFFFFFFF), C.byref(res)))
        return bool(res.value)

    def set_dirty_rect(self, x0, y0, x1, y1):
        self._ck(self.L.ggcuda_set_dirty_rect(self.h, x0, y0, x1, y1))

    def register_target(self, arr):
        """Page-lock a numpy pixel buffer in place (flushes into it are DMA'd directly); the context keeps it alive."""
        assert arr.dtype == np.uint8 and arr.flags.c_contiguous
        self._ck(self.L.ggcuda_register_target(self.h, _p(arr), arr.nbytes))
        if not hasattr(self, "_pinned"):
            self._pinned = {}
        self._pinned[arr.ctypes.data] = arr

    def unregister_target(self, arr):
        self._ck(self.L.ggcuda_unregister_target(self.h, _p(arr)))
        getattr(self, "_pinned", {}).pop(arr.ctypes.data, None)

    def comm_init(self, n_ranks, rank, uid):
        u = np.frombuffer(bytes(uid), dtype=np.uint8).copy()
        assert u.size == 128
        self._ck(self.L.ggcuda_comm_init(self.h, n_ranks, rank, _p(u)))

    def comm_destroy(self):
        self._ck(self.L.ggcuda_comm_destroy(self.h))

    def all_gather_bands(self, frame_ptr, band_bytes):
        self._ck(self.L.ggcuda_all_gather_bands(self.h, C.c_void_p(frame_ptr), band_bytes))

    def sync(self):
        self._ck(self.L.ggcuda_sync(self.h))

    def set_background(self, rgba_premul):
        a = np.asarray(rgba_premul, dtype=np.uint8)
        self._ck(self.L.ggcuda_set_background(self.h, _p(a)))

    def set_band(self, y0, y1):
        self._ck(self.L.ggcuda_set_band(self.h, y0, y1))

    def fill_path(self, verbs, coords, rgba_straight, fill_rule=0):
        v = np.ascontiguousarray(verbs, dtype=np.uint8)
        c = np.ascontiguousarray(coords, dtype=np.float64).ravel()
        col = np.asarray(rgba_straight, dtype=np.uint8)
        self._ck(self.L.ggcuda_fill_path(self.h, _p(v), v.size, _p(c), c.size, _p(col), int(fill_rule)))

    def fill_path_gradient(self, verbs, coords, kind, geom, stops, extend=0, fill_rule=0):
        """kind 0 linear (x0, y0, x1, y1) / 1 radial (cx, cy, r0, r1); stops: [(offset, r, g, b, a), ...] straight alpha."""
        v = np.ascontiguousarray(verbs, dtype=np.uint8)
        c = np.ascontiguousarray(coords, dtype=np.float64).ravel()
        g = np.zeros(6, dtype=np.float64)
        g[:len(geom)] = geom
        st = np.ascontiguousarray(stops, dtype=np.float64).reshape(-1, 5)
        self._ck(self.L.ggcuda_fill_path_gradient(self.h, _p(v), v.size, _p(c), c.size, int(kind), _p(g), _p(st), len(st), int(extend), int(fill_rule)))

    def add_image(self, pixels):
        """Register an image of this frame: (h, w, 4) uint8, premultiplied RGBA. Returns its index (TagImage refers to it)."""
        a = np.ascontiguousarray(pixels, dtype=np.uint8)
        assert a.ndim == 3 and a.shape[2] == 4
        ix = C.c_uint32(0)
        self._ck(self.L.ggcuda_add_image(self.h, a.shape[1], a.shape[0], _p(a), C.byref(ix)))
        return int(ix.value)

    def stroke_path(self, verbs, coords, rgba_straight, width, cap=0, join=0, miter_limit=4.0):
        v = np.ascontiguousarray(verbs, dtype=np.uint8)
        c = np.ascontiguousarray(coords, dtype=np.float64).ravel()
        col = np.asarray(rgba_straight, dtype=np.uint8)
        self._ck(self.L.ggcuda_stroke_path(self.h, _p(v), v.size, _p(c), c.size, _p(col), float(width), int(cap), int(join),
                                           float(miter_limit)))

    def push_clip(self, verbs, coords):
        v = np.ascontiguousarray(verbs, dtype=np.uint8)
        c = np.ascontiguousarray(coords, dtype=np.float64).ravel()
        self._ck(self.L.ggcuda_push_clip(self.h, _p(v), v.size, _p(c), c.size))

    def push_layer(self, blend_mode, alpha):
        self._ck(self.L.ggcuda_push_layer(self.h, int(blend_mode), float(alpha)))

    def pop(self):
        self._ck(self.L.ggcuda_pop(self.h))

    def add_encoding(self, tags, path_data, draw_data, transforms, brushes):
        t = np.ascontiguousarray(tags, dtype=np.uint8)
        pd = np.ascontiguousarray(path_data, dtype=np.float32)
        dd = np.ascontiguousarray(draw_data, dtype=np.uint32)
        tr = np.ascontiguousarray(transforms, dtype=np.float32).ravel()
        br = np.ascontiguousarray(brushes, dtype=np.float64).ravel()
        self._ck(self.L.ggcuda_add_encoding(self.h, _p(t), t.size, _p(pd), pd.size, _p(dd), dd.size, _p(tr), tr.size,
                                            _p(br), br.size // 4))

    def flush(self, dst, stride=None, flags=0):
        """dst: (H, W, 4) uint8 premultiplied RGBA (GPURenderTarget.Data)."""
        assert dst.dtype == np.uint8 and dst.flags.c_contiguous
        stride = stride or dst.strides[0]
        self._ck(self.L.ggcuda_flush(self.h, _p(dst), stride, flags))

    def upload(self):
        self._ck(self.L.ggcuda_upload(self.h))

    def render_device(self, dptr, stride, flags=0):
        self._ck(self.L.ggcuda_render_device(self.h, C.c_void_p(dptr), stride, flags))

    def render_device_multi(self, dptr, mirrors, stride, flags=0, multicast=False):
        """Band to `dptr` and to every address in `mirrors` (peer pointers, or one multicast address)."""
        arr = (C.c_void_p * max(1, len(mirrors)))(*[C.c_void_p(int(m)) for m in mirrors])
        self._ck(self.L.ggcuda_render_device_multi(self.h, C.c_void_p(dptr), arr, len(mirrors), 1 if multicast else 0, stride, flags))

    def broadcast_band(self, band_ptr, mirrors, nbytes, stream, multicast=False):
        """Copy the band of the last render into every address of `mirrors` on `stream` (a CUDA stream handle of the caller's)."""
        arr = (C.c_void_p * len(mirrors))(*[C.c_void_p(int(m)) for m in mirrors])
        self._ck(self.L.ggcuda_broadcast_band(self.h, C.c_void_p(band_ptr), arr, len(mirrors), 1 if multicast else 0, nbytes, C.c_void_p(stream)))

    def pack_host(self):
        """(packed scene words, LAYOUT record) exactly as ggcuda_upload would send them; works on host-only contexts."""
        n = self.L.ggcuda_pack_host(self.h, None, 0, None)
        if n < 0:
            self._ck(int(n))
        words = np.zeros(n, dtype=np.uint32)
        lay = np.zeros(1, dtype=LAYOUT)
        m = self.L.ggcuda_pack_host(self.h, _p(words), n, _p(lay))
        if m < 0:
            self._ck(int(m))
        return words, lay[0]   # n_scene_words of scene, 8 tail words, the gradient table

    def set_timing(self, on):
        self._ck(self.L.ggcuda_set_timing(self.h, int(on)))

    def stats(self):
        s = Stats()
        self._ck(self.L.ggcuda_get_stats(self.h, C.byref(s)))
        return s.as_dict()

    def debug_read(self, which, dtype):
        n = self.L.ggcuda_debug_read(self.h, which, None, 0)
        if n < 0:
            self._ck(int(n))
        dt = np.dtype(dtype)
        out = np.zeros(n // dt.itemsize, dtype=dt)
        if n:
            m = self.L.ggcuda_debug_read(self.h, which, _p(out), n)
            if m < 0:
                self._ck(int(m))
        return out
