"""Mirror of gg's retained scene encoding (scene/encoding.go, scene/tag.go, scene/scene.go): the
input format of the hot path. Only the producers needed to build parity and benchmark scenes are
mirrored; the streams are byte-identical in layout to scene.Encoding's
(tags []u8 | pathData []f32 | drawData []u32 | transforms []Affine | brushes)."""
import struct

import numpy as np

# scene/tag.go:25-110
TagTransform, TagSetAntiAlias = 0x01, 0x02
TagBeginPath, TagMoveTo, TagLineTo, TagQuadTo, TagCubicTo, TagClosePath, TagEndPath = 0x10, 0x11, 0x12, 0x13, 0x14, 0x16, 0x17
TagFill, TagStroke, TagFillRoundRect = 0x20, 0x21, 0x22
TagPushLayer, TagPopLayer, TagBeginClip, TagEndClip = 0x30, 0x31, 0x40, 0x41
TagBrush, TagImage, TagText = 0x50, 0x51, 0x60

# scene/encoding.go:17-48
(BlendNormal, BlendMultiply, BlendScreen, BlendOverlay, BlendDarken, BlendLighten, BlendColorDodge, BlendColorBurn,
 BlendHardLight, BlendSoftLight, BlendDifference, BlendExclusion, BlendHue, BlendSaturation, BlendColor, BlendLuminosity,
 BlendClear, BlendCopy, BlendDestination, BlendSourceOver, BlendDestinationOver, BlendSourceIn, BlendDestinationIn,
 BlendSourceOut, BlendDestinationOut, BlendSourceAtop, BlendDestinationAtop, BlendXor, BlendPlus) = range(29)
NUM_BLEND_MODES = 29

FillNonZero, FillEvenOdd = 0, 1
LineCapButt, LineCapRound, LineCapSquare = 0, 1, 2
LineJoinMiter, LineJoinRound, LineJoinBevel = 0, 1, 2
MOVE, LINE, QUAD, CUBIC, CLOSE = 0, 1, 2, 3, 4
_VERB_TAG = {MOVE: TagMoveTo, LINE: TagLineTo, QUAD: TagQuadTo, CUBIC: TagCubicTo, CLOSE: TagClosePath}
IDENTITY = (1.0, 0.0, 0.0, 0.0, 1.0, 0.0)   # scene.Affine{A,B,C,D,E,F}: x' = A x + B y + C


def _f32bits(f):
    return struct.unpack("<I", struct.pack("<f", f))[0]


class Encoding:
    """scene.Encoding (scene/encoding.go:407-444) with its encoders (:477-709)."""

    def __init__(self):
        self.tags = bytearray()
        self.path_data = []      # float32 values
        self.draw_data = []      # uint32 values
        self.transforms = []     # 6-tuples
        self.brushes = []        # (r, g, b, a) float64 straight alpha (solid brushes only)
        self._chunks = []        # pre-built numpy chunks appended by the bulk encoders
        self.images = []         # (h, w, 4) uint8 arrays, premultiplied RGBA (scene.Image.Data), referenced by TagImage

    def EncodeTransform(self, t):
        self.tags.append(TagTransform)
        self.transforms.append(tuple(float(x) for x in t))

    def EncodePath(self, verbs, coords):
        """EncodePath (encoding.go:501-540): BeginPath, per-verb tags + float32 coords, EndPath."""
        if len(verbs) == 0:
            return
        self.tags.append(TagBeginPath)
        self.tags.extend(_VERB_TAG[int(v)] for v in verbs)
        self.path_data.extend(float(np.float32(c)) for c in coords)
        self.tags.append(TagEndPath)

    def EncodeFill(self, color, style=FillNonZero):
        self.tags.append(TagFill)
        self.draw_data += [len(self.brushes), int(style)]
        self.brushes.append(tuple(color))

    def EncodeStroke(self, color, width=1.0, miter_limit=4.0, cap=LineCapButt, join=LineJoinMiter):
        self.tags.append(TagStroke)
        self.draw_data += [len(self.brushes), _f32bits(width), _f32bits(miter_limit), int(cap), int(join)]
        self.brushes.append(tuple(color))

    def EncodeFillRoundRect(self, color, rect, rx, ry, style=FillNonZero):
        self.tags.append(TagFillRoundRect)
        self.draw_data += [len(self.brushes), int(style)]
        self.brushes.append(tuple(color))
        self.path_data += [float(np.float32(v)) for v in (*rect, rx, ry)]

    def EncodeImage(self, image_index, transform):
        """EncodeImage (encoding.go:656-661): the image's affine goes to the transform stream WITHOUT a TagTransform."""
        self.tags.append(TagImage)
        self.draw_data.append(int(image_index))
        self.transforms.append(tuple(float(x) for x in transform))

    def AddImage(self, pixels):
        """Register premultiplied RGBA8 pixels (h, w, 4); returns the index EncodeImage takes."""
        self.images.append(np.ascontiguousarray(pixels, dtype=np.uint8))
        return len(self.images) - 1

    def EncodePushLayer(self, blend, alpha):
        self.tags.append(TagPushLayer)
        self.draw_data += [int(blend), _f32bits(alpha)]

    def EncodePopLayer(self):
        self.tags.append(TagPopLayer)

    def EncodeBeginClip(self):
        self.tags.append(TagBeginClip)

    def EncodeEndClip(self):
        self.tags.append(TagEndClip)

    def streams(self):
        """(tags u8, pathData f32, drawData u32, transforms f32[n*6], brushes f64[n*4]) as numpy arrays -- the Go
        Encoding holds exactly these slices; they are materialised once per encoding state, not per render."""
        key = (len(self.tags), len(self.path_data), len(self.draw_data), len(self.transforms), len(self.brushes))
        if getattr(self, "_cache_key", None) != key:
            self._cache = (np.frombuffer(bytes(self.tags), dtype=np.uint8), np.asarray(self.path_data, dtype=np.float32),
                           np.asarray(self.draw_data, dtype=np.uint32), np.asarray(self.transforms, dtype=np.float32).reshape(-1),
                           np.asarray(self.brushes, dtype=np.float64).reshape(-1))
            self._cache_key = key
        return self._cache


    def Hash(self):
        """Encoding.Hash (encoding.go:752-802): FNV-1a over tags, path data, draw data and transforms."""
        return _hash(self, False)

    def CacheKey(self):
        """Hash continued over the brush colours (the reference's Hash ignores them): the key of a resident scene."""
        return _hash(self, True)


def _hash(enc, with_brushes):
    from . import _lib
    s = enc.streams()
    memo = enc.__dict__.setdefault("_hash_memo", {})
    k = (id(s[0]), with_brushes, tuple(id(im) for im in getattr(enc, "images", ())))   # (image arrays are treated as immutable once added)
    if k not in memo:
        memo.clear()
        hv = _lib.encoding_hash(s[0], s[1], s[2], s[3], s[4] if with_brushes else None)
        if with_brushes:   # ... and over the pixels of the images the encoding refers to
            for im in getattr(enc, "images", ()):
                hv = (hv * 0x100000001B3 ^ _lib.encoding_hash(im.reshape(-1), np.asarray(im.shape[:2], np.float32), [], [], None)) & 0xFFFFFFFFFFFFFFFF
        memo[k] = hv
    return memo[k]


class ArrayEncoding:
    """An Encoding held directly as numpy streams (used by the vectorised generators for 10^5-10^6 paths)."""

    Hash = Encoding.Hash
    CacheKey = Encoding.CacheKey

    def __init__(self, tags, path_data, draw_data, transforms, brushes):
        self._s = (np.ascontiguousarray(tags, np.uint8), np.ascontiguousarray(path_data, np.float32),
                   np.ascontiguousarray(draw_data, np.uint32), np.ascontiguousarray(transforms, np.float32).reshape(-1),
                   np.ascontiguousarray(brushes, np.float64).reshape(-1))

    def streams(self):
        return self._s


def circle_verbs_coords(cx, cy, r):
    """scene/path.go:192-213 (kappa = 0.5522847498, float32 arithmetic)."""
    f = np.float32
    cx, cy, r = f(cx), f(cy), f(r)
    k = f(r * f(0.5522847498))
    verbs = [MOVE, CUBIC, CUBIC, CUBIC, CUBIC, CLOSE]
    coords = [cx + r, cy, cx + r, cy + k, cx + k, cy + r, cx, cy + r, cx - k, cy + r, cx - r, cy + k, cx - r, cy,
              cx - r, cy - k, cx - k, cy - r, cx, cy - r, cx + k, cy - r, cx + r, cy - k, cx + r, cy]
    return verbs, [float(c) for c in coords]


def rect_verbs_coords(x0, y0, x1, y1):
    return [MOVE, LINE, LINE, LINE, CLOSE], [x0, y0, x1, y0, x1, y1, x0, y1]


class Scene:
    """The slice of scene.Scene (scene/scene.go:151-420) used to build test scenes. Shapes are given as
    (verbs, coords); transforms are delta-encoded like emitTransformIfNeeded (:143-148)."""

    def __init__(self):
        self.enc = Encoding()
        self._last_t = IDENTITY
        self._layers = []   # had_clip flags

    def _emit_t(self, t):
        t = tuple(float(x) for x in t)
        if t != self._last_t:
            self.enc.EncodeTransform(t)
            self._last_t = t

    def Fill(self, style, transform, color, shape):
        self._emit_t(transform)
        self.enc.EncodePath(*shape)
        self.enc.EncodeFill(color, style)

    def Stroke(self, stroke, transform, color, shape):
        self._emit_t(transform)
        self.enc.EncodePath(*shape)
        self.enc.EncodeStroke(color, **stroke)

    def PushLayer(self, blend, alpha, clip=None):
        alpha = min(1.0, max(0.0, alpha))
        self.enc.EncodePushLayer(blend, alpha)
        if clip is not None:
            self.enc.EncodePath(*clip)
            self.enc.EncodeBeginClip()
        self._layers.append(clip is not None)

    def PopLayer(self):
        if not self._layers:
            return False
        if self._layers.pop():
            self.enc.EncodeEndClip()
        self.enc.EncodePopLayer()
        return True

    def PushClip(self, shape, transform=IDENTITY):
        self._emit_t(transform)
        self.enc.EncodePath(*shape)
        self.enc.EncodeBeginClip()

    def PopClip(self):
        self.enc.EncodeEndClip()

    def Encoding(self):
        while self._layers:
            self.PopLayer()
        return self.enc
