"""Small scenes through the C ABI for compute-sanitizer (no torch: numpy + ctypes only).
    compute-sanitizer --tool memcheck  python tools/sanitizer_scene.py
    compute-sanitizer --tool racecheck python tools/sanitizer_scene.py
Fills + clips through the per-draw entry points, then strokes + layers with all 29 blend modes + a 6-deep clip stack through
the encoding entry, composite-over, a resident re-render with a dirty rectangle, a banded render."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import parity_util as U  # noqa: E402
from gg_b200 import _lib, scenes  # noqa: E402

ctx = _lib.Context(0)
w, h = 200, 136
elems = U.random_scene(7, w, h, 60, clips=True)
a = U.gpu_scene(ctx, elems, w, h)
enc, w2, h2 = scenes.config3(n=120, w=230, h=150, layer_every=4)
b = U.gpu_encoding(ctx, enc, w2, h2)
buf = np.full((h2, w2, 4), 77, dtype=np.uint8)
ctx.flush(buf, flags=_lib.KEEP_SCENE | _lib.COMPOSITE_OVER)
key = enc.CacheKey()
ctx.begin_keyed(w2, h2, key)
ctx.add_encoding(*enc.streams())
ctx.flush(buf)
assert ctx.begin_keyed(w2, h2, key)
ctx.set_dirty_rect(40, 30, 120, 90)
ctx.flush(buf)
c = U.gpu_encoding(ctx, enc, w2, h2, band=(2, 7))
# brushes beyond solid colours: the four gradient kinds, an SDF round rect, images (nearest and bilinear)
from gg_b200 import scene as S  # noqa: E402
e2 = S.Encoding()
rng = np.random.default_rng(3)
img = rng.integers(0, 256, (9, 13, 4)).astype(np.uint8)
img[..., :3] = np.minimum(img[..., :3], img[..., 3:])
e2.AddImage(img)
e2.EncodeTransform(S.IDENTITY)
e2.EncodeFillRoundRect((0.9, 0.2, 0.1, 0.8), (10.5, 8.5, 150.0, 90.0), 12.0, 12.0)
e2.EncodeImage(0, (1, 0, 20, 0, 1, 30))
e2.EncodeImage(0, (6.5, 1.0, 60, -1.0, 5.0, 40))
ctx.begin(w2, h2)
ctx.add_image(img)
ctx.add_encoding(*e2.streams())
stops = [(0.0, 1, 0, 0, 1), (0.5, 0, 1, 0, 0.5), (1.0, 0, 0, 1, 1)]
sq = ([0, 1, 1, 1, 4], [5, 5, 220, 5, 220, 140, 5, 140])
for kind, geom in ((0, (0, 0, 200, 100)), (1, (100, 70, 5, 80)), (2, (100, 70, 0.3, 5.0)), (3, (100, 70, 5, 80, 120, 60))):
    ctx.fill_path_gradient(*sq, kind, geom, stops, extend=kind % 3)
g = np.zeros((h2, w2, 4), dtype=np.uint8)
ctx.flush(g)
print("sanitizer scene ok", int(a.sum()), int(b.sum()), int(buf.sum()), int(c.sum()), int(g.sum()), "launches", ctx.stats()["kernel_launches"])
ctx.close()
