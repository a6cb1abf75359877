"""Synthetic scenes of the shapes BASELINE.json names, as scene.Encoding streams.

config1: 512x512, 1 000 filled circles / cubic blobs, SrcOver (cmd/ggdemo/examples/scene shape)
config3: 3840x2160, 10 000 filled + stroked Bezier paths, 29 blend modes, layers, clips  <- bench workload
config5: 16384x16384, 1 000 000 paths (vectorised generator)
All generators are deterministic in `seed`; nothing is read from disk.
"""
import numpy as np

from . import scene as S


def _blob(rng, cx, cy, box, nseg):
    """Closed path of `nseg` cubic segments with control points uniform in a box around (cx, cy)."""
    verbs = [S.MOVE]
    p0 = np.array([cx, cy]) + rng.uniform(-box / 2, box / 2, 2)
    coords = [float(p0[0]), float(p0[1])]
    for _ in range(nseg):
        pts = np.array([cx, cy]) + rng.uniform(-box / 2, box / 2, (3, 2))
        verbs.append(S.CUBIC)
        coords += [float(v) for v in pts.ravel()]
    verbs.append(S.CLOSE)
    return verbs, coords


def config1(seed=1, w=512, h=512, n=1000):
    rng = np.random.default_rng(seed)
    sc = S.Scene()
    for i in range(n):
        col = (*rng.uniform(0, 1, 3), rng.uniform(0.5, 1.0))
        cx, cy = rng.uniform(0, w), rng.uniform(0, h)
        if i % 2 == 0:
            shape = S.circle_verbs_coords(cx, cy, rng.uniform(4, 48))
        else:
            shape = _blob(rng, cx, cy, 96.0, 4)
        sc.Fill(S.FillNonZero, S.IDENTITY, col, shape)
    return sc.Encoding(), w, h


# scene.BlendModes whose layer changes the backdrop even where the layer is empty (Porter-Duff with Fb != 1 at Sa == 0)
_WIPING = (S.BlendClear, S.BlendCopy, S.BlendSourceIn, S.BlendDestinationIn, S.BlendSourceOut, S.BlendDestinationAtop)


def _config3_frame(sc, seed, w, h, n, layer_every, y_off, bound=None):
    rng = np.random.default_rng(seed)
    depth = 0
    mode = 0
    deep_done = False
    for i in range(n):
        if i % layer_every == 0:
            while depth > 0 and (depth >= 3 or rng.random() < 0.6):
                sc.PopLayer(); depth -= 1
            if not deep_done and i >= n // 2:
                for _ in range(6):   # one 6-deep stack to exercise the blend-stack spill (> 4 levels)
                    sc.PushLayer(mode % S.NUM_BLEND_MODES, float(rng.uniform(0.3, 1.0)), bound if mode % S.NUM_BLEND_MODES in _WIPING else None)
                    mode += 1; depth += 1
                deep_done = True
            else:
                clip = bound if mode % S.NUM_BLEND_MODES in _WIPING else None
                if rng.random() < 0.25:
                    clip = S.circle_verbs_coords(rng.uniform(0, w), y_off + rng.uniform(0, h), rng.uniform(200, 900))
                sc.PushLayer(mode % S.NUM_BLEND_MODES, float(rng.uniform(0.3, 1.0)), clip); mode += 1; depth += 1
        box = rng.uniform(32, 512)
        cx, cy = rng.uniform(0, w), y_off + rng.uniform(0, h)
        shape = _blob(rng, cx, cy, box, int(rng.integers(3, 7)))
        col = (*rng.uniform(0, 1, 3), rng.uniform(0.2, 1.0))
        if rng.random() < 0.5:
            sc.Fill(S.FillEvenOdd if rng.random() < 0.25 else S.FillNonZero, S.IDENTITY, col, shape)
        else:
            sc.Stroke(dict(width=float(rng.uniform(1, 12)), miter_limit=4.0, cap=int(rng.integers(0, 3)), join=int(rng.integers(0, 3))),
                      S.IDENTITY, col, shape)
        if deep_done and depth > 3 and i % layer_every == layer_every - 1:
            while depth > 0:
                sc.PopLayer(); depth -= 1
    while depth > 0:
        sc.PopLayer(); depth -= 1


def config3(seed=3, w=3840, h=2160, n=10000, layer_every=50, bands=1):
    """10k random filled + stroked paths; every `layer_every` paths wrapped in a PushLayer cycling the 29
    scene.BlendModes, 25 % of layers with a circular clip, nesting depth <= 3 plus one 6-deep case.
    `bands` > 1: the weak-scaling canvas for N GPUs -- that many such frames (seeds seed, seed + 1, ...) stacked
    vertically, one per GPU band. A layer without a clip of its own whose blend mode wipes whatever the layer covers gets
    its frame's rectangle as clip shape, so that it wipes its own frame, as it does on the single 4K canvas."""
    sc = S.Scene()
    if bands == 1:
        _config3_frame(sc, seed, w, h, n, layer_every, 0.0)
        return sc.Encoding(), w, h
    for b in range(bands):
        frame = S.rect_verbs_coords(0.0, float(b * h), float(w), float((b + 1) * h))
        _config3_frame(sc, seed + b, w, h, n, layer_every, float(b * h), bound=frame)
    return sc.Encoding(), w, h * bands


def config5(seed=5, size=16384, n=1_000_000, min_box=16.0, max_box=256.0):
    """n closed 4-cubic paths (half kappa circles, half random blobs), built as numpy streams."""
    rng = np.random.default_rng(seed)
    f = np.float32
    cx = rng.uniform(0, size, n).astype(f)
    cy = rng.uniform(0, size, n).astype(f)
    box = rng.uniform(min_box, max_box, n).astype(f)
    pts = np.empty((n, 13, 2), dtype=f)   # move + 4 x (c1, c2, end)
    # blobs
    off = rng.uniform(-0.5, 0.5, (n, 13, 2)).astype(f) * box[:, None, None]
    pts[:, :, 0] = cx[:, None] + off[:, :, 0]
    pts[:, :, 1] = cy[:, None] + off[:, :, 1]
    pts[:, 12] = pts[:, 0]   # closed
    # circles for even indices
    ev = np.arange(0, n, 2)
    r = (box[ev] * f(0.5)).astype(f)
    k = (r * f(0.5522847498)).astype(f)
    x, y = cx[ev], cy[ev]
    circ = np.stack([
        np.stack([x + r, y], -1),
        np.stack([x + r, y + k], -1), np.stack([x + k, y + r], -1), np.stack([x, y + r], -1),
        np.stack([x - k, y + r], -1), np.stack([x - r, y + k], -1), np.stack([x - r, y], -1),
        np.stack([x - r, y - k], -1), np.stack([x - k, y - r], -1), np.stack([x, y - r], -1),
        np.stack([x + k, y - r], -1), np.stack([x + r, y - k], -1), np.stack([x + r, y], -1)], axis=1).astype(f)
    pts[ev] = circ
    tags = np.tile(np.array([S.TagBeginPath, S.TagMoveTo, S.TagCubicTo, S.TagCubicTo, S.TagCubicTo, S.TagCubicTo,
                             S.TagClosePath, S.TagEndPath, S.TagFill], dtype=np.uint8), n)
    draw = np.empty((n, 2), dtype=np.uint32)
    draw[:, 0] = np.arange(n, dtype=np.uint32)
    draw[:, 1] = S.FillNonZero
    brushes = np.concatenate([rng.uniform(0, 1, (n, 3)), rng.uniform(0.5, 1.0, (n, 1))], axis=1)
    return S.ArrayEncoding(tags, pts.reshape(-1), draw.reshape(-1), np.zeros(0, np.float32), brushes), size, size


WORKLOADS = {
    "config1_512_1k_fills": config1,
    "config3_4k_10k_paths_blend_layers_clips": config3,
    "config5_16k_1m_paths": config5,
}
