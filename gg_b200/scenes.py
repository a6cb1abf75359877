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


def _config3_frame(sc, seed, w, h, n, layer_every, y_off, bound=None, nowipe=False):
    rng = np.random.default_rng(seed)
    depth = 0
    mode = 0
    deep_done = False
    for i in range(n):
        while nowipe and mode % S.NUM_BLEND_MODES in _WIPING:
            mode += 1
        if i % layer_every == 0:
            while depth > 0 and (depth >= 3 or rng.random() < 0.6):
                sc.PopLayer(); depth -= 1
            if not deep_done and i >= n // 2:
                for _ in range(6):   # one 6-deep stack to exercise the blend-stack spill (> 4 levels)
                    while nowipe and mode % S.NUM_BLEND_MODES in _WIPING:
                        mode += 1
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


def config3(seed=3, w=3840, h=2160, n=10000, layer_every=50, bands=1, nowipe=False):
    """10k random filled + stroked paths; every `layer_every` paths wrapped in a PushLayer cycling the 29
    scene.BlendModes, 25 % of layers with a circular clip, nesting depth <= 3 plus one 6-deep case.
    `bands` > 1: the weak-scaling canvas for N GPUs -- that many copies of this frame stacked vertically, one per GPU
    band (the same frame, so that the work per GPU is the same at every N; different random frames differ by +-40 % in
    cost, which would measure the scene generator, not the scaling). A layer without a clip of its own whose blend mode wipes whatever the layer covers gets
    its frame's rectangle as clip shape, so that it wipes its own frame, as it does on the single 4K canvas.
    `nowipe`: the layer cycle uses only the 23 blend modes that leave the backdrop alone where the layer is empty -- no tile is
    ever blanked by a layer, so fine's restart points (which this scene's wiping layers hand out generously) only come from
    opaque fills: the variant bench.py reports beside the headline scene."""
    sc = S.Scene()
    if bands == 1:
        _config3_frame(sc, seed, w, h, n, layer_every, 0.0, nowipe=nowipe)
        return sc.Encoding(), w, h
    for b in range(bands):
        frame = S.rect_verbs_coords(0.0, float(b * h), float(w), float((b + 1) * h))
        _config3_frame(sc, seed, w, h, n, layer_every, float(b * h), bound=frame, nowipe=nowipe)
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


# The in-tree icons of the reference's svg tests (svg/golden_test.go:23-27 folder, svg/stroke_hint_test.go:8-11 terminal):
# (path data, fill RGBA or None, stroke RGBA or None, stroke width, cap, join), viewBox size
ICON_FOLDER = ([("M0 0H20V20H0Z", (0x3C / 255, 0x3F / 255, 0x41 / 255, 1.0), None, 0, 0, 0),
                ("M10.5199 5.57617L10.7285 5.75H11H17C17.6904 5.75 18.25 6.30964 18.25 7V15.1667C18.25 16.0671 17.553 16.75 16.75 16.75"
                 "H3.25C2.44705 16.75 1.75 16.0671 1.75 15.1667V4.83333C1.75 3.93294 2.44705 3.25 3.25 3.25H7.63795C7.69643 3.25 7.75307 "
                 "3.2705 7.798 3.30794L10.5199 5.57617Z", None, (0xCE / 255, 0xD0 / 255, 0xD6 / 255, 1.0), 1.0, 0, 0)], 20)
ICON_TERMINAL = ([("M3 5L7 8L3 11", None, (0x6C / 255, 0x70 / 255, 0x7E / 255, 1.0), 1.0, 1, 1),
                  ("M9 11H13", None, (0x6C / 255, 0x70 / 255, 0x7E / 255, 1.0), 1.0, 1, 0)], 16)


def add_icon(sc, icon, x, y, scale):
    """What svg.RenderToScene emits for one icon (svg/scene_renderer.go): fills and strokes under a scale + translate; the
    stroke width is in user units, so it is scaled here (the scene strokes in device units, scene/renderer.go:655-713)."""
    from .svgpath import parse_path
    shapes, _ = icon
    t = (scale, 0.0, x, 0.0, scale, y)
    for d, fill, stroke, sw, cap, join in shapes:
        shape = parse_path(d)
        if fill is not None:
            sc.Fill(S.FillNonZero, t, fill, shape)
        if stroke is not None:
            sc.Stroke(dict(width=float(sw * scale), miter_limit=4.0, cap=cap, join=join), t, stroke, shape)


def config2(w=1920, h=1080, cell=16, seed=2):
    """BASELINE configs[1] stand-in (the named SVG files do not exist in the reference and its brushes
    are solid): the two in-tree test icons tiled over 1920x1080 at 16-px cells through the RenderToScene shape of calls,
    every row of icons inside a group-opacity layer. ~8 000 icons, ~24 000 fill + stroke paths."""
    rng = np.random.default_rng(seed)
    sc = S.Scene()
    for row in range(h // cell + 1):
        sc.PushLayer(S.BlendNormal, float(rng.uniform(0.6, 1.0)), None)
        for col in range(w // cell):
            icon = ICON_FOLDER if (row + col) % 2 == 0 else ICON_TERMINAL
            add_icon(sc, icon, float(col * cell), float(row * cell), cell / icon[1])
        sc.PopLayer()
    return sc.Encoding(), w, h


def config4(glyphs, w=3840, h=2160, n=50000, seed=4):
    """BASELINE configs[3]: a text-heavy 4K scene of `n` glyph outlines as paths (the reference resolves TagText to
    outlines before the GPU path; here they come pre-extracted, tests/golden/make_glyph_outlines.py): rows of printable
    ASCII in the Go Regular face at 14-48 px, quads, each glyph under its own scale + flip + translate transform;
    NonZero fills, with 2 000 EvenOdd fills and 2 000 one-pixel stroked outlines mixed in."""
    rng = np.random.default_rng(seed)
    upem = float(glyphs["units_per_em"])
    names = sorted(glyphs["glyphs"])
    tmpl = {}
    for ch in names:
        verbs, coords = [], []
        for contour in glyphs["glyphs"][ch]["contours"]:
            for seg in contour:
                if seg[0] == "M":
                    verbs.append(S.TagMoveTo); coords += seg[1:]
                elif seg[0] == "L":
                    verbs.append(S.TagLineTo); coords += seg[1:]
                else:
                    verbs.append(S.TagQuadTo); coords += seg[1:]
            verbs.append(S.TagClosePath)
        tmpl[ch] = (np.array([S.TagTransform, S.TagBeginPath] + verbs + [S.TagEndPath], np.uint8), np.array(coords, np.float32),
                    glyphs["glyphs"][ch]["advance"])
    tags, pdata, ddata, trs, brushes = [], [], [], [], []
    x, y, size = 8.0, 40.0, float(rng.uniform(14, 48))
    special = rng.permutation(n)
    kind = np.zeros(n, np.uint8)
    kind[special[:2000]] = 1      # even-odd
    kind[special[2000:4000]] = 2  # stroked outline
    for i in range(n):
        ch = names[int(rng.integers(0, len(names)))]
        tg, pc, adv = tmpl[ch]
        s = size / upem
        if x + adv * s > w - 8:
            x = 8.0
            y += size * 1.2
            size = float(rng.uniform(14, 48))
            if y > h - 8:
                y = 40.0
            s = size / upem
        tags.append(tg)
        pdata.append(pc)
        trs.append((s, 0.0, x, 0.0, -s, y))
        if kind[i] == 2:
            tags.append(np.array([S.TagStroke], np.uint8))
            ddata += [len(brushes), S._f32bits(1.0), S._f32bits(4.0), 0, 0]
        else:
            tags.append(np.array([S.TagFill], np.uint8))
            ddata += [len(brushes), int(kind[i])]
        brushes.append((*rng.uniform(0, 0.6, 3), 1.0))
        x += adv * s
    return S.ArrayEncoding(np.concatenate(tags), np.concatenate(pdata), np.array(ddata, np.uint32), np.array(trs, np.float32),
                           np.array(brushes, np.float64)), w, h


WORKLOADS = {
    "config1_512_1k_fills": config1,
    "config3_4k_10k_paths_blend_layers_clips": config3,
    "config5_16k_1m_paths": config5,
}
