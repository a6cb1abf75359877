"""GPU parity tests proper: the CUDA pipeline, called through the C ABI, against the CPU twin
oracle on identical scenes. Integer stages bit-exact; pixels within 1/255 of the twin (the only
difference is FMA contraction and segment order inside a tile)."""
import numpy as np
import pytest
from PIL import Image

import parity_util as U
from gg_b200 import _lib as G

pytestmark = pytest.mark.gpu
GOLDEN = __import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "vello-gpu-pipeline")


def _check(ctx, elems, w, h, bg=(0, 0, 0, 0), ptcl=True):
    out = U.gpu_scene(ctx, elems, w, h, bg)
    oc = U.oracle_scene(elems, w, h)
    rep = U.compare_stages(ctx, oc, elems, w, h, check_ptcl=ptcl)
    _, o_premul = oc.fine(_straight_bg(bg), straight=False, premul=True)
    mx, mean, frac = U.pixel_diff(out, o_premul)
    assert mx <= 1, f"max pixel diff {mx}"
    assert frac <= 0.002, f"{frac*100:.3f}% pixels differ"
    return out, oc, rep


def _straight_bg(bg_premul):
    r, g, b, a = bg_premul
    if a == 0:
        return (0, 0, 0, 0)
    return (min(255, round(r * 255 / a)), min(255, round(g * 255 / a)), min(255, round(b * 255 / a)), a)


@pytest.mark.parametrize("name,size,color,bg,eo,shape", [
    ("filled_circle", 100, (0, 255, 0, 255), (255, 255, 255, 255), False, ("circle", 50, 50, 45)),
    ("filled_triangle", 100, (0, 255, 0, 255), (255, 255, 255, 255), False, ("poly", [(5, 5), (95, 50), (5, 95)])),
    ("filling_nonzero_rule", 100, (128, 0, 0, 255), (255, 255, 255, 255), False, ("poly", [(50, 10), (75, 90), (10, 40), (90, 40), (25, 90)])),
    ("filling_evenodd_rule", 100, (128, 0, 0, 255), (255, 255, 255, 255), True, ("poly", [(50, 10), (75, 90), (10, 40), (90, 40), (25, 90)])),
    ("smoke_filled_circle", 20, (0, 0, 255, 255), (0, 0, 0, 255), False, ("circle", 10, 10, 7)),
    ("smoke_filled_square", 20, (0, 0, 255, 255), (0, 0, 0, 255), False, ("poly", [(7, 7), (13, 7), (13, 13), (7, 13)])),
])
def test_vello_goldens(ctx, name, size, color, bg, eo, shape):
    """The reference's own golden scenes (tilecompute/rasterizer_test.go:39-135) through the CUDA path."""
    if shape[0] == "circle":
        v, c = U.circle_path(*[np.float32(x) for x in shape[1:]])
    else:
        v, c = U.polygon_path(shape[1])
    elems = [dict(type="draw", verbs=v, coords=c, color=color, even_odd=eo)]
    out, oc, _ = _check(ctx, elems, size, size, bg)
    ref = np.array(Image.open(f"{GOLDEN}/{name}.png").convert("RGBA"))
    # goldens are opaque, so premultiplied == straight
    ndiff = int((out != ref).any(axis=2).sum())
    thr = {"filling_nonzero_rule": 0.15, "filling_evenodd_rule": 0.15}.get(name, 0.0)
    assert ndiff / (size * size) * 100 <= thr, f"{ndiff} px differ from the Vello golden"


@pytest.mark.parametrize("seed,w,h,n", [(1, 256, 256, 60), (2, 512, 384, 300), (3, 1000, 700, 800), (4, 130, 70, 40)])
def test_random_fills(ctx, seed, w, h, n):
    _check(ctx, U.random_scene(seed, w, h, n), w, h, bg=(0, 0, 0, 0))


@pytest.mark.parametrize("seed,w,h,n", [(11, 256, 256, 80), (12, 512, 512, 400), (13, 333, 222, 200)])
def test_random_clips(ctx, seed, w, h, n):
    _check(ctx, U.random_scene(seed, w, h, n, clips=True), w, h, bg=(255, 255, 255, 255))


def test_empty_scene(ctx):
    out = U.gpu_scene(ctx, [], 64, 48, bg=(10, 20, 30, 255))
    assert (out == np.array([10, 20, 30, 255], dtype=np.uint8)).all()


def test_offcanvas_and_degenerate(ctx):
    elems = []
    for (cx, cy) in [(-100, 50), (50, -100), (400, 50), (50, 400), (-5, -5), (130, 130)]:
        v, c = U.circle_path(np.float32(cx), np.float32(cy), np.float32(30))
        elems.append(dict(type="draw", verbs=v, coords=c, color=(255, 0, 0, 200)))
    v, c = U.polygon_path([(10, 10), (10, 10), (10, 10)])
    elems.append(dict(type="draw", verbs=v, coords=c, color=(0, 255, 0, 255)))
    out = U.gpu_scene(ctx, elems, 128, 128)
    oc = U.oracle_scene([e for e in elems[4:6]], 128, 128)
    _, op = oc.fine((0, 0, 0, 0), straight=False)
    # only the two partially visible circles contribute
    mx, _, _ = U.pixel_diff(out, op)
    assert mx <= 1


def test_bands_match_full_frame(ctx):
    """Rendering the canvas in horizontal bands (the multi-GPU decomposition) gives the full-frame pixels."""
    w, h = 512, 512
    elems = U.random_scene(21, w, h, 300, clips=True)
    full = U.gpu_scene(ctx, elems, w, h, bg=(0, 0, 0, 0)).copy()
    ht = (h + 15) // 16
    acc = np.zeros_like(full)
    for (y0, y1) in [(0, 5), (5, 13), (13, 14), (14, ht)]:
        part = U.gpu_scene(ctx, elems, w, h, bg=(0, 0, 0, 0), band=(y0, y1))
        acc[y0 * 16:y1 * 16] = part[y0 * 16:y1 * 16]
    assert (acc == full).all()


def _check_encoding(ctx, enc, w, h, bg=(0, 0, 0, 0)):
    out = U.gpu_encoding(ctx, enc, w, h, bg)
    oc = U.oracle_from_ctx(ctx, w, h)
    rep = U.compare_stages(ctx, oc, None, w, h)
    _, ref = oc.fine(_straight_bg(bg), straight=False, premul=True)
    mx, mean, frac = U.pixel_diff(out, ref)
    assert mx <= 1 and frac <= 0.002, (mx, frac)
    return out, oc, rep


def test_encoding_config1(ctx):
    """BASELINE configs[0]: 512x512, 1 000 filled circles / cubic blobs through the scene.Encoding entry."""
    from gg_b200 import scenes
    enc, w, h = scenes.config1()
    _check_encoding(ctx, enc, w, h)


def test_encoding_layers_clips_transforms(ctx):
    """Layers (elided where empty), clips, even-odd fills, strokes, round rects and non-identity transforms."""
    from gg_b200 import scene as S
    rng = np.random.default_rng(5)
    sc = S.Scene()
    w, h = 640, 480
    for i in range(120):
        if i % 10 == 0:
            clip = S.circle_verbs_coords(rng.uniform(0, w), rng.uniform(0, h), rng.uniform(60, 200)) if i % 20 == 0 else None
            sc.PushLayer(int(rng.integers(0, 16)), float(rng.uniform(0.3, 1.0)), clip)
        t = (1.0, 0.0, 0.0, 0.0, 1.0, 0.0) if i % 3 else (float(rng.uniform(0.5, 1.5)), float(rng.uniform(-0.3, 0.3)), float(rng.uniform(-20, 20)),
                                                             float(rng.uniform(-0.3, 0.3)), float(rng.uniform(0.5, 1.5)), float(rng.uniform(-20, 20)))
        shape = S.circle_verbs_coords(rng.uniform(0, w), rng.uniform(0, h), rng.uniform(5, 60))
        col = (*rng.uniform(0, 1, 3), rng.uniform(0.3, 1.0))
        if i % 4 == 1:
            sc.Stroke(dict(width=float(rng.uniform(1, 9)), cap=int(rng.integers(0, 3)), join=int(rng.integers(0, 3))), t, col, shape)
        elif i % 4 == 2:
            sc.Fill(S.FillEvenOdd, t, col, shape)
        else:
            sc.Fill(S.FillNonZero, t, col, shape)
        if i % 10 == 9:
            sc.PopLayer()
        if i % 17 == 0:
            sc.enc.EncodeFillRoundRect(col, (rng.uniform(0, w / 2), rng.uniform(0, h / 2), rng.uniform(w / 2, w), rng.uniform(h / 2, h)), 12.0, 9.0)
    _check_encoding(ctx, sc.Encoding(), w, h, bg=(255, 255, 255, 255))


def test_buffer_growth_from_cold_context():
    """A fresh context starts with small buffers: the first render must grow them (several passes) and still be right."""
    from gg_b200 import _lib, scenes
    c = _lib.Context(0)
    try:
        enc, w, h = scenes.config3(n=1500, w=1920, h=1080)
        out = U.gpu_encoding(c, enc, w, h)
        st = c.stats()
        assert st["passes"] >= 1
        oc = U.oracle_from_ctx(c, w, h)
        U.compare_stages(c, oc, None, w, h)
        _, ref = oc.fine((0, 0, 0, 0), straight=False, premul=True)
        mx, _, frac = U.pixel_diff(out, ref)
        assert mx <= 1 and frac <= 0.002
        out2 = U.gpu_encoding(c, enc, w, h)
        assert c.stats()["passes"] == 1, "steady state must be a single pass"
        assert U.pixel_diff(out, out2)[0] <= 1
    finally:
        c.close()


@pytest.mark.parametrize("bg", [(0, 0, 0, 0), (255, 255, 255, 255), (40, 80, 120, 200)])
def test_all_29_blend_modes(ctx, bg):
    """One layer per scene.BlendMode over a varied backdrop; CUDA fine vs the oracle's float32 definition."""
    from gg_b200 import scene as S
    rng = np.random.default_rng(29)
    sc = S.Scene()
    w, h = 29 * 40, 160
    for i in range(40):   # backdrop: translucent blobs over the whole strip
        sc.Fill(S.FillNonZero, S.IDENTITY, (*rng.uniform(0, 1, 3), rng.uniform(0.3, 1.0)),
                S.circle_verbs_coords(rng.uniform(0, w), rng.uniform(0, h), rng.uniform(20, 90)))
    for mode in range(S.NUM_BLEND_MODES):
        x0 = mode * 40
        sc.PushLayer(mode, float(rng.uniform(0.5, 1.0)), S.rect_verbs_coords(x0 + 2, 4, x0 + 38, h - 4) if mode % 2 else None)
        for _ in range(3):
            sc.Fill(S.FillNonZero, S.IDENTITY, (*rng.uniform(0, 1, 3), rng.uniform(0.2, 1.0)),
                    S.circle_verbs_coords(x0 + rng.uniform(5, 35), rng.uniform(10, h - 10), rng.uniform(8, 30)))
        sc.PopLayer()
    _check_encoding(ctx, sc.Encoding(), w, h, bg=bg)


def test_deep_clip_stack_spills(ctx):
    """Seven nested layers/clips: levels 0-1 shared memory, 2-3 local memory, 4+ global spill (fine.go:58-62)."""
    from gg_b200 import scene as S
    sc = S.Scene()
    w, h = 200, 200
    sc.Fill(S.FillNonZero, S.IDENTITY, (0.2, 0.4, 0.9, 1.0), S.rect_verbs_coords(10, 10, 190, 190))
    for d in range(7):
        sc.PushLayer(d % 3, 0.9, S.circle_verbs_coords(100, 100, 95 - 8 * d))
        sc.Fill(S.FillNonZero, S.IDENTITY, (0.9 - 0.1 * d, 0.1 * d, 0.5, 0.7), S.rect_verbs_coords(20 + 5 * d, 30, 180 - 5 * d, 170))
    _check_encoding(ctx, sc.Encoding(), w, h, bg=(255, 255, 255, 255))


def test_fine_restart_points(ctx):
    """Opaque full-tile fills and backdrop-wiping layers let fine skip the head of a tile's PTCL (coarse records the
    restart point); pixels must be what the full replay (oracle) gives, PTCL unchanged."""
    from gg_b200 import scene as S
    rng = np.random.default_rng(77)
    w, h = 320, 240

    def blobs(sc, n):
        for _ in range(n):
            sc.Fill(S.FillNonZero, S.IDENTITY, (*rng.uniform(0, 1, 3), rng.uniform(0.3, 1.0)),
                    S.circle_verbs_coords(rng.uniform(0, w), rng.uniform(0, h), rng.uniform(10, 80)))
    sc = S.Scene()
    blobs(sc, 30)
    sc.Fill(S.FillNonZero, S.IDENTITY, (0.2, 0.3, 0.4, 1.0), S.rect_verbs_coords(0, 0, w, 128))       # opaque: restart (solid tiles)
    blobs(sc, 10)
    sc.PushLayer(S.BlendCopy, 1.0); sc.PopLayer()                                                        # empty Copy layer wipes everything
    blobs(sc, 10)
    sc.PushLayer(S.BlendClear, 0.8); blobs(sc, 5); sc.PopLayer()                                         # Clear wipes whatever it holds
    blobs(sc, 10)
    sc.PushLayer(S.BlendSourceIn, 1.0); blobs(sc, 3); sc.PopLayer()                                      # not empty: no restart, real SrcIn
    blobs(sc, 5)
    sc.PushLayer(S.BlendNormal, 1.0, S.circle_verbs_coords(160, 120, 100))
    sc.Fill(S.FillNonZero, S.IDENTITY, (0.9, 0.1, 0.1, 1.0), S.rect_verbs_coords(0, 0, w, h))           # opaque but inside a clip: no restart
    sc.PopLayer()
    blobs(sc, 5)
    for bg in ((0, 0, 0, 0), (255, 255, 255, 255)):
        _check_encoding(ctx, sc.Encoding(), w, h, bg=bg)


def test_long_hit_lists(ctx):
    """More than 1024 draws over one tile: coarse sorts such lists in place in global memory instead of shared memory."""
    rng = np.random.default_rng(3)
    elems = []
    for i in range(1500):
        x0, y0 = rng.uniform(0, 20, 2)
        v, c = U.polygon_path(np.array([(x0, y0), (x0 + rng.uniform(5, 40), y0), (x0 + rng.uniform(5, 40), y0 + rng.uniform(5, 40))], dtype=np.float32))
        elems.append(dict(type="draw", verbs=v, coords=c, color=tuple(int(q) for q in rng.integers(0, 256, 3)) + (40,)))
    _check(ctx, elems, 64, 64, bg=(255, 255, 255, 255))


@pytest.mark.parametrize("name", ["polygon", "float-rect-aa", "star-aa"])
def test_cuda_vs_gg_cpu_aaa_goldens(ctx, name):
    """CUDA coverage against gg's CPU filler on the Skia-AAA golden paths (see test_cpu_oracle.test_exact_area_vs_gg_cpu_aaa
    for why max |d| is large on snapped edges): same statistics as the CPU twin, i.e. the gap is the algorithm's, not CUDA's."""
    import test_cpu_oracle as O
    cov_g = O.skia_golden_coverage(name)
    v, c = U.polygon_path(np.array(O.SKIA_AAA[name], dtype=np.float32))
    out = U.gpu_scene(ctx, [dict(type="draw", verbs=v, coords=c, color=(255, 255, 255, 255))], 100, 100)
    d = np.abs(out[..., 3].astype(np.float64) - cov_g)
    assert d.mean() <= 0.7 and (d <= 2.5).mean() >= 0.97 and d.max() <= 61


# ---- device-side stroke expansion (gg_b200/csrc/stroke.cuh) against its CPU statement (oracle/twin.c ot_stroke_segment) ----
def _stroke_scene(seed, w, h, n, transforms=False):
    from gg_b200 import scene as S
    rng = np.random.default_rng(seed)
    sc = S.Scene()
    for i in range(n):
        kind = i % 6
        cx, cy = rng.uniform(0, w), rng.uniform(0, h)
        box = rng.uniform(8, 160)
        closed = bool(rng.random() < 0.5)
        if kind == 0:     # polyline
            m = int(rng.integers(1, 7))
            pts = np.stack([cx + rng.uniform(-box, box, m + 1), cy + rng.uniform(-box, box, m + 1)], 1)
            verbs, coords = [S.MOVE] + [S.LINE] * m, list(pts.ravel())
        elif kind == 1:   # cubics, possibly with cusps and loops
            m = int(rng.integers(1, 5))
            verbs, coords = [S.MOVE] + [S.CUBIC] * m, list(np.concatenate([[cx, cy], (np.array([cx, cy]) + rng.uniform(-box, box, (3 * m, 2))).ravel()]))
        elif kind == 2:   # quads
            m = int(rng.integers(1, 5))
            verbs, coords = [S.MOVE] + [S.QUAD] * m, list(np.concatenate([[cx, cy], (np.array([cx, cy]) + rng.uniform(-box, box, (2 * m, 2))).ravel()]))
        elif kind == 3:   # small circle under a wide stroke (offset > radius of curvature everywhere)
            verbs, coords = S.circle_verbs_coords(cx, cy, rng.uniform(0.5, 6))
            closed = False
        elif kind == 4:   # two subpaths, degenerate pieces: repeated points, a zero-length curve, a lone MoveTo (dot)
            verbs = [S.MOVE, S.LINE, S.LINE, S.CUBIC, S.MOVE, S.MOVE, S.LINE, S.QUAD]
            a, b = (cx, cy), (cx + box, cy + box / 3)
            coords = [*a, *a, *b, *b, *b, *b, cx - 9, cy - 9, cx + 5, cy - box, cx + 5, cy, cx + 5, cy, cx + box / 2, cy]
        else:             # mixed
            verbs = [S.MOVE, S.LINE, S.CUBIC, S.QUAD, S.LINE]
            coords = list(np.concatenate([[cx, cy], (np.array([cx, cy]) + rng.uniform(-box, box, (7, 2))).ravel()]))
        if closed:
            verbs = list(verbs) + [S.CLOSE]
        t = S.IDENTITY
        if transforms and i % 3 == 0:
            t = (float(rng.uniform(0.5, 1.5)), float(rng.uniform(-0.4, 0.4)), float(rng.uniform(-20, 20)),
                 float(rng.uniform(-0.4, 0.4)), float(rng.uniform(0.5, 1.5)), float(rng.uniform(-20, 20)))
        col = (*rng.uniform(0, 1, 3), rng.uniform(0.3, 1.0))
        sc.Stroke(dict(width=float(rng.choice([0.3, 1.0, 2.5, 7.0, 12.0, 30.0])), miter_limit=float(rng.choice([1.0, 4.0, 10.0])),
                       cap=int(rng.integers(0, 3)), join=int(rng.integers(0, 3))), t, col, (verbs, [float(v) for v in coords]))
        if i % 5 == 0:    # fills in between: the two kinds share the work list
            sc.Fill(S.FillNonZero, t, col, S.circle_verbs_coords(rng.uniform(0, w), rng.uniform(0, h), rng.uniform(4, 40)))
    return sc.Encoding()


@pytest.mark.parametrize("seed,w,h,n,tr", [(21, 512, 384, 240, False), (22, 700, 500, 400, True), (23, 128, 96, 60, True)])
def test_strokes_device_expansion(ctx, seed, w, h, n, tr):
    """Stroked paths (all caps / joins / miter limits, open and closed, cusps, degenerate pieces, transforms): the lines the
    device writes are the oracle's, bit for bit and in order; every later stage follows as for fills."""
    _check_encoding(ctx, _stroke_scene(seed, w, h, n, tr), w, h)


def test_strokes_per_draw_api_and_bands(ctx):
    """GPUAccelerator.StrokePath entry (ggcuda_stroke_path) and band culling of stroked segments."""
    w, h = 400, 320
    rng = np.random.default_rng(31)

    def build():
        ctx.begin(w, h)
        for i in range(80):
            m = int(rng.integers(1, 4))
            c = rng.uniform(0, [w, h] * (1 + 3 * m))
            ctx.stroke_path([0] + [3] * m + ([4] if i % 2 else []), c, tuple(int(q) for q in rng.integers(60, 256, 4)),
                            float(rng.uniform(0.5, 20)), int(i % 3), int((i // 3) % 3), 4.0)
    build()
    full = np.zeros((h, w, 4), np.uint8)
    ctx.set_band(0, (h + 15) // 16)
    ctx.flush(full, flags=G.KEEP_SCENE)
    oc = U.oracle_from_ctx(ctx, w, h)
    U.compare_stages(ctx, oc, None, w, h)
    _, ref = oc.fine((0, 0, 0, 0), straight=False, premul=True)
    mx, _, frac = U.pixel_diff(full, ref)
    assert mx <= 1 and frac <= 0.002
    parts = np.zeros_like(full)
    for y0, y1 in ((0, 7), (7, 13), (13, 20)):
        ctx.set_band(y0, y1)
        ctx.flush(parts, flags=G.KEEP_SCENE)
    assert (parts == full).all()


def test_strokes_device_vs_host_stroker(ctx):
    """Same scene through the diagnostic host polyline stroker (GGCUDA_CREATE_HOST_STROKES): the two outlines are
    different constructions of the same stroke, so pixels agree except along edges (flattening differs) and where the
    host stroker is wrong (it leaves holes / spikes where the offset exceeds the radius of curvature)."""
    from gg_b200 import scene as S
    rng = np.random.default_rng(41)
    sc = S.Scene()
    w, h = 512, 512
    for i in range(60):
        pts = rng.uniform(0, w, (5, 2))
        sc.Stroke(dict(width=float(rng.uniform(1, 10)), cap=int(i % 3), join=int((i // 3) % 3)), S.IDENTITY, (1, 1, 1, 1),
                  ([S.MOVE] + [S.LINE] * 4 + ([S.CLOSE] if i % 2 else []), [float(v) for v in pts.ravel()]))
    a = U.gpu_encoding(ctx, sc.Encoding(), w, h)
    hc = G.Context(0, G.CREATE_HOST_STROKES)
    try:
        b = U.gpu_encoding(hc, sc.Encoding(), w, h)
    finally:
        hc.close()
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    assert d.mean() < 0.02 and (d > 8).mean() < 1e-3, (d.max(), d.mean())   # round joins / caps are tessellated differently


# ---- BASELINE.json's configurations at their full sizes ----
def _full_size(ctx, enc, w, h, bg=(0, 0, 0, 0), threads=None):
    import os
    out = U.gpu_encoding(ctx, enc, w, h, bg)
    oc = U.oracle_from_ctx(ctx, w, h)
    rep = U.compare_stages_fast(ctx, oc, w, h)
    from oracle import twin as T
    ref, _ = T.render_packed(ctx.debug_read(G.BUF_SCENE, np.uint32), ctx.debug_read(G.BUF_LAYOUT, G.LAYOUT)[0], w, h, bg, threads or os.cpu_count())
    d = np.abs(out.astype(np.int32) - ref.astype(np.int32))
    assert d.max() <= 1 and (d.max(axis=-1) > 0).mean() <= 0.002, (d.max(), (d.max(axis=-1) > 0).mean())
    return out, rep


def test_config2_svg_icons_1080p(ctx):
    """configs[1] stand-in: ~8 000 SVG icons (fills + round / miter strokes under scale transforms, group opacity)."""
    from gg_b200 import scenes
    enc, w, h = scenes.config2()
    _full_size(ctx, enc, w, h, bg=(255, 255, 255, 255))


def test_config3_4k_full_size(ctx):
    """configs[2], the benchmark scene itself: 3840x2160, 10 000 filled + stroked paths, 29 blend modes, layers, clips --
    every integer stage bit-exact against the oracle, pixels within 1/255."""
    from gg_b200 import scenes
    enc, w, h = scenes.config3()
    _, rep = _full_size(ctx, enc, w, h)
    assert rep["lines"] > 1_000_000 and rep["segments"] > 2_000_000


def test_config4_glyph_outlines_4k(ctx):
    """configs[3]: 50 000 glyph outlines (quads) under per-glyph transforms, NonZero / EvenOdd fills and stroked outlines."""
    import json, os
    from gg_b200 import scenes
    glyphs = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "fixtures", "goregular_ascii.json")))
    enc, w, h = scenes.config4(glyphs)
    _full_size(ctx, enc, w, h, bg=(255, 255, 255, 255))


def test_config5_16k_one_million_paths(ctx):
    """configs[4]: 16384x16384, 1 000 000 paths. Too large for the oracle as a whole, so: (a) the oracle renders the
    top-left 1024x1024 of the same encoding (a smaller canvas only clips) and the CUDA frame must match it there;
    (b) eight bands rendered one after the other are the full frame; (c) a second render is the same frame."""
    from gg_b200 import scenes
    from oracle import twin as T
    import os
    enc, w, h = scenes.config5()
    c5 = G.Context(0)
    try:
        full = U.gpu_encoding(c5, enc, w, h)
        st = c5.stats()
        assert st["n_draws"] == 1_000_000 and st["n_lines"] > 10_000_000
        again = np.zeros_like(full)
        c5.flush(again, flags=G.KEEP_SCENE)
        assert (full == again).all()       # fine accumulates coverage in fixed point: the order segments arrive in a tile does not matter
        parts = np.zeros_like(full)
        ht = h // 16
        for b in range(8):
            c5.set_band(b * ht // 8, (b + 1) * ht // 8)
            c5.flush(parts, flags=G.KEEP_SCENE)
        assert (full == parts).all()
    finally:
        c5.close()
    crop = 1024
    hc = G.Context(-1)
    hc.begin(crop, crop)
    hc.add_encoding(*enc.streams())
    words, lay = hc.pack_host()
    hc.close()
    ref, _ = T.render_packed(words, lay, crop, crop, (0, 0, 0, 0), os.cpu_count())
    d = np.abs(full[:crop, :crop].astype(np.int32) - ref.astype(np.int32))
    assert d.max() <= 1 and (d.max(axis=-1) > 0).mean() <= 0.002, (d.max(), (d.max(axis=-1) > 0).mean())


def test_degenerate_random_scenes(ctx):
    """The structured random scenes of test_cpu_host.test_random_scenes_pack_and_render (empty paths, repeated points,
    collapsed curves, zero-width and huge strokes, unbalanced pops, clips of empty paths, all 29 layer modes, transforms)
    on the device: every stage against the oracle."""
    from gg_b200 import scene as S
    rng = np.random.default_rng(2)

    def rand_path():
        v, co = [], []
        for _ in range(int(rng.integers(0, 7))):
            k = int(rng.choice([S.MOVE, S.LINE, S.LINE, S.QUAD, S.CUBIC, S.CLOSE]))
            m = {S.MOVE: 2, S.LINE: 2, S.QUAD: 4, S.CUBIC: 6, S.CLOSE: 0}[k]
            pts = rng.uniform(-20, 90, m)
            if rng.random() < 0.2 and m >= 2 and len(co) >= 2:
                pts[:2] = co[-2:]
            if rng.random() < 0.1 and m:
                pts[:] = pts[0]
            v.append(k)
            co += list(pts)
        return v, co
    done = 0
    for _ in range(120):
        sc, depth = S.Scene(), 0
        for _ in range(int(rng.integers(1, 12))):
            r = rng.random()
            t = S.IDENTITY if rng.random() < 0.7 else tuple(rng.uniform(-1.5, 1.5, 6))
            col = tuple(rng.uniform(0, 1, 4))
            if r < 0.35:
                sc.Fill(int(rng.integers(0, 2)), t, col, rand_path())
            elif r < 0.7:
                sc.Stroke(dict(width=float(rng.choice([0, 0.01, 1, 5, 40])), miter_limit=float(rng.choice([0, 1, 4, 100])),
                               cap=int(rng.integers(0, 3)), join=int(rng.integers(0, 3))), t, col, rand_path())
            elif r < 0.8:
                sc.PushLayer(int(rng.integers(0, 29)), float(rng.uniform(0, 1)), rand_path() if rng.random() < 0.5 else None)
                depth += 1
            elif r < 0.88 and depth:
                sc.PopLayer()
                depth -= 1
            elif r < 0.94:
                sc.PushClip(rand_path(), t)
            else:
                sc.PopClip()
        w, h = int(rng.integers(1, 80)), int(rng.integers(1, 80))
        try:
            _check_encoding(ctx, sc.Encoding(), w, h)
        except G.GGCudaError as e:
            assert e.code in (G.ERR_INVALID, G.ERR_UNSUPPORTED)
            continue
        done += 1
    assert done > 100


# ---------------------------------------------------------------- round 2: paths nobody tested in round 1
def composite_over_bytes(src, dst):
    """VelloAccelerator.compositeOver (vello_accelerator.go:388-442) on premultiplied RGBA8 arrays: sA == 0 keeps the
    target, sA == 255 overwrites, otherwise s + (d * (255 - sA) + 127) / 255 per byte (uint8 wrap as in Go)."""
    s, d = src.astype(np.uint32), dst.astype(np.uint32)
    inv = 255 - s[..., 3:4]
    o = (s + (d * inv + 127) // 255) & 0xFF
    sa = src[..., 3:4]
    return np.where(sa == 0, dst, np.where(sa == 255, src, o.astype(np.uint8))).astype(np.uint8)


def _busy_target(w, h, seed=3, stride_pad=0):
    """A target that already holds content: premultiplied RGBA8 gradient + noise, optionally with a padded stride."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, (h, w, 1), dtype=np.uint8)
    a[: h // 3] = 255
    a[h // 3: h // 2, : w // 2] = 0
    rgb = (rng.integers(0, 256, (h, w, 3)).astype(np.uint32) * a // 255).astype(np.uint8)
    buf = np.zeros((h, w + stride_pad, 4), dtype=np.uint8)
    buf[:, :w, :3] = rgb
    buf[:, :w, 3:] = a
    buf[:, w:] = 0x5A   # padding bytes must survive
    return buf


@pytest.mark.parametrize("w,h,pad", [(512, 384, 0), (333, 222, 5)])
def test_flush_composite_over(ctx, w, h, pad):
    """GGCUDA_COMPOSITE_OVER (row a14): the scene is rasterised on transparent and source-overed onto what the target
    already holds with the reference's byte formula -- bit-exact against that formula applied to the CUDA scene pixels,
    and the scene pixels within 1/255 of the oracle's. Row-padded target: the padding survives."""
    elems = U.random_scene(31, w, h, 150, clips=True)
    plain = U.gpu_scene(ctx, elems, w, h, bg=(0, 0, 0, 0)).copy()
    oc = U.oracle_scene(elems, w, h)
    _, ref = oc.fine((0, 0, 0, 0), straight=False, premul=True)
    mx, _, frac = U.pixel_diff(plain, ref)
    assert mx <= 1 and frac <= 0.002
    buf = _busy_target(w, h, stride_pad=pad)
    before = buf.copy()
    ctx.begin(w, h)
    ctx.set_background((9, 9, 9, 9))   # must be ignored under composite-over
    ctx.set_band(0, (h + 15) // 16)
    for e in elems:
        if e["type"] == "draw":
            ctx.fill_path(e["verbs"], e["coords"], e["color"], 1 if e.get("even_odd") else 0)
        elif e["type"] == "begin_clip":
            ctx.push_clip(e["verbs"], e["coords"])
        else:
            ctx.pop()
    view = buf[:, :w]
    ctx.L.ggcuda_flush(ctx.h, buf.ctypes.data_as(__import__("ctypes").c_void_p), buf.strides[0], G.COMPOSITE_OVER)
    want = composite_over_bytes(plain, before[:, :w])
    assert (view == want).all(), f"{int((view != want).any(axis=2).sum())} px differ from compositeOver(scene, target)"
    assert (buf[:, w:] == 0x5A).all()
    # and against the oracle's scene: at most the 1/255 the scene pixels themselves may differ by
    want_o = composite_over_bytes(ref, before[:, :w])
    mx, _, frac = U.pixel_diff(view, want_o)
    assert mx <= 1 and frac <= 0.002


def test_composite_over_after_buffer_growth():
    """Composite-over through a cold context: the grow-and-retry passes must not composite twice."""
    w, h = 400, 300
    elems = U.random_scene(33, w, h, 400)
    c = G.Context(0)
    try:
        plain = U.gpu_scene(G.Context(0), elems, w, h)
        buf = _busy_target(w, h, seed=5)
        before = buf.copy()
        c.begin(w, h)
        for e in elems:
            c.fill_path(e["verbs"], e["coords"], e["color"], 1 if e.get("even_odd") else 0)
        c.flush(buf, flags=G.COMPOSITE_OVER)
        assert c.stats()["passes"] >= 1
        assert (buf == composite_over_bytes(plain, before)).all()
    finally:
        c.close()


def test_accelerator_fill_stroke_flush():
    """The GPUAccelerator mirror end to end (accelerator.go:104-140): FillPath / StrokePath accumulate, Flush composites
    over a target that already holds content (what a gg.Context user hits, vello_accelerator.go:197-386)."""
    from gg_b200 import accelerator as A
    w, h = 320, 240
    acc = A.CUDAAccelerator(0)
    acc.Init()
    try:
        assert acc.CanAccelerate(A.AccelFill) and acc.CanAccelerate(A.AccelStroke) and not acc.CanAccelerate(A.AccelCircleSDF)
        buf = _busy_target(w, h, seed=9)
        before = buf.copy()
        tgt = A.GPURenderTarget(w, h, buf)
        rng = np.random.default_rng(2)
        for i in range(60):
            p = A.Path()
            if i % 3 == 0:
                p.Circle(rng.uniform(0, w), rng.uniform(0, h), rng.uniform(4, 50))
            else:
                p.MoveTo(rng.uniform(0, w), rng.uniform(0, h))
                for _ in range(3):
                    p.CubicTo(*rng.uniform(0, w, 6) * np.array([1, h / w] * 3))
                if i % 2:
                    p.Close()
            paint = A.Paint(color=(*rng.uniform(0, 1, 3), rng.uniform(0.3, 1.0)), fill_rule=int(i % 5 == 0),
                            line_width=float(rng.uniform(1, 9)), line_cap=int(rng.integers(0, 3)), line_join=int(rng.integers(0, 3)))
            if i % 4 == 1:
                acc.StrokePath(tgt, p, paint)
            else:
                acc.FillPath(tgt, p, paint)
        with pytest.raises(A.ErrFallbackToCPU):
            acc.StrokePath(tgt, A.Path().MoveTo(0, 0).LineTo(5, 5), A.Paint(dashed=True))
        with pytest.raises(A.ErrFallbackToCPU):
            acc.FillShape(tgt, None, A.Paint())
        acc.Flush(tgt)
        assert (buf != before).any()
        # oracle: the packed scene of that flush, rendered on transparent, composited with the reference's formula
        oc = U.oracle_from_ctx(acc.ctx, w, h)
        rep = U.compare_stages_fast(acc.ctx, oc, w, h)
        assert rep["lines"] > 0
        _, ref = oc.fine((0, 0, 0, 0), straight=False, premul=True)
        want = composite_over_bytes(ref, before)
        mx, _, frac = U.pixel_diff(buf, want)
        assert mx <= 1 and frac <= 0.002, (mx, frac)
        # a second Flush without new draws must leave the target alone
        again = buf.copy()
        acc.Flush(tgt)
        assert (buf == again).all()
    finally:
        acc.Close()


def test_render_encoding_composite_over(ctx):
    """RenderEncoding(composite_over=True): layers with wiping blend modes (Clear, Copy, SrcIn ...) act on the SCENE only, the
    finished scene is then composited over the target as the reference does -- earlier content outside the scene's pixels stays."""
    from gg_b200 import accelerator as A, scenes
    enc, w, h = scenes.config3(n=200, w=480, h=320, layer_every=10)
    acc = A.CUDAAccelerator(0)
    acc.Init()
    try:
        t0 = A.GPURenderTarget(w, h)
        acc.RenderEncoding(t0, enc)
        plain = t0.Data.copy()
        buf = _busy_target(w, h, seed=11)
        before = buf.copy()
        acc.RenderEncoding(A.GPURenderTarget(w, h, buf), enc, composite_over=True)
        assert (buf == composite_over_bytes(plain, before)).all()
    finally:
        acc.Close()


def test_even_odd_hole_larger_than_a_tile(ctx):
    """two nested same-direction squares filled EvenOdd. The hole's interior tiles have backdrop 2 and no
    segments: the reference's coarse (coarse.go:425) paints them solid; here (and in the oracle) they stay empty."""
    w = h = 128
    sq = lambda a, b: [a, a, b, a, b, b, a, b]
    verbs = np.array([0, 1, 1, 1, 4, 0, 1, 1, 1, 4], dtype=np.uint8)
    coords = np.array(sq(4, 124) + sq(24, 104), dtype=np.float64)
    elems = [dict(type="draw", verbs=verbs, coords=coords, color=(255, 0, 0, 255), even_odd=True)]
    out, oc, _ = _check(ctx, elems, w, h)
    assert out[64, 64, 3] == 0 and out[40, 40, 3] == 0      # inside the hole: interior tile and edge tile
    assert out[10, 10, 3] == 255 and out[64, 10, 3] == 255  # the ring
    # non-zero: the same geometry is solid throughout
    elems[0]["even_odd"] = False
    out, _, _ = _check(ctx, elems, w, h)
    assert out[64, 64, 3] == 255


def test_float_target(ctx):
    """GGCUDA_TARGET_F32: premultiplied float RGBA in device memory (the north star's 128-bit f32 store variant);
    quantised with fine.wgsl's (v * 255 + 0.5) it is the RGBA8 frame."""
    import torch
    from gg_b200 import scenes
    enc, w, h = scenes.config3(n=200, w=500, h=300, layer_every=25)
    ref8 = U.gpu_encoding(ctx, enc, w, h).copy()
    f = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    ctx.begin(w, h)
    ctx.set_band(0, (h + 15) // 16)
    ctx.add_encoding(*enc.streams())
    ctx.render_device(f.data_ptr(), w * 16, G.TARGET_F32)
    torch.cuda.synchronize()
    q = (f.clamp(0, 1) * 255.0 + 0.5).to(torch.uint8).cpu().numpy()
    assert (q == ref8).all()


def test_resident_scene_and_dirty_rect():
    """a scene whose key is still resident skips ingest / upload / flatten / binning / coarse (fine only), a dirty
    rectangle re-rasterises and reads back only the tiles it touches."""
    from gg_b200 import scenes
    enc, w, h = scenes.config3(n=300, w=640, h=480, layer_every=20)
    c = G.Context(0)
    try:
        key = enc.CacheKey()
        assert not c.begin_keyed(w, h, key)
        c.add_encoding(*enc.streams())
        a = np.zeros((h, w, 4), dtype=np.uint8)
        c.flush(a)
        cold = c.stats()["kernel_launches"]
        assert c.begin_keyed(w, h, key)                      # resident
        with pytest.raises(G.GGCudaError):
            c.add_encoding(*enc.streams())                   # nothing may be added to a resident scene
        b = np.zeros_like(a)
        c.flush(b)
        assert c.stats()["kernel_launches"] == 1 < cold and (a == b).all()
        # dirty rectangle: only the tiles it touches are written
        assert c.begin_keyed(w, h, key)
        c.set_dirty_rect(100, 50, 230, 140)
        d = np.full_like(a, 7)
        c.flush(d)
        x0, x1, y0, y1 = 96, 256, 48, 144                    # tile pairs are 32 px wide
        assert (d[y0:y1, x0:x1] == a[y0:y1, x0:x1]).all()
        d[y0:y1, x0:x1] = 7
        assert (d == 7).all()
        # another key: not resident, full pipeline again
        assert not c.begin_keyed(w, h, key + 1)
        c.add_encoding(*enc.streams())
        c.flush(b)
        assert 1 < c.stats()["kernel_launches"] <= cold and (a == b).all()
    finally:
        c.close()


def test_registered_target_matches_staged_flush():
    """ggcuda_register_target: the page-locked, slice-by-slice read-back gives the same frame as the staged copy."""
    w, h = 700, 1100
    elems = U.random_scene(41, w, h, 300)
    c = G.Context(0)
    try:
        staged = U.gpu_scene(c, elems, w, h).copy()
        pinned = np.zeros((h, w, 4), dtype=np.uint8)
        c.register_target(pinned)
        c.flush(pinned, flags=G.KEEP_SCENE)
        assert (pinned == staged).all()
        before = pinned.copy()
        c.flush(pinned, flags=G.KEEP_SCENE | G.COMPOSITE_OVER)      # composite-over reads the registered target too
        assert (pinned == composite_over_bytes(staged, before)).all()
        c.unregister_target(pinned)
        c.flush(pinned, flags=G.KEEP_SCENE)
        assert (pinned == staged).all()
    finally:
        c.close()


def test_deterministic_frames(ctx):
    """Two renders of the same scene give the same bytes (round 1 accepted a difference of 1: float atomics in fine)."""
    from gg_b200 import scenes
    enc, w, h = scenes.config3(n=600, w=1024, h=768, layer_every=30)
    a = U.gpu_encoding(ctx, enc, w, h).copy()
    for _ in range(3):
        assert (U.gpu_encoding(ctx, enc, w, h) == a).all()


def test_banded_curved_fills_match_full_frame(ctx):
    """curved fills straddling band edges -- bands of ragged height against the full frame, with the
    ingest-time culling of paths outside the band active."""
    from gg_b200 import scenes
    enc, w, h = scenes.config1()
    full = U.gpu_encoding(ctx, enc, w, h).copy()
    n_full = ctx.stats()["n_draws"]
    ht = (h + 15) // 16
    acc = np.zeros_like(full)
    culled = 0
    for (y0, y1) in [(0, 3), (3, 4), (4, 17), (17, ht)]:
        part = U.gpu_encoding(ctx, enc, w, h, band=(y0, y1))
        culled += n_full - ctx.stats()["n_draws"]
        acc[y0 * 16:y1 * 16] = part[y0 * 16:y1 * 16]
    assert culled > 0, "ingest did not drop any path outside its band"
    assert (acc == full).all()


def test_config1_cuda_vs_gg_cpu_path(ctx):
    """Row a15 on a BASELINE config: the CUDA frame of configs[0] (512x512, 1 000 SrcOver fills) against gg's own CPU path
    restated in oracle/aaa.c (pinned diff == 0 to the reference's AAA goldens, tests/test_cpu_aaa.py). The north star asks
    max 2/255 and mean 0.25/255; an exact-area rasteriser over RGBA8-packed brushes does not get there
    -- the measured distance is recorded here and guarded against regressions."""
    import test_cpu_aaa as TA
    from gg_b200 import scenes
    enc, w, h = scenes.config1()
    out = U.gpu_encoding(ctx, enc, w, h)
    cpu = TA.gg_cpu_render(enc, w, h)
    d = np.abs(out.astype(int) - cpu.astype(int))
    mean, mx, beyond = float(d.mean()), int(d.max()), float((d.max(axis=2) > 2).mean())
    print(f"config1 CUDA vs gg CPU (AAA + truncating source-over): mean |d| = {mean:.3f}/255, max = {mx}, {beyond * 100:.2f}% of pixels beyond 2/255")
    assert mean < 2.0 and beyond < 0.2
    # a single opaque shape on an empty canvas: no overlap, no truncation chain -- only the rim differs: the tile pipeline
    # flattens to 0.25 px (flatten.go:19; chords up to a quarter pixel inside the arc = up to 64/255 where the rim is nearly
    # horizontal), gg's CPU edges subdivide to 0.1 px and then run forward differences (its own two modes differ by up to 60,
    # circle_render_test.go:731)
    from gg_b200 import scene as S
    sc = S.Scene()
    sc.Fill(S.FillNonZero, S.IDENTITY, (0.2, 0.4, 0.8, 1.0), S.circle_verbs_coords(256.0, 256.0, 200.0))
    enc1 = sc.Encoding()
    out1 = U.gpu_encoding(ctx, enc1, w, h)
    cpu1 = TA.gg_cpu_render(enc1, w, h)
    d1 = np.abs(out1.astype(int) - cpu1.astype(int))
    print(f"one circle r=200: mean |d| = {d1.mean():.4f}/255, max = {d1.max()}, {(d1.max(axis=2) > 2).mean() * 100:.3f}% beyond 2/255")
    assert d1.mean() < 0.3 and (d1.max(axis=2) > 2).mean() < 0.01


def test_gradient_brushes(ctx):
    """linear / radial / sweep / focal-radial gradient fills (pad, repeat, reflect) through the per-draw entry, mixed with solid fills and
    clips: PTCL word for word (CmdGrad where the oracle has it), pixels within 1/255 of the oracle, which evaluates gg's
    ColorAt exactly (the device does the same arithmetic but for powf in the linear -> sRGB step)."""
    w, h = 400, 300
    rng = np.random.default_rng(8)
    ctx.begin(w, h)
    ctx.set_background((0, 0, 0, 0))
    ctx.set_band(0, (h + 15) // 16)
    n_grad = 0
    for i in range(60):
        v, c = U.circle_path(np.float32(rng.uniform(0, w)), np.float32(rng.uniform(0, h)), np.float32(rng.uniform(10, 90)))
        if i % 7 == 3:
            ctx.push_clip(v, c)
            continue
        if i % 7 == 6:
            ctx.pop()
        if i % 2:
            stops = [(float(o), *rng.uniform(0, 1, 3), float(rng.uniform(0.3, 1.0))) for o in sorted(rng.uniform(0, 1, int(rng.integers(2, 5))))]
            if i % 8 == 1:
                ctx.fill_path_gradient(v, c, 0, tuple(rng.uniform(0, w, 4)), stops, extend=i % 3, fill_rule=0)
            elif i % 8 == 3:
                ctx.fill_path_gradient(v, c, 1, (float(rng.uniform(0, w)), float(rng.uniform(0, h)), 0.0, float(rng.uniform(20, 150))), stops, extend=i % 3)
            elif i % 8 == 5:     # sweep: centre, start / end angle
                ctx.fill_path_gradient(v, c, 2, (float(rng.uniform(0, w)), float(rng.uniform(0, h)), float(rng.uniform(-3, 3)), float(rng.uniform(-6, 6))), stops, extend=i % 3)
            else:                # radial with the focus off the centre
                gx, gy, r1 = float(rng.uniform(0, w)), float(rng.uniform(0, h)), float(rng.uniform(40, 150))
                ctx.fill_path_gradient(v, c, 3, (gx, gy, 5.0, r1, gx + 0.4 * r1, gy - 0.3 * r1), stops, extend=i % 3)
            n_grad += 1
        else:
            ctx.fill_path(v, c, tuple(int(x) for x in rng.integers(0, 256, 4)), 0)
    out = np.zeros((h, w, 4), dtype=np.uint8)
    ctx.flush(out, flags=G.KEEP_SCENE)
    oc = U.oracle_from_ctx(ctx, w, h)
    rep = U.compare_stages_fast(ctx, oc, w, h)
    assert n_grad > 20 and (oc.ptcl_words == 6).sum() > 0 and rep["ptcl_words"] > 0
    words = ctx.debug_read(G.BUF_SCENE, np.uint32)
    lay = ctx.debug_read(G.BUF_LAYOUT, G.LAYOUT)[0]
    from oracle import twin as T
    ref, _ = T.render_packed(words, lay, w, h, (0, 0, 0, 0), 4)
    d = np.abs(out.astype(int) - ref.astype(int))
    assert d.max() <= 2 and (d.max(axis=2) > 1).mean() < 1e-4, (d.max(), (d.max(axis=2) > 1).mean(), (d.max(axis=2) > 0).mean())


@pytest.mark.gpu
def test_round_rect_sdf_encoding(ctx):
    """TagFillRoundRect in a scene.Encoding renders as the reference's SDF coverage (scene/renderer.go:986-1043)
    inside the inflated outline's tiles; mixed with path fills, a clip and a rotated round rect (outline fallback). PTCL word
    for word, pixels within 1/255 of the oracle (same float32 formula on both sides, sqrt rounding aside)."""
    from gg_b200 import scene as S
    w, h = 500, 380
    rng = np.random.default_rng(21)
    enc = S.Encoding()
    for i in range(40):
        col = (*rng.uniform(0, 1, 3), float(rng.uniform(0.3, 1.0)))
        x0, y0 = rng.uniform(-20, w - 40), rng.uniform(-20, h - 40)
        rect = (x0, y0, x0 + rng.uniform(2, 200), y0 + rng.uniform(2, 160))
        if i % 9 == 4:
            enc.EncodeTransform((0.8, -0.6, 60, 0.6, 0.8, -40))
        elif i % 5 == 0:
            enc.EncodeTransform((float(rng.uniform(0.5, 1.5)), 0, float(rng.uniform(-10, 10)), 0, float(rng.uniform(0.5, 1.5)), 0))
        else:
            enc.EncodeTransform(S.IDENTITY)
        if i % 4 == 1:
            enc.EncodePath(*U.circle_path(np.float32(rng.uniform(0, w)), np.float32(rng.uniform(0, h)), np.float32(rng.uniform(10, 70))))
            enc.EncodeFill(col)
        else:
            enc.EncodeFillRoundRect(col, rect, float(rng.uniform(0, 40)), float(rng.uniform(0, 40)))
        if i == 12:
            enc.EncodeTransform(S.IDENTITY)
            enc.EncodePath(*U.circle_path(np.float32(250), np.float32(190), np.float32(170)))
            enc.EncodeBeginClip()
        if i == 30:
            enc.EncodeEndClip()
    ctx.begin(w, h)
    ctx.set_background((0, 0, 0, 0))
    ctx.set_band(0, (h + 15) // 16)      # the shared context keeps the band of the test before
    ctx.add_encoding(*enc.streams())
    out = np.zeros((h, w, 4), dtype=np.uint8)
    ctx.flush(out, flags=G.KEEP_SCENE)
    oc = U.oracle_from_ctx(ctx, w, h)
    rep = U.compare_stages_fast(ctx, oc, w, h)
    assert (oc.ptcl_words == 6).sum() > 0 and rep["ptcl_words"] > 0
    words = ctx.debug_read(G.BUF_SCENE, np.uint32)
    lay = ctx.debug_read(G.BUF_LAYOUT, G.LAYOUT)[0]
    from oracle import twin as T
    ref, _ = T.render_packed(words, lay, w, h, (0, 0, 0, 0), 4)
    d = np.abs(out.astype(int) - ref.astype(int))
    assert d.max() <= 1 and (d.max(axis=2) > 0).mean() < 0.01, (d.max(), (d.max(axis=2) > 0).mean())


@pytest.mark.gpu
def test_images_in_encoding(ctx):
    """TagImage (scene/renderer.go:1093-1243) through the encoding entry: translated (nearest texel), scaled,
    rotated and mirrored images over fills, inside a clip, under a later fill. PTCL word for word, pixels within 1/255 of the
    oracle (the same float32 arithmetic on both sides)."""
    from gg_b200 import scene as S
    w, h = 420, 300
    rng = np.random.default_rng(33)
    enc = S.Encoding()
    imgs = []
    for k, (ih, iw) in enumerate([(24, 40), (64, 64), (7, 5), (120, 90)]):
        a = rng.integers(0, 256, (ih, iw, 1)).astype(np.float32)
        a[rng.random((ih, iw, 1)) < 0.2] = 0
        a[rng.random((ih, iw, 1)) < 0.3] = 255
        rgb = rng.integers(0, 256, (ih, iw, 3)).astype(np.float32)
        imgs.append(np.concatenate([np.floor(rgb * a / 255.0), a], axis=2).astype(np.uint8))
        enc.AddImage(imgs[-1])
    enc.EncodeTransform(S.IDENTITY)
    enc.EncodePath(*S.rect_verbs_coords(0, 0, w, h)); enc.EncodeFill((0.9, 0.9, 0.8, 1.0))
    enc.EncodeImage(0, (1, 0, 30, 0, 1, 20))
    enc.EncodeImage(1, (2.25, 0, 100.5, 0, 1.5, 10.25))
    enc.EncodePath(*U.circle_path(np.float32(260), np.float32(170), np.float32(90)))
    enc.EncodeBeginClip()
    enc.EncodeImage(3, (0.8, -0.6, 200, 0.6, 0.8, 60))
    enc.EncodeImage(2, (20, 0, 180, 0, 20, 120))
    enc.EncodePath(*U.circle_path(np.float32(300), np.float32(200), np.float32(40))); enc.EncodeFill((0.1, 0.2, 0.9, 0.6))
    enc.EncodeEndClip()
    enc.EncodeImage(1, (-1, 0, 90, 0, 1, 200))
    enc.EncodeImage(0, (1, 0, -20, 0, 1, 280))          # partly off the canvas
    ctx.begin(w, h)
    ctx.set_background((0, 0, 0, 0))
    ctx.set_band(0, (h + 15) // 16)
    for im in imgs:
        ctx.add_image(im)
    ctx.add_encoding(*enc.streams())
    out = np.zeros((h, w, 4), dtype=np.uint8)
    ctx.flush(out, flags=G.KEEP_SCENE)
    oc = U.oracle_from_ctx(ctx, w, h)
    rep = U.compare_stages_fast(ctx, oc, w, h)
    assert (oc.ptcl_words == 6).sum() > 0 and rep["ptcl_words"] > 0
    words = ctx.debug_read(G.BUF_SCENE, np.uint32)
    lay = ctx.debug_read(G.BUF_LAYOUT, G.LAYOUT)[0]
    from oracle import twin as T
    ref, _ = T.render_packed(words, lay, w, h, (0, 0, 0, 0), 4)
    d = np.abs(out.astype(int) - ref.astype(int))
    assert d.max() <= 1 and (d.max(axis=2) > 0).mean() < 0.01, (d.max(), (d.max(axis=2) > 0).mean())
    assert (out[20:44, 30:70] != np.array([230, 230, 204, 255])).any()      # image 0 is there
    # through the accelerator mirror (RenderEncoding registers the encoding's images), resident on the second call
    from gg_b200 import accelerator as A
    acc = A.CUDAAccelerator(); acc.Init()
    tgt = A.GPURenderTarget(w, h)
    acc.RenderEncoding(tgt, enc)
    acc.RenderEncoding(tgt, enc)
    assert acc.resident_hits == 1 and (tgt.Data == out).all()
    acc.Close()


@pytest.mark.gpu
def test_graph_replay_no_wait_and_band_broadcast(ctx):
    """The steady-state conveniences of round 2 on one device: a pass into a device target is replayed as a CUDA graph (two
    cached, for a double-buffered target), GGCUDA_NO_WAIT returns before the device is done once the scene's counts are known
    to fit, and ggcuda_broadcast_band copies a finished band on a side stream (here into two 'frames' of the same device) while
    the next frame is rasterised. Every frame must equal the plain render bit for bit; changing what a graph was built from
    (background, target, dirty rectangle) must not replay a stale one."""
    import torch
    from gg_b200 import scenes
    w, h = 640, 400
    enc, _, _ = scenes.config3(n=300, w=w, h=h, layer_every=5, nowipe=True)   # no wiping layers: the background shows
    stream = torch.cuda.Stream()
    side = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    try:
        ctx.begin(w, h)
        ctx.set_background((10, 20, 30, 255))
        ctx.set_band(0, (h + 15) // 16)
        ctx.add_encoding(*enc.streams())
        truth = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
        ctx.set_timing(True)                                   # events between the stages: the stream path, no graph
        ctx.render_device(truth.data_ptr(), w * 4, G.KEEP_SCENE)
        ctx.set_timing(False)
        assert int(truth.sum().item()) > 0
        bands_ = [torch.zeros_like(truth), torch.zeros_like(truth)]
        frames = [torch.zeros_like(truth), torch.zeros_like(truth)]
        for k in range(6):                                     # queued behind one another
            b = bands_[k & 1]
            ctx.render_device(b.data_ptr(), w * 4, G.KEEP_SCENE | G.NO_WAIT)
            ctx.broadcast_band(b.data_ptr(), [f.data_ptr() for f in frames], b.numel(), side.cuda_stream)
        ctx.sync()
        torch.cuda.synchronize()
        for t in bands_ + frames:
            assert bool((t == truth).all().item())
        # a different background: same buffers, same target -> the configuration differs, the graph must be rebuilt
        ctx.set_background((200, 0, 0, 255))
        other = torch.zeros_like(truth)
        ctx.render_device(other.data_ptr(), w * 4, G.KEEP_SCENE)
        ctx.render_device(bands_[0].data_ptr(), w * 4, G.KEEP_SCENE)
        torch.cuda.synchronize()
        assert bool((bands_[0] == other).all().item()) and not bool((other == truth).all().item())
        # a dirty rectangle applies to one render and is part of what a graph is keyed by
        ctx.set_background((10, 20, 30, 255))
        ctx.set_dirty_rect(64, 48, 200, 120)
        ctx.render_device(other.data_ptr(), w * 4, G.KEEP_SCENE)
        torch.cuda.synchronize()
        inside = other[48:112, 64:192]
        assert bool((inside == truth[48:112, 64:192]).all().item()) and not bool((other == truth).all().item())
        ctx.render_device(other.data_ptr(), w * 4, G.KEEP_SCENE)
        torch.cuda.synchronize()
        assert bool((other == truth).all().item())
    finally:
        ctx.sync()
        ctx.set_stream(0)


@pytest.mark.gpu
def test_pipelined_renderer_frames_in_flight():
    """Two contexts, two host threads (accelerator.PipelinedRenderer): different encodings submitted back to back come out as
    the same pixels as the one-at-a-time accelerator renders them."""
    from gg_b200 import accelerator as A, scenes
    w, h = 320, 240
    encs = [scenes.config3(n=150, w=w, h=h, layer_every=4, seed=s)[0] for s in (1, 2, 3, 4, 5)]
    acc = A.CUDAAccelerator(); acc.Init()
    want = []
    for e in encs:
        t = A.GPURenderTarget(w, h)
        acc.RenderEncoding(t, e, resident=False)
        want.append(t.Data.copy())
    acc.Close()
    pr = A.PipelinedRenderer(0, depth=2)
    tgts = [A.GPURenderTarget(w, h) for _ in encs]
    futs = [pr.submit(t, e, resident=False) for t, e in zip(tgts, encs)]
    for f in futs:
        f.result()
    pr.Close()
    for t, wnt in zip(tgts, want):
        assert (t.Data == wnt).all()
    assert not (want[0] == want[1]).all()
