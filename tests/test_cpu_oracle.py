"""CPU tests (no GPU): the oracle against the reference's own golden images and known-answer vectors."""
import json
import os

import numpy as np
import pytest
from PIL import Image

from oracle import twin as T

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "vello-gpu-pipeline")
STAR = [(50, 10), (75, 90), (10, 40), (90, 40), (25, 90)]
WHITE, LIME, MAROON, BLACK, BLUE = (255, 255, 255, 255), (0, 255, 0, 255), (128, 0, 0, 255), (0, 0, 0, 255), (0, 0, 255, 255)

# tilecompute/rasterizer_test.go:39-135 (scene definitions and thresholds in % of differing pixels)
CASES = [
    ("filled_circle", 100, LIME, WHITE, 0, lambda: T.flatten_fill(T.circle_cubics(50, 50, 45)), 0.0),
    ("filled_triangle", 100, LIME, WHITE, 0, lambda: T.polygon_lines([(5, 5), (95, 50), (5, 95)]), 0.0),
    ("filling_nonzero_rule", 100, MAROON, WHITE, 0, lambda: T.polygon_lines(STAR), 0.15),
    ("filling_evenodd_rule", 100, MAROON, WHITE, 1, lambda: T.polygon_lines(STAR), 0.15),
    ("smoke_filled_circle", 20, BLUE, BLACK, 0, lambda: T.flatten_fill(T.circle_cubics(10, 10, 7)), 0.0),
    ("smoke_filled_square", 20, BLUE, BLACK, 0, lambda: T.polygon_lines([(7, 7), (13, 7), (13, 13), (7, 13)]), 0.0),
]


@pytest.mark.parametrize("name,size,color,bg,eo,lines,thr", CASES, ids=[c[0] for c in CASES])
def test_vello_golden_images(name, size, color, bg, eo, lines, thr):
    """TestVelloPortVsGPUPipeline / ...Smoke: RasterizeScene and the PTCL pipeline both reproduce Vello's CPU goldens."""
    ref = np.array(Image.open(os.path.join(GOLDEN, name + ".png")).convert("RGBA"))
    e, l = T.make_elements([dict(lines=lines(), color=color, even_odd=eo)])
    per_path = T.rasterize_scene(bg, e, l, size, size)
    ptcl_straight, _ = T.Coarse(e, l, size, size).fine(bg)
    for out in (per_path, ptcl_straight):
        pct = (out != ref).any(axis=2).mean() * 100
        assert pct <= thr, f"{name}: {pct:.3f}% pixels differ (threshold {thr})"
    if thr == 0.15:   # the reference's comments record exactly 8 / 10 differing pixels against the Rust output
        assert (per_path != ref).any(axis=2).sum() == {"filling_nonzero_rule": 8, "filling_evenodd_rule": 10}[name]


def test_flatten_line_counts():
    """lines per 4-cubic circle grow with sqrt(r)."""
    got = {r: len(T.flatten_fill(T.circle_cubics(200, 200, r))) for r in (4, 20, 45, 100)}
    assert got == {4: 12, 20: 20, 45: 32, 100: 48}


def test_path_monoid_known_answers():
    """scene_encode_test.go:239-310 TestPathMonoidNew."""
    LINETO, PATH, TRANSFORM, STYLE = 0x9, 0x10, 0x20, 0x40
    kat = [
        (0, dict()),
        (LINETO, dict(path_seg_ix=1, path_seg_offset=2)),
        (TRANSFORM, dict(trans_ix=1)),
        (STYLE, dict(style_ix=1)),
        (PATH, dict(path_ix=1)),
        (LINETO | (LINETO << 8), dict(path_seg_ix=2, path_seg_offset=4)),
        (TRANSFORM | (STYLE << 8) | (LINETO << 16) | (PATH << 24), dict(trans_ix=1, style_ix=1, path_seg_ix=1, path_seg_offset=2, path_ix=1)),
        # ggcuda's MoveTo tag 0x0C: two floats of path data, no segment, under the unchanged bit tricks
        (0x0C, dict(path_seg_offset=2)),
        (0x0C | (0x0B << 8) | (0x0A << 16) | (LINETO << 24), dict(path_seg_ix=3, path_seg_offset=2 + 6 + 4 + 2)),
    ]
    for word, want in kat:
        m = T.path_monoid(word)
        for f in ("trans_ix", "path_seg_ix", "path_seg_offset", "style_ix", "path_ix"):
            assert int(m[f]) == want.get(f, 0), (hex(word), f)


def test_draw_monoid_known_answers():
    """scene_encode_test.go:418-466 TestDrawMonoidNew."""
    for tag, want in [(0, (0, 0, 0, 0)), (0x44, (1, 0, 1, 1)), (0x9, (1, 1, 2, 0)), (0x21, (1, 1, 0, 0))]:
        m = T.draw_monoid(tag)
        assert (int(m["path_ix"]), int(m["clip_ix"]), int(m["scene_offset"]), int(m["info_offset"])) == want


def test_color_packing():
    """scene_encode_test.go:597 colour packing: premultiplied with +0.5 rounding, R | G<<8 | B<<16 | A<<24."""
    e, l = T.make_elements([dict(lines=T.polygon_lines([(1, 1), (9, 1), (5, 9)]), color=(255, 128, 0, 128))])
    c = T.Coarse(e, l, 16, 16)
    a = np.float32(128) / np.float32(255)
    want = int(np.float32(255) * a + np.float32(0.5)) | (int(np.float32(128) * a + np.float32(0.5)) << 8) | (128 << 24)
    assert int(c.info[0]) == want


def test_ptcl_words_fill_solid_and_order():
    """coarse_test.go:14-165, 218-259, 344-395: Fill/Solid/Color encoding, draw order, Solid vs Fill tiles."""
    big = T.polygon_lines([(0, 0), (64, 0), (64, 64), (0, 64)])          # covers 4x4 tiles
    tri = T.polygon_lines([(20, 20), (44, 24), (22, 44)])
    e, l = T.make_elements([dict(lines=big, color=(255, 0, 0, 255)), dict(lines=tri, color=(0, 0, 255, 255), even_odd=1)])
    c = T.Coarse(e, l, 64, 64)
    inner = c.ptcl(1 * 4 + 1)                                            # tile (1,1): interior of the square, crossed by the triangle
    assert inner[0] == 0                                                 # blend offset word
    assert inner[1] == 3 and inner[2] == 5 and inner[3] == 0xFF0000FF    # CmdSolid, CmdColor red
    assert inner[4] == 1 and (inner[5] & 1) == 1 and inner[8] == 5 and inner[9] == 0xFFFF0000   # CmdFill even-odd, CmdColor blue
    assert inner[-1] == 0                                                # CmdEnd
    edge = c.ptcl(0)                                                     # tile (0,0): square's corner has segments -> CmdFill
    assert edge[1] == 1 and (edge[2] >> 1) >= 1


def test_fine_tile_known_answers():
    """fine_ptcl_test.go:29-189: solid colour, fill + colour, two shapes composited."""
    def words(*w):
        return np.array([0, *w], dtype=np.uint32)
    out = T.fine_tile(words(3, 5, 0xFF0000FF, 0), np.zeros(0, dtype=T.SEGMENT), (0, 0, 0, 1))
    assert np.allclose(out, [1, 0, 0, 1], atol=1e-4)
    segs = np.array([((2, 2), (14, 8), 1e9), ((14, 8), (2, 14), 1e9), ((2, 14), (2, 2), 1e9)], dtype=T.SEGMENT)
    out = T.fine_tile(words(1, 3 << 1, 0, 0, 5, 0xFF00FF00, 0), segs, (1, 1, 1, 1)).reshape(16, 16, 4)
    assert out[8, 6, 1] > 0.5 and out[8, 6, 0] < 0.1                     # inside: green
    assert (out[0, 0, :3] > 0.9).all() and (out[1, 15, :3] > 0.9).all()  # outside: white
    # solid red then 50 % blue: source-over in premultiplied float
    out = T.fine_tile(words(3, 5, 0xFF0000FF, 3, 5, 0x80800000, 0), np.zeros(0, dtype=T.SEGMENT), (0, 0, 0, 0))
    a = 128 / 255
    assert np.allclose(out[0], [1 - a, 0, a, 1.0], atol=2e-3)


def test_fine_clip_depths():
    """fine_clip_test.go:14-160: clip depth 1, 2 and > 4 (spill) give saved*(1-fg.a)+fg with fg = rgba*area*alpha."""
    empty = np.zeros(0, dtype=T.SEGMENT)
    one = np.float32(1.0).view(np.uint32)
    half = np.float32(0.5).view(np.uint32)
    w = np.array([0, 10, 3, 5, 0xFF0000FF, 3, 11, 0x8003, int(half), 0], dtype=np.uint32)   # Begin, solid red, Solid, End(alpha .5)
    out = T.fine_tile(w, empty, (0, 0, 1, 1))
    assert np.allclose(out[0], [0.5, 0, 0.5, 1.0], atol=1e-5)
    deep = [0] + [10] * 6 + [3, 5, 0xFF00FF00] + [3, 11, 0x8003, int(one)] * 6 + [0]
    out = T.fine_tile(np.array(deep, dtype=np.uint32), empty, (0, 0, 0, 1))
    assert np.allclose(out[0], [0, 1, 0, 1], atol=1e-5)


def test_fine_tile_clip_vector():
    """fine_ptcl_test.go:150-189 TestFineRasterizeTileClip: Solid, BeginClip, Solid + green, EndClip(alpha .5) over white gives
    (0.5, 1, 0.5, 1). The reference writes blend word 0 and its fine ignores the word; here the word selects the composite
    and source-over is 0x8003 (clip) or 3 (Normal layer) -- word 0 is scene.BlendClear."""
    half = int(np.float32(0.5).view(np.uint32))
    empty = np.zeros(0, dtype=T.SEGMENT)
    for blend, want in ((0x8003, [0.5, 1, 0.5, 1]), (3, [0.5, 1, 0.5, 1]), (0, [0, 0, 0, 0])):
        w = np.array([0, 3, 10, 3, 5, 0xFF00FF00, 11, blend, half, 0], dtype=np.uint32)
        assert np.allclose(T.fine_tile(w, empty, (1, 1, 1, 1))[0], want, atol=1e-6), hex(blend)


def test_clip_scene_zero_depth_and_bbox():
    """clip_integration_test.go:24-326: content outside the clip's bbox or in its empty tiles is suppressed, inside it is kept."""
    clip = T.polygon_lines([(16, 16), (48, 16), (48, 48), (16, 48)])
    big = T.polygon_lines([(0, 0), (64, 0), (64, 64), (0, 64)])
    e, l = T.make_elements([dict(type=T.ELEM_BEGIN_CLIP, lines=clip, blend=0x8003, alpha=1.0), dict(lines=big, color=(255, 0, 0, 255)),
                            dict(type=T.ELEM_END_CLIP)])
    c = T.Coarse(e, l, 64, 64)
    assert list(c.ptcl(0)) == [0, 0]                                     # outside the clip bbox: nothing
    inside = list(c.ptcl(1 * 4 + 1))
    assert inside[1] == 10 and 11 in inside                              # BeginClip ... EndClip
    s, _ = c.fine((255, 255, 255, 255))
    assert tuple(s[32, 32]) == (255, 0, 0, 255) and tuple(s[4, 4]) == (255, 255, 255, 255)


def test_blend_byte_known_answers():
    """internal/blend/{porter_duff,advanced,hsl}_test.go table rows (tests/golden/blend_kats.json)."""
    import ctypes as C
    L = T.lib()
    L.ot_blend_bytes.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    kats = json.load(open(os.path.join(HERE, "golden", "blend_kats.json")))
    assert len(kats) >= 40
    for k in kats:
        s, d, o = np.array(k["s"], np.uint8), np.array(k["d"], np.uint8), np.zeros(4, np.uint8)
        L.ot_blend_bytes(k["mode"], s.ctypes.data, d.ctypes.data, o.ctypes.data)
        assert list(o) == k["want"], (k["func"], k["case"], list(o))


def test_float_blend_tracks_byte_blend():
    """The float32 layer composite (ot_blend_f32, what fine applies) stays within 2/255 of gg's byte functions for
    opaque inputs -- except Overlay / HardLight where the reference's `2*d` wraps at d == 128 (blend_funcs.go:170,223)."""
    import ctypes as C
    L = T.lib()
    L.ot_blend_bytes.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ot_blend_f32.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(0)
    for mode in range(29):
        word = (mode << 8) | 3 if mode < 16 else mode - 16
        for _ in range(400):
            s = np.array([*rng.integers(0, 256, 3), 255], dtype=np.uint8)
            d = np.array([*rng.integers(0, 256, 3), 255], dtype=np.uint8)
            if mode in (3, 8) and (128 in s[:3] or 128 in d[:3]):
                continue
            o = np.zeros(4, np.uint8)
            L.ot_blend_bytes(mode, s.ctypes.data, d.ctypes.data, o.ctypes.data)
            bf, ff, of = (d / 255).astype(np.float32), (s / 255).astype(np.float32), np.zeros(4, np.float32)
            L.ot_blend_f32(word, bf.ctypes.data, ff.ctypes.data, of.ctypes.data)
            assert np.abs(np.clip(of, 0, 1) * 255 - o).max() <= 2.0, (mode, s, d, o, of * 255)


# ---- exact-area coverage (this pipeline, = Vello) against gg's CPU rasteriser (Skia AAA) ----
SKIA_AAA = {   # internal/raster/analytic_filler_golden_test.go:466-560 (TestCompositing_*RGB): paths of the Skia goldens
    "polygon": [(75.160671, 88.756136), (24.797274, 88.734053), (9.255130, 40.828792), (50.012955, 11.243795), (90.744819, 40.864522)],
    "float-rect-aa": [(10.3, 15.4), (90.8, 15.4), (90.8, 86.0), (10.3, 86.0)],
    "star-aa": [(50.0, 7.5), (75.0, 87.5), (10.0, 37.5), (90.0, 37.5), (25.0, 87.5)],
}


def skia_golden_coverage(name):
    """Invert renderWithAnalyticFillerOnWhite (analytic_filler_golden_test.go:102-139) to recover AAA's 8-bit coverage
    from the golden image (several coverages can share one RGB: the mean of the candidates is used, +-1)."""
    def div255(a, b):
        return (a * b + 128) // 255
    paint = (div255(50, 200), div255(127, 200), div255(150, 200), 200)
    lut = {}
    for cov in range(256):
        if cov == 0:
            rgb = (255, 255, 255)
        else:
            sc = cov + 1
            s = [(p * sc) >> 8 for p in paint]
            inv = (255 - s[3]) + 1
            rgb = tuple((s[k] + ((255 * inv) >> 8)) & 255 for k in range(3))
        lut.setdefault(rgb, []).append(cov)
    g = np.array(Image.open(os.path.join(HERE, "golden", "skia-aaa", f"skia-aaa-{name}-white.png")).convert("RGBA"))
    cov = np.zeros(g.shape[:2])
    for y in range(g.shape[0]):
        for x in range(g.shape[1]):
            cov[y, x] = np.mean(lut[tuple(int(v) for v in g[y, x, :3])])
    return cov


@pytest.mark.parametrize("name", list(SKIA_AAA))
def test_exact_area_vs_gg_cpu_aaa(name):
    """gg's CPU filler reproduces these Skia-AAA goldens pixel for pixel (the reference's own tests assert diff == 0), so
    they ARE gg's CPU output for these paths. Exact-area coverage differs from them where AAA snaps edge Y to a quarter
    pixel: mean |d| stays below 0.6/255 but single edge pixels are up to 57/255 apart -- the north-star bound of
    max 2/255 against the CPU path is not reachable by any exact-area rasteriser; recorded here, not hidden."""
    cov_g = skia_golden_coverage(name)
    a = T.rasterize(T.polygon_lines(np.array(SKIA_AAA[name], dtype=np.float32)), 0, 100, 100)
    d = np.abs(np.clip(a, 0, 1) * 255 - cov_g)
    assert d.mean() <= 0.6
    assert (d <= 2).mean() >= 0.975
    assert d.max() <= 60          # quarter-pixel snapping of a horizontal edge: up to ~0.125 * 255 per edge, 2 edges at a corner


# ---- the reference's scene-encoding tests, one for one (tilecompute/scene_encode_test.go) ----
def _encode(paths):
    e, l = T.make_elements([dict(lines=T.polygon_lines(p) if isinstance(p, list) else p, color=c, even_odd=eo) for p, c, eo in paths])
    return T.Coarse(e, l, 100, 100)


def _tag_bytes(c):
    L = c.layout
    return c.scene[L["path_tag_base"]:L["path_data_base"]].view(np.uint8)


def test_encode_scene_single_path():
    """TestEncodeSceneSinglePath (:14-83) + TestPackScene (:172-237) + TestPackPathTags (:553-595, the four_tags vector)."""
    c = _encode([([(10, 10), (90, 50), (10, 90)], (255, 0, 0, 255), 0)])
    L = c.layout
    assert (L["n_paths"], L["n_draw_objects"]) == (1, 1)
    assert L["path_tag_base"] == 0 and (L["path_data_base"] - L["path_tag_base"]) % 256 == 0 and L["path_data_base"] >= 256
    assert list(c.scene[L["draw_tag_base"]:L["draw_data_base"]]) == [0x44]
    assert list(c.scene[L["draw_data_base"]:L["transform_base"]]) == [0xFF0000FF]
    assert list(c.scene[L["transform_base"]:L["style_base"]].view(np.float32)) == [1, 0, 0, 1, 0, 0]
    assert list(c.scene[L["style_base"]:]) == [0]
    pd = c.scene[L["path_data_base"]:L["draw_tag_base"]]
    assert L["draw_tag_base"] == L["path_data_base"] + len(pd) and L["draw_data_base"] == L["draw_tag_base"] + 1
    assert len(c.scene) == L["path_data_base"] + len(pd) + 1 + 1 + 6 + 1             # TestPackScene: total size
    assert list(pd.view(np.float32)) == [10, 10, 90, 50, 10, 90, 10, 10]          # 4 + 2 + 2 floats for the connected triangle
    assert int(c.scene[0]) == 0x09094020                                          # Transform, Style, LineTo, LineTo packed 4 per word
    assert list(_tag_bytes(c)[:7]) == [0x20, 0x40, 0x09, 0x09, 0x09, 0x09, 0x10]   # the first point travels as a LineTo too (scene_encode.go:186-215)


def test_encode_scene_multi_path_and_even_odd():
    """TestEncodeSceneMultiPath (:85-169) and TestEncodeSceneEvenOddStyle (:653-671)."""
    c = _encode([([(10, 10), (30, 10), (30, 30), (10, 30)], (255, 0, 0, 255), 0),
                 ([(40, 10), (60, 50), (40, 50)], (0, 255, 0, 255), 0),
                 ([(70, 10), (90, 10), (90, 30), (70, 30)], (0, 0, 255, 128), 1)])
    L = c.layout
    assert (L["n_paths"], L["n_draw_objects"]) == (3, 3)
    assert list(c.scene[L["draw_tag_base"]:L["draw_data_base"]]) == [0x44] * 3
    dd = c.scene[L["draw_data_base"]:L["transform_base"]]
    assert len(dd) == 3
    blue = int(dd[2])
    assert (blue & 0xFF, (blue >> 8) & 0xFF, blue >> 24) == (0, 0, 128) and 127 <= (blue >> 16) & 0xFF <= 129
    assert L["style_base"] - L["transform_base"] == 18
    assert list(c.scene[L["style_base"]:]) == [0, 0, 2]
    c = _encode([([(50, 10), (75, 90), (10, 40), (90, 40), (25, 90)], (128, 0, 0, 255), 1)])
    assert list(c.scene[c.layout["style_base"]:]) == [2]


def test_pathtag_and_draw_scans():
    """TestPathtagReduceScan (:346-415), TestDrawReduceScan (:472-550), TestSceneRoundTrip (:674-733)."""
    c = _encode([([(10, 10), (50, 50), (10, 50)], (255, 0, 0, 255), 0), ([(60, 10), (90, 10), (90, 40), (60, 40)], (0, 0, 255, 255), 0)])
    tm = c.tag_monoids
    assert tuple(tm[0]) == (0, 0, 0, 0, 0)                                        # exclusive prefix starts at the identity
    for f in ("trans_ix", "path_seg_ix", "path_ix", "style_ix"):
        assert (np.diff(tm[f].astype(np.int64)) >= 0).all()                       # monotone
    words = c.scene[c.layout["path_tag_base"]:c.layout["path_data_base"]]
    last = T.path_monoid(int(words[len(tm) - 1]))
    total = {f: int(tm[f][-1]) + int(last[f]) for f in tm.dtype.names}
    assert (total["trans_ix"], total["style_ix"], total["path_ix"]) == (2, 2, 2) and total["path_seg_ix"] == 4 + 5     # every point of a closed polygon is a LineTo tag, the first one included
    c = _encode([([(10, 10), (50, 50), (10, 50)], (255, 0, 0, 255), 0), ([(60, 10), (90, 10), (90, 40), (60, 40)], (0, 255, 0, 255), 0),
                 ([(20, 60), (80, 60), (50, 90)], (0, 0, 255, 128), 1)])
    dm = c.draw_monoids
    assert len(dm) == 3 and tuple(dm[0]) == (0, 0, 0, 0)
    assert (int(dm[1]["path_ix"]), int(dm[1]["scene_offset"])) == (1, 1) and (int(dm[2]["path_ix"]), int(dm[2]["scene_offset"])) == (2, 2)
    dd = c.scene[c.layout["draw_data_base"]:c.layout["transform_base"]]
    assert len(c.info) == 3 and list(c.info) == list(dd)
    b = int(c.info[2])
    assert (b & 0xFF, (b >> 8) & 0xFF, b >> 24) == (0, 0, 128) and 127 <= (b >> 16) & 0xFF <= 129
    c = _encode([(T.flatten_fill(T.circle_cubics(50, 50, 20)), (255, 128, 0, 255), 0), ([(10, 10), (90, 10), (90, 90), (10, 90)], (0, 128, 255, 200), 1)])
    assert len(c.draw_monoids) == 2 and len(c.info) == 2 and len(c.tag_monoids) == c.layout["path_data_base"] - c.layout["path_tag_base"]


# ---- the reference's coarse tests, one for one (tilecompute/coarse_test.go:172-441) ----
def _cmds(ptcl):
    """[(tag, payload words...)] of one tile's PTCL (ptcl.go:17-177 word layout)."""
    out, o = [], 1
    size = {0: 1, 1: 4, 3: 1, 5: 2, 10: 1, 11: 3}
    while o < len(ptcl):
        t = int(ptcl[o])
        out.append((t, *[int(w) for w in ptcl[o + 1:o + size[t]]]))
        o += size[t]
        if t == 0:
            break
    return out


def test_coarse_single_triangle_multi_path_allocation():
    """TestCoarseRasterizeSingleTriangle (:172-214), TestCoarseRasterizeMultiPath (:218-259), TestCoarseTileAllocation (:262-313)."""
    e, l = T.make_elements([dict(lines=T.polygon_lines([(5, 5), (27, 5), (16, 27)]), color=(255, 0, 0, 255))])
    c = T.Coarse(e, l, 32, 32)
    assert (c.wt, c.ht) == (2, 2)
    with_cmds = [t for t in range(4) if len(c.ptcl(t)) > 2]
    assert with_cmds and any(any(a[0] in (1, 3) and b[0] == 5 for a, b in zip(_cmds(c.ptcl(t)), _cmds(c.ptcl(t))[1:])) for t in with_cmds)
    red, blue = T.polygon_lines([(0, 0), (32, 0), (32, 32), (0, 32)]), T.polygon_lines([(0, 0), (16, 0), (16, 16), (0, 16)])
    e, l = T.make_elements([dict(lines=red, color=(255, 0, 0, 255)), dict(lines=blue, color=(0, 0, 255, 255))])
    c = T.Coarse(e, l, 32, 32)
    colors = [cmd[1] for cmd in _cmds(c.ptcl(0)) if cmd[0] == 5]
    assert colors[:2] == [0xFF0000FF, 0xFFFF0000]                                # scene order: red, then blue
    e, l = T.make_elements([dict(lines=T.polygon_lines([(2, 2), (10, 2), (10, 10), (2, 10)]), color=(255, 0, 0, 255)),
                            dict(lines=T.polygon_lines([(5, 5), (25, 5), (25, 25), (5, 25)]), color=(0, 255, 0, 255))])
    c = T.Coarse(e, l, 32, 32)
    b0, b1 = c.paths[0]["bbox"], c.paths[1]["bbox"]
    assert (b0[2] - b0[0]) * (b0[3] - b0[1]) == 1 and (b1[2] - b1[0]) * (b1[3] - b1[1]) >= 4
    assert c.paths[1]["tiles"] > c.paths[0]["tiles"]


def test_coarse_empty_solid_evenodd_nolines():
    """TestCoarseEmptyTiles (:316-340), TestCoarseBackdropSolid (:344-395), TestCoarseEvenOdd (:399-425), TestCoarseNoLines (:428-441)."""
    e, l = T.make_elements([dict(lines=T.polygon_lines([(1, 1), (10, 1), (5, 10)]), color=(255, 0, 0, 255))])
    c = T.Coarse(e, l, 64, 64)
    assert (c.wt, c.ht) == (4, 4) and list(c.ptcl(3 * 4 + 3)) == [0, 0]           # blend offset + CmdEnd
    e, l = T.make_elements([dict(lines=T.polygon_lines([(0, 0), (48, 0), (48, 48), (0, 48)]), color=(0, 255, 0, 255))])
    c = T.Coarse(e, l, 48, 48)
    assert [cmd[0] for cmd in _cmds(c.ptcl(1 * 3 + 1))] == [3, 5, 0]              # centre tile: CmdSolid, CmdColor, CmdEnd
    star = [(16, 1), (20, 14), (30, 14), (22, 22), (26, 31), (16, 25), (6, 31), (10, 22), (2, 14), (12, 14)]
    e, l = T.make_elements([dict(lines=T.polygon_lines(star), color=(128, 0, 0, 255), even_odd=1)])
    c = T.Coarse(e, l, 32, 32)
    fills = [cmd for t in range(4) for cmd in _cmds(c.ptcl(t)) if cmd[0] == 1]
    assert fills and all(f[1] & 1 for f in fills)                                # even-odd flag in every CmdFill
    e, l = T.make_elements([])
    c = T.Coarse(e, l, 32, 32)
    assert all(list(c.ptcl(t)) == [0, 0] for t in range(4))


# ---- the reference's clip scenes, one for one (tilecompute/clip_integration_test.go) ----
def _rect(x0, y0, x1, y1):
    return T.polygon_lines([(x0, y0), (x1, y0), (x1, y1), (x0, y1)])


def _clip(lines, alpha=1.0):
    return dict(type=T.ELEM_BEGIN_CLIP, lines=lines, blend=0x8003, alpha=alpha)


def test_clip_scene_integration_nested_alpha():
    """TestClipSceneIntegration (:24-75), TestClipSceneNestedClip (:78-149), TestClipSceneAlpha (:152-195): straight-alpha
    RGBA8 of RasterizeSceneDefPTCL over a white background."""
    white = (255, 255, 255, 255)
    e, l = T.make_elements([dict(lines=_rect(0, 0, 64, 64), color=(0, 255, 0, 255)), _clip(_rect(16, 16, 48, 48)),
                            dict(lines=_rect(0, 0, 64, 64), color=(255, 0, 0, 255)), dict(type=T.ELEM_END_CLIP)])
    s, _ = T.Coarse(e, l, 64, 64).fine(white)
    assert tuple(s[32, 32]) == (255, 0, 0, 255) and tuple(s[4, 4]) == (0, 255, 0, 255)
    e, l = T.make_elements([dict(lines=_rect(0, 0, 64, 64), color=(0, 0, 255, 255)), _clip(_rect(8, 8, 56, 56)),
                            dict(lines=_rect(0, 0, 64, 64), color=(0, 255, 0, 255)), _clip(_rect(20, 20, 44, 44)),
                            dict(lines=_rect(0, 0, 64, 64), color=(255, 0, 0, 255)), dict(type=T.ELEM_END_CLIP), dict(type=T.ELEM_END_CLIP)])
    s, _ = T.Coarse(e, l, 64, 64).fine(white)
    assert tuple(s[32, 32]) == (255, 0, 0, 255) and tuple(s[12, 12]) == (0, 255, 0, 255) and tuple(s[2, 2]) == (0, 0, 255, 255)
    e, l = T.make_elements([_clip(_rect(0, 0, 32, 32), alpha=0.5), dict(lines=_rect(0, 0, 32, 32), color=(255, 0, 0, 255)), dict(type=T.ELEM_END_CLIP)])
    s, _ = T.Coarse(e, l, 32, 32).fine(white)
    assert tuple(s[16, 16]) == (255, 128, 128, 255)      # white * 0.5 + red * 0.5


def test_clip_scene_def_encoding_and_fixup():
    """TestEncodeSceneDefBasic (:220-281), TestClipLeafScan (:198-218): draw tags Color / BeginClip / Color / EndClip, the clip
    pair counted as two clip inputs, blend word and alpha bits in the draw data, and the EndClip monoid patched with its
    BeginClip's path index and scene offset."""
    e, l = T.make_elements([dict(lines=_rect(0, 0, 64, 64), color=(0, 255, 0, 255)), _clip(_rect(16, 16, 48, 48)),
                            dict(lines=_rect(0, 0, 64, 64), color=(255, 0, 0, 255)), dict(type=T.ELEM_END_CLIP)])
    c = T.Coarse(e, l, 64, 64)
    L = c.layout
    assert list(c.scene[L["draw_tag_base"]:L["draw_data_base"]]) == [0x44, 0x9, 0x44, 0x21]
    assert L["n_draw_objects"] == 4 and L["n_clips"] == 2
    dd = c.scene[L["draw_data_base"]:L["transform_base"]]
    assert int(dd[1]) == 0x8003 and dd[2:3].view(np.float32)[0] == 1.0
    dm = c.draw_monoids
    assert int(dm[3]["path_ix"]) == int(dm[1]["path_ix"]) and int(dm[3]["scene_offset"]) == int(dm[1]["scene_offset"])   # clip_leaf.go:48-51


def test_fine_clip_vectors():
    """fine_clip_test.go:14-160, the PTCLs of TestFineRasterizeClip / NestedClip / DeepClip word for word (their EndClip blend
    word 0 written as 0x8003, see test_fine_tile_clip_vector): red 50 % in a clip over white -> (1, .498, .498, 1); five nested
    clips, each over a red solid, blue innermost -> blue."""
    empty = np.zeros(0, dtype=T.SEGMENT)
    one = int(np.float32(1.0).view(np.uint32))

    def rgba(r, g, b, a):
        return r | g << 8 | b << 16 | a << 24
    w = [0, 3, 5, rgba(255, 255, 255, 255), 10, 3, 5, rgba(128, 0, 0, 128), 11, 0x8003, one, 0]
    px = T.fine_tile(np.array(w, dtype=np.uint32), empty, (1, 1, 1, 1))[8 * 16 + 8]
    a = np.float32(128) / np.float32(255)
    assert np.allclose(px, [1.0, 1 - a, 1 - a, 1.0], atol=1e-6)
    w = [0] + [3, 5, rgba(255, 0, 0, 255), 10] * 5 + [3, 5, rgba(0, 0, 255, 255)] + [11, 0x8003, one] * 5 + [0]
    px = T.fine_tile(np.array(w, dtype=np.uint32), empty, (1, 1, 1, 1))[8 * 16 + 8]
    assert np.allclose(px, [0, 0, 1, 1], atol=1e-6)       # blend-stack depth 5 > BlendStackSplit 4 (ptcl.go:31): spilled levels


def test_ptcl_pipeline_matches_per_path_pipeline():
    """fine_ptcl_test.go:193-410 TestRasterizeScenePTCL{SinglePath,MultiPath,MatchesExisting} and clip_integration_test.go:284-326:
    the PTCL pipeline and the per-path pipeline (RasterizeScene) agree within 1/255 on the reference's scenes."""
    scenes_ = [
        [dict(lines=T.polygon_lines([(10, 10), (54, 32), (10, 54)]), color=(255, 0, 0, 255))],
        [dict(lines=_rect(0, 0, 32, 32), color=(255, 0, 0, 255)), dict(lines=_rect(8, 8, 24, 24), color=(0, 0, 255, 128))],
        [dict(lines=T.flatten_fill(T.circle_cubics(32, 32, 25)), color=(0, 200, 0, 255)), dict(lines=T.polygon_lines(STAR), color=(128, 0, 0, 200), even_odd=1)],
    ]
    for sc in scenes_:
        e, l = T.make_elements(sc)
        a = T.rasterize_scene(WHITE, e, l, 64, 64)
        b, _ = T.Coarse(e, l, 64, 64).fine(WHITE)
        assert np.abs(a.astype(int) - b.astype(int)).max() <= 1


def test_exact_area_vs_gg_cpu_multicontour_folder():
    """internal/raster/multicontour_golden_test.go:16-110: the stroke-expanded folder outline (two contours, 11 quads) filled
    NonZero. `multicontour-fill-20x20.png` is gg's CPU output (its test demands diff == 0), `stroke-expanded-fill-20x20.png`
    Skia's; gg documents 21 pixels differing from Skia by up to 25 on the quad corners and none on straight edges. Exact-area
    coverage through the same compositing formula (:291-319): straight edges identical to both, the curve-corner pixels within
    the same band."""
    from gg_b200 import _lib, scene as S
    verbs = [0, 1, 1, 1, 1, 1, 2, 1, 2, 1, 2, 1, 2, 1, 2, 1, 4,
             0, 1, 2, 1, 2, 2, 1, 2, 1, 2, 1, 2, 1, 1, 1, 4]
    pts = [10.84, 5.1921, 11.0486, 5.3659, 10.7285, 5.75, 10.7285, 5.25, 11, 5.25, 17, 5.25, 18.75, 5.25, 18.75, 7, 18.75, 15.1667,
           18.75, 17.25, 16.75, 17.25, 3.25, 17.25, 1.25, 17.25, 1.25, 15.1667, 1.25, 4.8333, 1.25, 2.75, 3.25, 2.75, 7.6379, 2.75,
           7.9095, 2.75, 8.1181, 2.9238, 10.84, 5.1921, 10.1998, 5.9603, 7.4779, 3.6921, 7.5475, 3.75, 7.6379, 3.75, 3.25, 3.75,
           2.8507, 3.75, 2.5579, 4.052, 2.25, 4.3696, 2.25, 4.8333, 2.25, 15.1667, 2.25, 16.25, 3.25, 16.25, 16.75, 16.25,
           17.75, 16.25, 17.75, 15.1667, 17.75, 7, 17.75, 6.25, 17, 6.25, 11, 6.25, 10.5475, 6.25, 10.1998, 5.9603]
    c = _lib.Context(-1)
    c.begin(20, 20)
    c.fill_path(verbs, pts, (255, 255, 255, 255), 0)
    words, lay = c.pack_host()
    img, _ = T.render_packed(words, lay, 20, 20)
    cov = img[..., 3].astype(np.uint16)
    bg, fg = np.array([0x3C, 0x3F, 0x41], np.uint16), np.array([0xCE, 0xD0, 0xD6], np.uint16)
    scale = cov + 1
    src = (fg[None, None, :] * scale[..., None]) >> 8
    srca = (0xFF * scale) >> 8
    out = np.where(cov[..., None] == 0, bg[None, None, :], (src + ((bg[None, None, :] * ((255 - srca) + 1)[..., None]) >> 8)) & 0xFF).astype(int)
    for name, max_allowed in (("multicontour-fill-20x20", 32), ("stroke-expanded-fill-20x20", 20)):   # measured: max 29 / 18, mean 1.21 / 0.98
        g = np.array(Image.open(os.path.join(HERE, "golden", "skia-aaa", name + ".png")).convert("RGBA"))[..., :3].astype(int)
        d = np.abs(out - g).max(-1)
        assert d.mean() <= 1.5 and (d <= 2).mean() >= 0.85 and d.max() <= max_allowed, (name, d.mean(), d.max(), (d <= 2).mean())


# ---- gradient brushes: the oracle's CmdGrad against gg's formulas restated independently in numpy ----
def _gg_color_at(stops, t, extend):
    """gradient.go:42-131 (applyExtendMode, colorAtOffset, interpolateColorLinear) + internal/color/convert.go:8-23."""
    f = np.float32
    if extend == 1:
        t = t - np.floor(t)
    elif extend == 2:
        t = abs(t); per = np.floor(t); t -= per
        if int(per) % 2 == 1:
            t = 1 - t
    else:
        t = min(1.0, max(0.0, t))
    stops = sorted(stops, key=lambda s: s[0])
    idx = next((i for i, s in enumerate(stops) if s[0] >= t), len(stops))
    if idx == 0:
        return list(stops[0][1:])
    if idx >= len(stops):
        return list(stops[-1][1:])
    a, b = stops[idx - 1], stops[idx]
    if a[0] == b[0]:
        return list(a[1:])
    lt = f((t - a[0]) / (b[0] - a[0]))
    s2l = lambda s: f(s) / f(12.92) if f(s) <= f(0.04045) else f(((float(f(s)) + 0.055) / 1.055) ** 2.4)   # noqa: E731
    l2s = lambda l: f(l) * f(12.92) if f(l) <= f(0.0031308) else f(1.055) * f(float(l) ** (1 / 2.4)) - f(0.055)   # noqa: E731
    rgb = [float(l2s(s2l(a[1 + k]) + lt * (s2l(b[1 + k]) - s2l(a[1 + k])))) for k in range(3)]
    return rgb + [float(f(a[4]) + lt * (f(b[4]) - f(a[4])))]


def _gg_focal_t(x, y, cx, cy, r1, fx0, fy0):
    """gradient_radial.go:131-196 computeTFocal."""
    dx, dy, fx, fy = x - fx0, y - fy0, cx - fx0, cy - fy0
    a, b, c = dx * dx + dy * dy, -2 * (dx * fx + dy * fy), fx * fx + fy * fy - r1 * r1
    if a == 0:
        return 0.0
    disc = b * b - 4 * a * c
    if disc < 0:
        return 1.0
    t1, t2 = (-b - np.sqrt(disc)) / (2 * a), (-b + np.sqrt(disc)) / (2 * a)
    pos = [v for v in (t1, t2) if v > 0]
    if not pos:
        return 0.0
    idist = min(pos) * np.sqrt(a)
    return 0.0 if idist == 0 else float(np.sqrt(a) / idist)


def _gg_sweep_t(x, y, cx, cy, a0, a1):
    """gradient_sweep.go:79-150."""
    sweep = a1 - a0
    if sweep == 0:
        return 0.0
    rel = float(np.arctan2(y - cy, x - cx)) - a0
    two_pi = 2 * np.pi
    if sweep > 0:
        while rel < 0:
            rel += two_pi
        while rel >= two_pi:
            rel -= two_pi
    else:
        while rel > 0:
            rel -= two_pi
        while rel <= -two_pi:
            rel += two_pi
    return rel / sweep


@pytest.mark.parametrize("kind,geom,extend", [(0, (10, 5, 80, 40), 0), (0, (30, 0, 50, 0), 1), (0, (30, 10, 50, 30), 2),
                                              (1, (48, 32, 4, 40), 0), (1, (20, 20, 0, 12), 2), (0, (5, 5, 5, 5), 0),
                                              (2, (48, 32, 0.0, 6.283185307179586), 0), (2, (40.5, 20.5, 1.0, -2.5), 1), (2, (10, 60, 0.5, 0.5), 0),
                                              (3, (48, 32, 0, 40, 60, 40), 0), (3, (30, 30, 5, 25, 20, 22), 2), (3, (30, 30, 7, 7, 20, 22), 0)])
def test_gradient_fill_matches_gg_formulas(kind, geom, extend):
    from gg_b200 import _lib
    w, h = 96, 64
    stops = [(1.0, 0.1, 0.2, 0.9, 1.0), (0.0, 1.0, 0.0, 0.0, 1.0), (0.4, 0.0, 1.0, 0.0, 0.5)]    # unsorted on purpose
    c = _lib.Context(-1)
    c.begin(w, h)
    c.fill_path_gradient([0, 1, 1, 1, 4], [0, 0, w, 0, w, h, 0, h], kind, geom, stops, extend)
    words, lay = c.pack_host()
    c.close()
    out, _ = T.render_packed(words, lay, w, h, (0, 0, 0, 0), 1)
    rng = np.random.default_rng(1)
    for _ in range(60):
        x, y = int(rng.integers(0, w)), int(rng.integers(0, h))
        fx, fy = x + 0.5, y + 0.5
        g = [float(np.float32(v)) for v in geom]
        if kind == 0:
            dx, dy = g[2] - g[0], g[3] - g[1]
            l2 = dx * dx + dy * dy
            col = sorted(stops)[0][1:] if l2 == 0 else _gg_color_at(stops, ((fx - g[0]) * dx + (fy - g[1]) * dy) / l2, extend)
        elif kind == 1:
            col = _gg_color_at(stops, (np.hypot(fx - g[0], fy - g[1]) - g[2]) / (g[3] - g[2]), extend)
        elif kind == 2:
            col = sorted(stops)[0][1:] if (fx == g[0] and fy == g[1]) else _gg_color_at(stops, _gg_sweep_t(fx, fy, *g[:4]), extend)
        else:
            col = sorted(stops)[0][1:] if g[3] - g[2] == 0 else _gg_color_at(stops, _gg_focal_t(fx, fy, g[0], g[1], g[3], g[4], g[5]), extend)
        want = [int(min(1.0, max(0.0, col[k] * col[3])) * 255 + 0.5) for k in range(3)] + [int(col[3] * 255 + 0.5)]
        assert np.abs(out[y, x].astype(int) - np.array(want)).max() <= 1, (x, y, out[y, x], want)


# ---- TagFillRoundRect: the oracle's SDF command against scene/shape.go:243-274 + scene/renderer.go:986-1070 ----
def _sdf_cov(px, py, cx, cy, hw, hh, r):
    f = np.float32
    px, py, cx, cy, hw, hh, r = (f(v) for v in (px, py, cx, cy, hw, hh, r))
    dx = abs(px - cx) - hw + r
    dy = abs(py - cy) - hh + r
    mx, my = max(dx, f(0)), max(dy, f(0))
    outside = f(np.sqrt(np.float64(mx * mx + my * my)))
    inside = min(max(dx, dy), f(0))
    d = outside + inside - r
    aa = f(0.7)
    if d >= aa:
        return f(0)
    if d <= -aa:
        return f(1)
    t = (d + aa) / (f(2) * aa)
    return f(1) - (t * t * (f(3) - f(2) * t))


def _rrect_encoding(rects, w, h):
    from gg_b200 import scene as S
    enc = S.Encoding()
    for (rect, rx, ry, col, tr) in rects:
        enc.EncodeTransform(tr)
        enc.EncodeFillRoundRect(col, rect, rx, ry)
    return enc


def test_round_rect_sdf_known_answers():
    """scene/roundrect_shape_test.go:165-219: the coverage table of the reference's own tests."""
    cx, cy, hw, hh, r = 50, 40, 50, 40, 10
    for (px, py, inside) in [(50, 40, True), (30, 30, True), (49.5, 0.5, True), (200, 200, False), (0.1, 0.1, False)]:
        assert (_sdf_cov(px, py, cx, cy, hw, hh, r) > 0.5) == inside
    from gg_b200 import _lib
    c = _lib.Context(-1)
    c.begin(128, 96)
    c.add_encoding(*_rrect_encoding([((0, 0, 100, 80), 10, 10, (1, 1, 1, 1), (1, 0, 0, 0, 1, 0))], 128, 96).streams())
    words, lay = c.pack_host()
    c.close()
    out, _ = T.render_packed(words, lay, 128, 96, (0, 0, 0, 0), 1)
    assert out[40, 50, 3] == 255 and out[30, 30, 3] == 255 and out[0, 49, 3] > 127      # pixel (49, 0) has its centre at (49.5, 0.5)
    assert out[90, 120, 3] == 0 and out[0, 0, 3] < 128
    # smoothstep on the edge: a pixel centre exactly on the boundary is half covered (TestSmoothstepCoverage32)
    assert abs(float(_sdf_cov(100, 40, cx, cy, hw, hh, r)) - 0.5) < 0.01


@pytest.mark.parametrize("tr", [(1, 0, 0, 0, 1, 0), (1.5, 0, 7.25, 0, 0.75, -3.5), (-1, 0, 120, 0, 1, 2)])
def test_round_rect_sdf_matches_reference_formula(tr):
    """Every pixel of a frame of overlapping translucent round rects (axis-aligned transforms, incl. a mirror) equals the
    reference's renderFillRoundRect + blendSDF restated in numpy, to 1/255 (the oracle composites in float32, the reference's
    pixmap holds float32 too)."""
    from gg_b200 import _lib
    w, h = 160, 112
    rects = [((10.3, 8.6, 90.2, 70.9), 12, 12, (0.9, 0.2, 0.1, 0.8), tr), ((40, 30, 70, 100), 30, 8, (0.1, 0.3, 0.9, 0.5), tr),
             ((60.5, 5.5, 64.5, 9.5), 9, 9, (0, 1, 0, 1), tr)]
    c = _lib.Context(-1)
    c.begin(w, h)
    c.add_encoding(*_rrect_encoding(rects, w, h).streams())
    words, lay = c.pack_host()
    c.close()
    tail = words[int(lay["n_scene_words"]):]
    assert tail[6] == 3                                       # three SDF records behind the packed scene
    out, _ = T.render_packed(words, lay, w, h, (0, 0, 0, 0), 1)
    want = np.zeros((h, w, 4), np.float32)
    f = np.float32
    for (rect, rx, ry, col, t) in rects:
        x0, y0 = f(t[0]) * f(rect[0]) + f(t[1]) * f(rect[1]) + f(t[2]), f(t[3]) * f(rect[0]) + f(t[4]) * f(rect[1]) + f(t[5])
        x1, y1 = f(t[0]) * f(rect[2]) + f(t[1]) * f(rect[3]) + f(t[2]), f(t[3]) * f(rect[2]) + f(t[4]) * f(rect[3]) + f(t[5])
        x0, x1 = min(x0, x1), max(x0, x1)
        y0, y1 = min(y0, y1), max(y0, y1)
        cx, cy, hw, hh = (x0 + x1) / f(2), (y0 + y1) / f(2), (x1 - x0) / f(2), (y1 - y0) / f(2)
        r = min(min(f(rx), f(ry)), min(hw, hh))
        # the product quantises the brush to 8 bits at ingest like every other draw (path_convert.go:131-140)
        q = [f(int(min(255.0, v * 255.0 + 0.5))) / f(255) for v in col]
        for py in range(h):
            for px in range(w):
                cov = _sdf_cov(px + 0.5, py + 0.5, cx, cy, hw, hh, r)
                if cov <= 0:
                    continue
                sa = q[3] * cov
                inv = f(1) - sa
                pm = np.array([f(int(q[0] * q[3] * f(255) + f(0.5))) / f(255), f(int(q[1] * q[3] * f(255) + f(0.5))) / f(255),
                               f(int(q[2] * q[3] * f(255) + f(0.5))) / f(255)], np.float32)
                want[py, px, :3] = want[py, px, :3] * inv + pm * cov
                want[py, px, 3] = want[py, px, 3] * inv + sa
    wq = np.floor(np.clip(want, 0, 1) * 255 + 0.5).astype(int)
    d = np.abs(out.astype(int) - wq)
    assert d.max() <= 1, (d.max(), np.argwhere(d > 1)[:5])
    assert out[..., 3].max() > 100


def test_round_rect_rotated_falls_back_to_the_outline():
    """A transform that is not axis-aligned has no SDF form in the reference either (renderFillRoundRect transforms two
    corners only); ggcuda fills the transformed outline with the exact-area path instead: no SDF record."""
    from gg_b200 import _lib
    c = _lib.Context(-1)
    c.begin(128, 128)
    c.add_encoding(*_rrect_encoding([((20, 20, 90, 70), 10, 10, (1, 0, 0, 1), (0.8, -0.6, 30, 0.6, 0.8, 0))], 128, 128).streams())
    words, lay = c.pack_host()
    c.close()
    assert words[int(lay["n_scene_words"]):][6] == 0
    out, _ = T.render_packed(words, lay, 128, 128, (0, 0, 0, 0), 1)
    assert out[..., 3].max() == 255


# ---- TagImage: the oracle's image command against scene/renderer.go:1093-1243 restated in numpy, bytes and all ----
def _ref_blit_image(pm, img, t):
    """blitImageToTile + blitBilinearPixel + blitNearestPixel on a whole-canvas premultiplied RGBA8 pixmap (tile = canvas)."""
    f = np.float32
    H, W = pm.shape[:2]
    ih, iw = img.shape[:2]
    A, B, C_, D, E, F = (f(v) for v in t)
    det = A * E - B * D
    if det == 0:
        return
    inv_det = f(1.0) / det
    iA, iB, iC = E * inv_det, -B * inv_det, (B * F - E * C_) * inv_det
    iD, iE, iF = -D * inv_det, A * inv_det, (D * C_ - A * F) * inv_det
    xs, ys = [f(0)], [f(0)]                                  # the box is seeded with the untransformed (0, 0)
    for cx, cy in ((0, 0), (iw, 0), (iw, ih), (0, ih)):
        xs.append(A * f(cx) + B * f(cy) + C_)
        ys.append(D * f(cx) + E * f(cy) + F)
    sx0, sy0 = max(int(min(xs)), 0), max(int(min(ys)), 0)
    ex, ey = min(int(max(xs)) + 1, W), min(int(max(ys)) + 1, H)
    cb = lambda v: 0 if v <= 0 else (255 if v >= 255 else int(f(v) + f(0.5)))   # noqa: E731  clampByte
    for py in range(sy0, ey):
        cyp = f(py) + f(0.5)
        for px in range(sx0, ex):
            cxp = f(px) + f(0.5)
            sx = iA * cxp + iB * cyp + iC - f(0.5)
            sy = iD * cxp + iE * cyp + iF - f(0.5)
            flx, fly = np.floor(sx), np.floor(sy)
            ix0, iy0 = int(flx), int(fly)
            if ix0 + 1 < 0 or iy0 + 1 < 0 or ix0 >= iw or iy0 >= ih:
                continue
            wx, wy = sx - flx, sy - fly
            d = pm[py, px]
            if wx == 0 and wy == 0:
                if 0 <= ix0 < iw and 0 <= iy0 < ih:
                    s = img[iy0, ix0]
                    sa = int(s[3])
                    if sa == 0:
                        continue
                    if sa == 255:
                        d[:3] = s[:3]; d[3] = 255
                    else:
                        inv = 255 - sa
                        for k in range(4):
                            d[k] = (int(s[k]) + (int(d[k]) * inv + 127) // 255) & 0xFF
                continue
            cx0, cx1 = min(max(ix0, 0), iw - 1), min(max(ix0 + 1, 0), iw - 1)
            cy0, cy1 = min(max(iy0, 0), ih - 1), min(max(iy0 + 1, 0), ih - 1)
            tex = [img[cy0, cx0], img[cy0, cx1], img[cy1, cx0], img[cy1, cx1]]
            tex = [np.zeros(4, f) if p[3] == 0 else p.astype(f) for p in tex]
            ifx, ify = f(1) - wx, f(1) - wy
            wts = [ifx * ify, wx * ify, ifx * wy, wx * wy]
            s = [tex[0][k] * wts[0] + tex[1][k] * wts[1] + tex[2][k] * wts[2] + tex[3][k] * wts[3] for k in range(4)]
            if s[3] < f(0.5 / 255.0):
                continue
            inv = f(1) - s[3] / f(255)
            for k in range(4):
                d[k] = cb(s[k] + f(d[k]) * inv)


def _test_image(h, w, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, (h, w, 1)).astype(np.float32)
    a[rng.random((h, w, 1)) < 0.15] = 0          # transparent holes (fetchPremul's special case)
    a[rng.random((h, w, 1)) < 0.3] = 255
    rgb = rng.integers(0, 256, (h, w, 3)).astype(np.float32)
    return np.concatenate([np.floor(rgb * a / 255.0), a], axis=2).astype(np.uint8)   # premultiplied: colour <= alpha


@pytest.mark.parametrize("t", [(1, 0, 10, 0, 1, 7), (1, 0, 10.5, 0, 1, 7.25), (2.5, 0, 3, 0, 1.75, 2), (0.8, -0.6, 30, 0.6, 0.8, 5),
                               (-1, 0, 60, 0, 1, 4), (0.3, 0, 20, 0, 0.4, 20), (1, 0, -6, 0, 1, -5)])
def test_image_matches_reference_blit(t):
    """One image over an opaque and a translucent fill: every pixel within 1/255 of the reference's byte arithmetic (the
    nearest-texel path on integer translations, bilinear with clamp-to-edge otherwise, the half-texel border, holes)."""
    from gg_b200 import _lib, scene as S
    W, H = 96, 64
    img = _test_image(14, 20, 5)
    enc = S.Encoding()
    enc.EncodeTransform(S.IDENTITY)
    enc.EncodePath(*S.rect_verbs_coords(0, 0, 50, 64)); enc.EncodeFill((0.2, 0.4, 0.6, 1.0))
    enc.EncodePath(*S.rect_verbs_coords(40, 10, 96, 50)); enc.EncodeFill((1.0, 0.5, 0.0, 0.5))
    ix = enc.AddImage(img)
    enc.EncodeImage(ix, t)
    enc.EncodeTransform((1, 0, 0, 0, 1, 0))     # the image's affine must not have become the current transform ...
    enc.EncodePath(*S.rect_verbs_coords(90, 60, 96, 64)); enc.EncodeFill((0.0, 1.0, 0.0, 1.0))
    c = _lib.Context(-1)
    c.begin(W, H)
    c.add_image(img)
    c.add_encoding(*enc.streams())
    words, lay = c.pack_host()
    c.close()
    out, _ = T.render_packed(words, lay, W, H, (0, 0, 0, 0), 1)
    pm = np.zeros((H, W, 4), np.uint8)
    pm[:, :50] = (51, 102, 153, 255)
    q = lambda v: int(min(255.0, v * 255.0 + 0.5))   # noqa: E731
    src = np.array([q(1.0) * q(0.5) / 255.0, q(0.5) * q(0.5) / 255.0, 0, q(0.5)], np.float32)   # premultiplied translucent orange
    reg = pm[10:50, 40:96].astype(np.float32)
    pm[10:50, 40:96] = np.floor(src + reg * (1 - src[3] / 255.0) + 0.5).astype(np.uint8)
    base = pm.copy()
    _ref_blit_image(pm, img, t)
    pm[60:64, 90:96] = (0, 255, 0, 255)
    base[60:64, 90:96] = (0, 255, 0, 255)
    d = np.abs(out.astype(int) - pm.astype(int))
    assert (pm != base).any(), "the reference drew nothing: bad test transform"
    assert d.max() <= 1, (d.max(), np.argwhere(d.max(axis=2) > 1)[:5], out[tuple(np.argwhere(d.max(axis=2) > 1)[0])], pm[tuple(np.argwhere(d.max(axis=2) > 1)[0])])
    # exactly the pixels the reference touches are touched
    touched_ref = (pm != base).any(axis=2)
    out0, _ = T.render_packed(*_pack_without_image(enc, W, H), W, H, (0, 0, 0, 0), 1)
    touched = (out != out0).any(axis=2)
    assert (touched & ~touched_ref).sum() <= touched_ref.sum() * 0.02 + 2 and (touched_ref & ~touched).sum() <= touched_ref.sum() * 0.02 + 2


def _pack_without_image(enc, W, H):
    from gg_b200 import _lib
    c = _lib.Context(-1)
    c.begin(W, H)
    c.add_encoding(*enc.streams())     # no image registered: TagImage draws nothing (renderer.go:773)
    words, lay = c.pack_host()
    c.close()
    return words, lay


def test_image_stream_consumption_and_degenerates():
    """TagImage consumes one drawData word and one transform; a singular affine, an unknown index draw nothing."""
    from gg_b200 import _lib, scene as S
    img = _test_image(4, 4, 1)
    enc = S.Encoding()
    enc.AddImage(img)
    enc.EncodeImage(0, (0, 0, 5, 0, 0, 5))          # det == 0
    enc.EncodeImage(7, (1, 0, 5, 0, 1, 5))          # no such image
    enc.EncodeTransform((1, 0, 2, 0, 1, 3))
    enc.EncodePath(*S.rect_verbs_coords(0, 0, 4, 4)); enc.EncodeFill((1, 1, 1, 1))
    c = _lib.Context(-1)
    c.begin(32, 32)
    c.add_image(img)
    c.add_encoding(*enc.streams())
    words, lay = c.pack_host()
    c.close()
    out, _ = T.render_packed(words, lay, 32, 32, (0, 0, 0, 0), 1)
    want = np.zeros((32, 32, 4), np.uint8)
    want[3:7, 2:6] = 255                             # the rectangle under ITS transform, nothing else
    assert (out == want).all()
