"""CPU tests (no GPU): the C-ABI library loads and exports every declared symbol, host-side scene packing
(host-only context, no compute), and the multi-rank band assembly over gloo."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from gg_b200 import _lib, scene as S, scenes
from oracle import twin as T

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "ggcuda.h")).read()
    declared = set(re.findall(r"GGCUDA_API\s+[\w\s\*]+?\b(ggcuda_\w+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    L = ctypes.CDLL(_lib.LIB_PATH)
    for sym in declared:
        assert hasattr(L, sym), sym
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l and "ggcuda_" in l}
    assert exported == declared, exported ^ declared   # nothing else leaks out of the shared object


def test_no_cpu_fallback():
    """A host-only context packs scenes but refuses to render; there is no software path in the product."""
    c = _lib.Context(-1)
    c.begin(64, 64)
    c.fill_path([0, 1, 1, 4], [1, 1, 30, 1, 15, 30], (255, 0, 0, 255))
    with pytest.raises(_lib.GGCudaError) as e:
        c.flush(np.zeros((64, 64, 4), np.uint8))
    assert e.value.code == _lib.ERR_UNSUPPORTED
    src = "".join(open(os.path.join(ROOT, "gg_b200", f)).read() for f in os.listdir(os.path.join(ROOT, "gg_b200")) if f.endswith(".py"))
    assert "oracle" not in src.replace("oracle/", "")   # the package never imports the test oracle


def _pack(build):
    c = _lib.Context(-1)
    c.begin(256, 128)
    build(c)
    words, lay = c.pack_host()
    c.close()
    return words, lay


def _tags(words, lay):
    return words[lay["path_tag_base"]:lay["path_tag_base"] + lay["n_tag_words"]].view(np.uint8)[:lay["n_tag_bytes"]]


def test_packed_layout_and_autoclose():
    """PackedScene layout (scene_encode.go:280-356) and the ingest rules: one Transform when it changes, one Style and
    one Path marker per draw, MoveTo tag 0x0C, open subpaths closed, zero-length closes dropped."""
    def build(c):
        c.fill_path([0, 1, 1], [10, 10, 50, 10, 30, 40], (255, 0, 0, 255), 0)           # open triangle -> closing line added
        c.fill_path([0, 1, 1, 1, 4], [60, 10, 90, 10, 90, 40, 60, 10], (0, 255, 0, 128), 1)   # explicit return to start + Close
    words, lay = _pack(build)
    assert lay["n_tag_words"] % 256 == 0 and lay["path_tag_base"] == 0
    assert lay["path_data_base"] == lay["n_tag_words"]
    assert (lay["n_draws"], lay["n_paths"], lay["n_clips"]) == (2, 2, 0)
    assert list(_tags(words, lay)) == [0x20, 0x40, 0x0C, 0x09, 0x09, 0x09, 0x10, 0x40, 0x0C, 0x09, 0x09, 0x09, 0x10]
    pd = words[lay["path_data_base"]:lay["draw_tag_base"]].view(np.float32)
    assert list(pd[:8]) == [10, 10, 50, 10, 30, 40, 10, 10]
    assert list(words[lay["draw_tag_base"]:lay["draw_data_base"]]) == [0x44, 0x44]
    a = np.float32(128) / np.float32(255)
    assert int(words[lay["draw_data_base"]]) == 0xFF0000FF
    assert int(words[lay["draw_data_base"] + 1]) == (int(np.float32(255) * a + np.float32(0.5)) << 8) | (128 << 24)
    assert list(words[lay["style_base"]:lay["style_base"] + 6]) == [0, 0, 0, 2, 0, 0]     # 3 words per path; even-odd = bit 1
    assert list(words[lay["transform_base"]:lay["style_base"]].view(np.float32)) == [1, 0, 0, 0, 1, 0]
    # monoids of the packed tags agree with the oracle's restatement of pathtag.go
    m = T.path_monoid(int(words[0]))
    assert (int(m["trans_ix"]), int(m["style_ix"]), int(m["path_seg_ix"]), int(m["path_seg_offset"])) == (1, 1, 1, 4)


def test_clip_and_layer_pairing():
    def build(c):
        c.push_layer(S.BlendMultiply, 0.5)
        c.push_clip([0, 1, 1, 4], [0, 0, 100, 0, 50, 100])
        c.fill_path([0, 1, 1, 4], [10, 10, 50, 10, 30, 40], (255, 0, 0, 255))
        c.pop()
        c.pop()
        c.push_layer(S.BlendClear, 1.0)
        # left open on purpose: packing closes it (scene/renderer.go:789-797)
    words, lay = _pack(build)
    assert (lay["n_draws"], lay["n_clips"]) == (7, 6)
    dt = list(words[lay["draw_tag_base"]:lay["draw_data_base"]])
    assert dt == [0x9, 0x9, 0x44, 0x21, 0x21, 0x9, 0x21]
    aux = words[lay["clip_aux_base"]:lay["clip_aux_base"] + 14].view(np.int32).reshape(7, 2)
    assert list(aux[:, 0]) == [-1, 0, 1, 1, 0, -1, 5]          # enclosing clip (EndClip: its own BeginClip)
    assert aux[0, 1] == 4 and aux[1, 1] == 3 and aux[5, 1] == 6  # Begin -> End links
    dd = words[lay["draw_data_base"]:lay["transform_base"]]
    assert dd[0] == ((S.BlendMultiply << 8) | 3) | 0xC0000000   # mix mode, SrcOver compose; implicit + elidable layer
    assert dd[1] == np.float32(0.5).view(np.uint32)
    assert dd[2] == 0x8003                                       # plain clip
    assert dd[5] == 0                                            # Clear compose: must cover the canvas, not elidable
    # the Multiply layer is implicit: its path is empty (the first path data belongs to the clip triangle);
    # the Clear layer keeps an explicit full-canvas rectangle
    pd = words[lay["path_data_base"]:lay["draw_tag_base"]].view(np.float32)
    assert list(pd[:6]) == [0, 0, 100, 0, 50, 100]
    assert list(pd[-10:]) == [0, 0, 256, 0, 256, 128, 0, 128, 0, 0]


def test_encoding_ingest_matches_per_draw_calls():
    """scene.Encoding streams (tags 0x01..0x41) produce the same packed scene as the equivalent per-draw calls."""
    sc = S.Scene()
    shape = S.circle_verbs_coords(40, 40, 20)
    sc.Fill(S.FillEvenOdd, S.IDENTITY, (1.0, 0.5, 0.0, 0.5), shape)
    sc.PushClip(S.rect_verbs_coords(0, 0, 50, 50))
    sc.Fill(S.FillNonZero, S.IDENTITY, (0.0, 0.0, 1.0, 1.0), shape)
    sc.PopClip()
    c = _lib.Context(-1)
    c.begin(256, 128)
    c.add_encoding(*sc.Encoding().streams())
    w1, l1 = c.pack_host()
    c.begin(256, 128)
    col = lambda rgba: tuple(int(min(255.0, v * 255.0 + 0.5)) for v in rgba)
    c.fill_path(*shape, col((1.0, 0.5, 0.0, 0.5)), 1)
    c.push_clip(*S.rect_verbs_coords(0, 0, 50, 50))
    c.fill_path(*shape, col((0.0, 0.0, 1.0, 1.0)), 0)
    c.pop()
    w2, l2 = c.pack_host()
    assert l1 == l2 and (w1 == w2).all()


def test_pooled_ingest_is_byte_identical():
    """GGCUDA_INGEST_THREADS=n (opt-in thread pool, host_scene.cpp IngestPool) packs the same bytes as the one-thread walk,
    whole frame and with a band's culling active."""
    import subprocess
    import sys
    tool = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "ingest_bench.py")
    for bands in ("1", "4"):
        md5 = set()
        for n in ("1", "4"):
            env = dict(os.environ, GGCUDA_INGEST_THREADS=n, GG_INGEST_REPS="1")
            out = subprocess.run([sys.executable, tool, bands], env=env, capture_output=True, text=True, timeout=300)
            assert out.returncode == 0, out.stderr[-2000:]
            md5.add(out.stdout.split("md5")[1].split()[0])
        assert len(md5) == 1, md5


def test_unsupported_tags_are_reported():
    c = _lib.Context(-1)
    c.begin(64, 64)
    with pytest.raises(_lib.GGCudaError) as e:
        c.add_encoding(np.array([S.TagText], np.uint8), [], [0], [1, 0, 0, 0, 1, 0], [])
    assert e.value.code == _lib.ERR_UNSUPPORTED     # the Go binding maps this to gg.ErrFallbackToCPU
    c.begin(64, 64)
    c.add_encoding(np.array([S.TagImage], np.uint8), [], [0], [1, 0, 0, 0, 1, 0], [])   # an image that was never registered: nothing drawn
    assert c.pack_host()[1]["n_draws"] == 0
    with pytest.raises(_lib.GGCudaError) as e:
        c.add_encoding(np.array([S.TagFill], np.uint8), [], [], [], [])
    assert e.value.code == _lib.ERR_INVALID         # stream underrun


def test_stroke_outline_area():
    """Stroke of a straight line (expanded by the oracle's statement of the device stroker): length x width (+ caps) within 2 %."""
    for cap, extra in ((0, 0.0), (2, 1.0), (1, np.pi / 4)):
        c = _lib.Context(-1)
        c.begin(128, 64)
        c.stroke_path([0, 1], [20, 32, 100, 32], (255, 255, 255, 255), 10.0, cap, 0, 4.0)
        words, lay = c.pack_host()
        img, _ = T.render_packed(words, lay, 128, 64)
        area = img[..., 3].astype(np.float64).sum() / 255
        want = 80 * 10 + extra * 10 * 10
        assert abs(area - want) / want < 0.02, (cap, area, want)


def test_config_generators_are_deterministic():
    a = scenes.config1()[0].streams()
    b = scenes.config1()[0].streams()
    assert all((x == y).all() for x, y in zip(a, b))
    enc, w, h = scenes.config3(n=200, w=640, h=480)
    tags = enc.streams()[0]
    assert (tags == S.TagPushLayer).sum() == (tags == S.TagPopLayer).sum() >= 4
    assert (tags == S.TagStroke).sum() > 0 and (tags == S.TagBeginClip).sum() == (tags == S.TagEndClip).sum()


def test_bands_assemble_over_gloo(tmp_path):
    """world_size-2 gloo run of the band decomposition: each rank produces its band (here with the oracle, since
    there is no GPU and no fallback) and one all_gather_into_tensor assembles the frame on every rank."""
    script = tmp_path / "rank.py"
    script.write_text(f'''
import os, sys
sys.path.insert(0, {ROOT!r})
import numpy as np, torch, torch.distributed as dist
from gg_b200 import _lib, bands, scenes
from oracle import twin as T
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
enc, w, h = scenes.config3(n=300, w=320, h=200)
c = _lib.Context(-1); c.begin(w, h); c.add_encoding(*enc.streams()); words, lay = c.pack_host()
full, _ = T.render_packed(words, lay, w, h)
y0, y1 = bands.band_rows(h, world, rank)
frame = bands.alloc_frame(w, h, world, "cpu")
band = bands.band_view(frame, h, world, rank)
rows = full[y0 * 16:min(y1 * 16, h)]
band[:rows.shape[0]] = torch.from_numpy(rows.copy())
bands.assemble(frame, h, world, rank)
assert (frame[:h].numpy() == full).all(), "assembled frame differs"
assert frame.shape[0] == bands.padded_height(h, world) and frame.shape[0] % (16 * world) == 0
dist.destroy_process_group()
print("rank", rank, "ok")
''')
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


def _stroke_alpha(v, c, width, cap, join, w, h, flags=0, miter=4.0):
    ctx = _lib.Context(-1, flags)
    ctx.begin(w, h)
    ctx.stroke_path(v, c, (255, 255, 255, 255), width, cap, join, miter)
    words, lay = ctx.pack_host()
    img, _ = T.render_packed(words, lay, w, h)
    return img[..., 3].astype(np.float64) / 255


def test_stroke_centre_line_encoding():
    """Stroked paths travel as centre lines: 3-word style, degenerate segments dropped, marker copy of the first segment
    (after a marker MoveTo when the subpath is open), dots only for round / square caps."""
    def build(c):
        c.stroke_path([0, 1, 1, 3], [10, 10, 10, 10, 50, 10, 50, 10, 50, 10, 50, 10], (255, 255, 255, 255), 4.0, 1, 2, 7.0)   # open; a repeated point and a null cubic
        c.stroke_path([0, 1, 1, 4], [10, 30, 50, 30, 30, 60], (255, 255, 255, 255), 2.0, 0, 0, 4.0)                            # closed: closing line added
        c.stroke_path([0, 0, 1], [5, 5, 80, 80, 90, 80], (255, 255, 255, 255), 2.0, 0, 0, 4.0)                                 # lone MoveTo + butt cap: nothing
        c.stroke_path([0], [70, 20], (255, 255, 255, 255), 6.0, 2, 0, 4.0)                                                     # dot with a square cap
    words, lay = _pack(build)
    assert list(_tags(words, lay)) == [0x20, 0x40, 0x0C, 0x09, 0x8C, 0x89, 0x10,
                                       0x40, 0x0C, 0x09, 0x09, 0x09, 0x89, 0x10,
                                       0x40, 0x0C, 0x09, 0x8C, 0x89, 0x10,
                                       0x40, 0x0C, 0x09, 0x8C, 0x89, 0x10]
    st = words[lay["style_base"]:lay["style_base"] + 12]
    assert int(st[0]) == 1 | (2 << 2) | (1 << 4) and st[1:3].view(np.float32).tolist() == [4.0, 7.0]
    assert int(st[3]) == 1 and int(st[9]) == 1 | (2 << 4)
    pd = words[lay["path_data_base"]:lay["draw_tag_base"]].view(np.float32)
    assert list(pd[:8]) == [10, 10, 50, 10, 10, 10, 50, 10]      # MoveTo, LineTo, marker MoveTo, marker copy of the LineTo


def test_stroke_areas_closed_forms():
    """Oracle statement of the stroke expander against closed forms: circles (also with the offset beyond the radius: no
    hole), joins, caps."""
    k = 0.5522847498307936
    for r, wd in ((40, 10), (40, 1), (3, 12), (1, 20), (10, 20)):
        kk = r * k
        c = [64 + r, 64, 64 + r, 64 + kk, 64 + kk, 64 + r, 64, 64 + r, 64 - kk, 64 + r, 64 - r, 64 + kk, 64 - r, 64,
             64 - r, 64 - kk, 64 - kk, 64 - r, 64, 64 - r, 64 + kk, 64 - r, 64 + r, 64 - kk, 64 + r, 64]
        a = _stroke_alpha([0, 3, 3, 3, 3, 4], c, wd, 0, 1, 128, 128)
        want = np.pi * ((r + wd / 2) ** 2 - max(r - wd / 2, 0) ** 2)
        assert abs(a.sum() - want) / want < 0.012, (r, wd, a.sum(), want)
        if r < wd / 2:
            assert a[64, 64] == 1.0          # the centre is inside the stroke
    # right-angle corner, width 10: two 50 x 10 arms overlapping in a 5 x 5 square + the outer join
    v, c = [0, 1, 1], [20, 20, 70, 20, 70, 70]
    base = 2 * 50 * 10 - 25
    for join, extra in ((2, 12.5), (0, 25.0), (1, np.pi * 25 / 4)):
        a = _stroke_alpha(v, c, 10, 0, join, 100, 100)
        assert abs(a.sum() - (base + extra)) < 1.0, (join, a.sum(), base + extra)
    a = _stroke_alpha(v, c, 10, 0, 0, 100, 100, miter=1.0)      # miter limit exceeded -> bevel
    assert abs(a.sum() - (base + 12.5)) < 1.0


def test_stroke_against_distance_field():
    """Round joins + round caps: the stroke is the set of points within half a width of the curve. Random cubics with cusps
    and loops, widths up to 24 px: no pixel clearly inside is uncovered, none clearly outside is covered."""
    rng = np.random.default_rng(0)
    w = h = 96
    ys, xs = np.mgrid[0:h, 0:w]
    P = np.stack([xs + 0.5, ys + 0.5], -1).reshape(-1, 1, 2)
    for it in range(12):
        n = int(rng.integers(1, 4))
        closed = bool(rng.random() < 0.5)
        c = rng.uniform(8, 88, 2 + 6 * n)
        wd = float(rng.uniform(0.5, 24))
        pts, cur = [], c[0:2]
        t = np.linspace(0, 1, 200)[:, None]
        for s in range(n):
            p1, p2, p3 = c[2 + 6 * s:4 + 6 * s], c[4 + 6 * s:6 + 6 * s], c[6 + 6 * s:8 + 6 * s]
            pts.append((1 - t) ** 3 * cur + 3 * (1 - t) ** 2 * t * p1 + 3 * (1 - t) * t * t * p2 + t ** 3 * p3)
            cur = p3
        pts = np.concatenate(pts + ([pts[0][:1]] if closed else []))
        a0, ab = pts[:-1][None], (pts[1:] - pts[:-1])[None]
        l2 = np.maximum((ab ** 2).sum(-1), 1e-30)
        tt = np.clip(((P - a0) * ab).sum(-1) / l2, 0, 1)
        dist = np.sqrt((((a0 + ab * tt[..., None]) - P) ** 2).sum(-1)).min(1).reshape(h, w)
        a = _stroke_alpha([0] + [3] * n + ([4] if closed else []), c, wd, 1, 1, w, h)
        assert not ((dist < wd / 2 - 0.9) & (a < 0.98)).any(), it
        assert not ((dist > wd / 2 + 0.9) & (a > 0.02)).any(), it


def test_wiping_layer_stops_at_enclosing_clip():
    """A layer whose blend mode wipes what it covers (Copy) is a full-canvas rectangle -- inside a clip, only as far as the
    clip's bounds: same pixels (the clip hides everything beyond), far fewer tiles."""
    from gg_b200 import scene as S
    w, h = 128, 64
    sc = S.Scene()
    sc.Fill(S.FillNonZero, S.IDENTITY, (0, 1, 0, 1), S.rect_verbs_coords(0, 0, w, h))
    sc.PushClip(S.rect_verbs_coords(0, 0, 64, h))
    sc.Fill(S.FillNonZero, S.IDENTITY, (1, 0, 0, 1), S.rect_verbs_coords(0, 0, w, h))    # red, inside the clip group
    sc.PushLayer(S.BlendCopy, 1.0, None)
    sc.Fill(S.FillNonZero, S.IDENTITY, (0, 0, 1, 1), S.rect_verbs_coords(16, 16, 48, 48))
    sc.PopLayer()
    sc.PopClip()
    c = _lib.Context(-1)
    c.begin(w, h)
    c.add_encoding(*sc.Encoding().streams())
    words, lay = c.pack_host()
    pd = words[lay["path_data_base"]:lay["draw_tag_base"]].view(np.float32)
    assert list(pd[30:38]) == [0, 0, 64, 0, 64, 64, 0, 64]          # the layer's rectangle: the clip's bounds, not the canvas
    img, _ = T.render_packed(words, lay, w, h)
    assert (img[:, 64:] == (0, 255, 0, 255)).all()                  # beyond the clip: the green below, untouched
    assert (img[16:48, 16:48] == (0, 0, 255, 255)).all()            # the layer's content
    assert (img[:16, :64] == (0, 255, 0, 255)).all()                # Copy wiped the clip group's red: green shows through


def test_svg_path_parser():
    from gg_b200.svgpath import parse_path
    v, c = parse_path("M3 5L7 8L3 11")
    assert v == [S.MOVE, S.LINE, S.LINE] and c == [3, 5, 7, 8, 3, 11]
    v, c = parse_path("m1 1h2v3l-1-1c1 0 2 1 2 2s0 1-1 1q1 1 2 0t2 0z")
    assert v == [S.MOVE, S.LINE, S.LINE, S.LINE, S.CUBIC, S.CUBIC, S.QUAD, S.QUAD, S.CLOSE]
    assert c[:8] == [1, 1, 3, 1, 3, 4, 2, 3]
    assert c[8:14] == [3, 3, 4, 4, 4, 5] and c[14:20] == [4, 6, 4, 6, 3, 6]      # S reflects the previous second control point
    assert c[20:24] == [4, 7, 5, 6] and c[24:28] == [6, 5, 7, 6]                 # T reflects the previous control point
    v, c = parse_path("M10.5199 5.57617L10.7285 5.75H11H17C17.6904 5.75 18.25 6.30964 18.25 7V15.1667Z")
    assert v == [S.MOVE, S.LINE, S.LINE, S.LINE, S.CUBIC, S.LINE, S.CLOSE] and c[4:8] == [11, 5.75, 17, 5.75]
    with pytest.raises(NotImplementedError):
        parse_path("M0 0A5 5 0 0 1 10 10")


def test_stroke_against_gg_folder_golden():
    """svg/golden_test.go:21-47: gg's CPU renderer reproduces Skia on this stroke-only icon (diff == 0 asserted there), so
    the golden IS gg's CPU output. Exact-area coverage of our stroke outline against it: straight edges identical, mean
    |d| ~ 1.3/255; the few pixels further apart sit on the rounded corners (AAA quantises coverage, gg hints strokes)."""
    from PIL import Image
    sc = S.Scene()
    scenes.add_icon(sc, scenes.ICON_FOLDER, 0.0, 0.0, 1.0)
    c = _lib.Context(-1)
    c.begin(20, 20)
    c.add_encoding(*sc.Encoding().streams())
    words, lay = c.pack_host()
    img, _ = T.render_packed(words, lay, 20, 20)
    g = np.array(Image.open(os.path.join(os.path.dirname(__file__), "golden", "svg", "folder_stroke_20x20.png")).convert("RGBA"))
    d = np.abs(img.astype(int) - g.astype(int))
    assert d.mean() <= 1.5 and (d.max(-1) > 16).mean() <= 0.05 and (d.max(-1) <= 2).mean() >= 0.85, (d.mean(), d.max())


def test_glyph_fixture_and_text_scene():
    """The committed glyph outlines (Go Regular, printable ASCII) and the configs[3] generator built on them."""
    import json
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "fixtures", "goregular_ascii.json")))
    assert g["units_per_em"] == 2048 and len(g["glyphs"]) == 94
    assert len(g["glyphs"]["O"]["contours"]) == 2 and len(g["glyphs"]["i"]["contours"]) == 2 and len(g["glyphs"]["B"]["contours"]) == 3
    enc, w, h = scenes.config4(g, n=400, w=640, h=480)
    c = _lib.Context(-1)
    c.begin(w, h)
    c.add_encoding(*enc.streams())
    words, lay = c.pack_host()
    assert lay["n_draws"] == 400
    img, _ = T.render_packed(words, lay, w, h, (255, 255, 255, 255))
    ink = (img[..., :3].sum(-1) < 3 * 255).mean()
    assert 0.05 < ink < 0.6        # text, not blobs: a plausible share of inked pixels


def test_push_layer_with_clip_is_one_clip():
    """scene.PushLayer(blend, alpha, clip) arrives as PushLayer, clip path, BeginClip ... EndClip, PopLayer (scene/scene.go:307-375):
    one BeginClip/EndClip pair whose path is the clip shape and whose blend word / alpha are the layer's; implicit layers
    (no clip, non-wiping mode) are numbered in the width word of their empty path's style."""
    sc = S.Scene()
    sc.PushLayer(S.BlendMultiply, 0.5, S.rect_verbs_coords(10, 10, 50, 40))      # mix 1, SrcOver
    sc.Fill(S.FillNonZero, S.IDENTITY, (1, 0, 0, 1), S.rect_verbs_coords(0, 0, 30, 30))
    sc.PopLayer()
    sc.PushLayer(S.BlendCopy, 1.0, S.rect_verbs_coords(5, 5, 20, 20))            # wiping mode: never elided
    sc.PopLayer()
    sc.PushLayer(S.BlendScreen, 1.0, None)                                       # implicit layer 0
    sc.PushLayer(S.BlendNormal, 0.7, None)                                       # implicit layer 1
    sc.Fill(S.FillNonZero, S.IDENTITY, (0, 1, 0, 1), S.rect_verbs_coords(0, 0, 8, 8))
    sc.PopLayer()
    sc.PopLayer()
    c = _lib.Context(-1)
    c.begin(64, 64)
    c.add_encoding(*sc.Encoding().streams())
    words, lay = c.pack_host()
    dt = list(words[lay["draw_tag_base"]:lay["draw_data_base"]])
    assert dt == [0x9, 0x44, 0x21, 0x9, 0x21, 0x9, 0x9, 0x44, 0x21, 0x21]
    dd = words[lay["draw_data_base"]:lay["transform_base"]]
    assert int(dd[0]) == (1 << 8 | 3) | 0x80000000 and dd[1:2].view(np.float32)[0] == 0.5      # Multiply, droppable where empty
    assert int(dd[3]) == 1 and dd[4:5].view(np.float32)[0] == 1.0                             # Copy: compose 1, always written
    assert int(dd[5]) == (2 << 8 | 3) | 0xC0000000 and int(dd[7]) == 3 | 0xC0000000            # implicit layers
    st = words[lay["style_base"]:lay["clip_aux_base"]].reshape(-1, 3)
    assert int(st[5, 1]) == 0 and int(st[6, 1]) == 1                                         # their ordinals
    pd = words[lay["path_data_base"]:lay["draw_tag_base"]].view(np.float32)
    assert list(pd[:4]) == [10, 10, 50, 10]                                                  # the first clip's path is the clip shape
    img, _ = T.render_packed(words, lay, 64, 64)
    assert img[25, 25, 3] == 128 and img[35, 45, 3] == 0 and img[9, 9, 3] == 0              # the red fill shows inside its layer's clip only
    assert img[12, 12, 3] == 0 and img[20, 20, 3] == 128      # the (empty) Copy layer wiped its clip's 15 x 15 px, not the rest of the tile
    assert tuple(img[5, 5]) == (0, 179, 0, 179)               # green at the inner implicit layer's alpha


def test_ingest_survives_malformed_encodings():
    """Random tag / data streams (short streams, NaNs, unbalanced clips and layers, unknown nesting): add_encoding answers
    with an error or a well-formed packed scene, never a crash (every stream access is bounds-checked, scene/decoder.go does
    the same for the CPU renderer)."""
    rng = np.random.default_rng(1)
    tags_pool = [0x01, 0x02, 0x10, 0x11, 0x12, 0x13, 0x14, 0x16, 0x17, 0x20, 0x21, 0x22, 0x30, 0x31, 0x40, 0x41, 0x50]
    c = _lib.Context(-1)
    ok = err = 0
    for _ in range(800):
        tags = rng.choice(tags_pool, int(rng.integers(0, 60))).astype(np.uint8)
        pd = rng.uniform(-50, 300, int(rng.integers(0, 80))).astype(np.float32)
        if rng.random() < 0.1 and len(pd):
            pd[rng.integers(0, len(pd), 3)] = np.nan
        dd = rng.integers(0, 2 ** 32, int(rng.integers(0, 40)), dtype=np.uint64).astype(np.uint32)
        tr = rng.uniform(-2, 2, 6 * int(rng.integers(0, 4))).astype(np.float32)
        br = rng.uniform(0, 1, 4 * int(rng.integers(0, 5)))
        c.begin(int(rng.integers(1, 300)), int(rng.integers(1, 300)))
        try:
            c.add_encoding(tags, pd, dd, tr, br)
        except _lib.GGCudaError as e:
            assert e.code in (_lib.ERR_INVALID, _lib.ERR_UNSUPPORTED)
            err += 1
            continue
        words, lay = c.pack_host()
        tail = words[int(lay["n_scene_words"]):]
        assert len(words) == lay["n_scene_words"] + 8 + 16 * tail[6] and lay["n_draws"] == lay["n_paths"]   # one path marker per draw; 8 tail words; round-rect SDF records
        ok += 1
    assert ok > 100 and err > 100


def test_random_scenes_pack_and_render():
    """Structured random scenes with every degenerate the ingest rules name (empty paths, repeated points, collapsed curves,
    zero-width and huge strokes, unbalanced pops, clips of empty paths): packed scenes the CPU twin renders without incident."""
    rng = np.random.default_rng(2)
    c = _lib.Context(-1)

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
    rendered = 0
    for _ in range(250):
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
        c.begin(w, h)
        try:
            c.add_encoding(*sc.Encoding().streams())
        except _lib.GGCudaError:
            continue
        words, lay = c.pack_host()
        img, _ = T.render_packed(words, lay, w, h)
        assert img.shape == (h, w, 4)
        rendered += 1
    assert rendered > 200
