"""Stage-by-stage comparison of the CUDA pipeline (through the C ABI) with the CPU twin oracle.

A scene is a list of element dicts:
  {"type": "draw"|"begin_clip"|"end_clip", "verbs": u8[], "coords": f64[], "color": (r,g,b,a) straight u8,
   "even_odd": bool, "blend": u32, "alpha": float}
"""
import numpy as np

from gg_b200 import _lib as G
from oracle import twin as T


def oracle_scene(elems, w, h):
    """Flatten every path with the oracle's restatement of path_convert.go and run the CPU twin."""
    oe = []
    for e in elems:
        if e["type"] == "end_clip":
            oe.append(dict(type=T.ELEM_END_CLIP))
            continue
        lines = T.flatten_path(e["verbs"], e["coords"], auto_close=True)
        if e["type"] == "draw":
            oe.append(dict(type=T.ELEM_DRAW, lines=lines, color=e["color"], even_odd=e.get("even_odd", False)))
        else:
            oe.append(dict(type=T.ELEM_BEGIN_CLIP, lines=lines, blend=e.get("blend", 0x8003), alpha=e.get("alpha", 1.0)))
    el, ln = T.make_elements(oe)
    return T.Coarse(el, ln, w, h)


def gpu_scene(ctx, elems, w, h, bg=(0, 0, 0, 0), band=None):
    """Feed the same scene through the per-draw C ABI and render it; returns the premultiplied RGBA8 frame."""
    ctx.begin(w, h)
    ctx.set_background(bg)
    ht = (h + 15) // 16
    ctx.set_band(*(band or (0, ht)))
    for e in elems:
        if e["type"] == "draw":
            ctx.fill_path(e["verbs"], e["coords"], e["color"], 1 if e.get("even_odd") else 0)
        elif e["type"] == "begin_clip":
            ctx.push_clip(e["verbs"], e["coords"])
        else:
            ctx.pop()
    out = np.zeros((h, w, 4), dtype=np.uint8)
    ctx.flush(out, flags=G.KEEP_SCENE)
    return out


def gpu_encoding(ctx, enc, w, h, bg=(0, 0, 0, 0), band=None):
    """Render a scene.Encoding (gg_b200.scene.Encoding) through ggcuda_add_encoding + ggcuda_flush."""
    ctx.begin(w, h)
    ctx.set_background(bg)
    ht = (h + 15) // 16
    ctx.set_band(*(band or (0, ht)))
    ctx.add_encoding(*enc.streams())
    out = np.zeros((h, w, 4), dtype=np.uint8)
    ctx.flush(out, flags=G.KEEP_SCENE)
    return out


def oracle_from_ctx(ctx, w, h):
    """CPU twin run on the packed scene the context uploaded (oracle/packed.c)."""
    words = ctx.debug_read(G.BUF_SCENE, np.uint32)
    lay = ctx.debug_read(G.BUF_LAYOUT, G.LAYOUT)[0]
    return T.Coarse.from_packed(words, lay, w, h)


def sorted_rows(a):
    """Sort a structured array's rows by their raw bytes (order-independent multiset comparison)."""
    if len(a) == 0:
        return a
    v = np.ascontiguousarray(a).view(np.uint8).reshape(len(a), -1)
    idx = np.lexsort(v.T[::-1])
    return a[idx]


def compare_scans(ctx):
    """a3 / a4 / a5: the device's pathtag scan, draw scan + draw leaf (info, clip inputs) and clip-leaf fix-up against the
    oracle's restatement of pathtag.go:76-121, draw_leaf.go:54-151 and clip_leaf.go:27-56 run on the SAME packed scene."""
    lay = ctx.debug_read(G.BUF_LAYOUT, G.LAYOUT)[0]
    words = ctx.debug_read(G.BUF_SCENE, np.uint32)
    tm, dm, info, ci = T.scan_stages(words, lay)
    gtm = ctx.debug_read(G.BUF_TAG_MONOIDS, G.PATH_MONOID)
    assert len(gtm) == len(tm) and gtm.tobytes() == tm.astype(G.PATH_MONOID).tobytes(), "tag monoids differ"
    gdm = ctx.debug_read(G.BUF_DRAW_MONOIDS, G.DRAW_MONOID)
    assert len(gdm) == len(dm) and gdm.tobytes() == dm.astype(G.DRAW_MONOID).tobytes(), "draw monoids (after clip leaf) differ"
    ginfo = ctx.debug_read(G.BUF_INFO, np.uint32)
    assert (ginfo[:len(info)] == info).all(), "draw info differs"
    gci = ctx.debug_read(G.BUF_CLIP_INPS, G.CLIP_INP)
    assert len(gci) == len(ci) and (gci["ix"].astype(np.int64) == ci[:, 0].astype(np.uint32)).all() and (gci["path_ix"] == ci[:, 1]).all(), "clip inputs differ"
    return {"tag_words": len(tm), "draws": len(dm), "clips": len(ci)}


def compare_stages(ctx, oc, elems, w, h, check_ptcl=True):
    """Assert bit-exact equality of every integer stage; returns a dict of counts for reporting."""
    rep = {}
    # ---- a3 / a4 / a5: monoid scans, draw leaf, clip leaf
    rep.update(compare_scans(ctx))
    lay = ctx.debug_read(G.BUF_LAYOUT, G.LAYOUT)[0]
    # ---- a6: lines, per path, as multisets (the reference orders LineTo lines before flattened cubics)
    gl = ctx.debug_read(G.BUF_LINES, G.LINE)
    n_paths = len(oc.paths)
    order = np.argsort(gl["path_ix"], kind="stable")
    gls = gl[order]
    starts = np.searchsorted(gls["path_ix"], np.arange(n_paths + 1))
    if elems is None:
        # packed-scene oracle: both sides flatten in tag order, so for fills the arrays must be identical
        lay = ctx.debug_read(G.BUF_LAYOUT, G.LAYOUT)[0]
        words = ctx.debug_read(G.BUF_SCENE, np.uint32)
        ol = T.flatten_packed(words, lay).astype(G.LINE)
        styles = words[lay["style_base"]:lay["clip_aux_base"]:3]
        if not (styles & 1).any():
            assert len(ol) == len(gl), f"{len(gl)} lines, oracle {len(ol)}"
            assert ol.tobytes() == gl.tobytes(), "flattened lines differ"
        else:
            # stroked paths: the device keeps two fixed slots per outline piece (zero-length lines stay in place, the
            # oracle drops them) and appends the surplus of rare rectangle pieces out of order -> per-path multisets
            def canon(a):
                a = a[(a["p0"] != a["p1"]).any(axis=1)]
                return sorted_rows(a[np.argsort(a["path_ix"], kind="stable")])
            ga, oa = canon(gl), canon(ol)
            assert len(ga) == len(oa), f"{len(ga)} non-degenerate lines, oracle {len(oa)}"
            assert ga.tobytes() == oa.tobytes(), "flattened / stroked lines differ"
    for p, e in enumerate(elems or []):
        g = gls[starts[p]:starts[p + 1]]
        if e["type"] == "end_clip":
            assert len(g) == 0
            continue
        o = T.flatten_path(e["verbs"], e["coords"], auto_close=True).astype(G.LINE)
        o["path_ix"] = p
        assert len(g) == len(o), f"path {p}: {len(g)} lines, oracle {len(o)}"
        assert sorted_rows(g).tobytes() == sorted_rows(o).tobytes(), f"path {p}: flattened lines differ"
    rep["lines"] = len(gl)
    # ---- a7: paths
    gp = ctx.debug_read(G.BUF_PATHS, G.PATH)
    assert len(gp) == n_paths, (len(gp), n_paths)
    for p in range(n_paths):
        ob = oc.paths[p]["bbox"]
        gb = gp[p]["bbox"]
        ow, oh = int(ob[2]) - int(ob[0]), int(ob[3]) - int(ob[1])
        if ow > 0 and oh > 0:
            assert (ob == gb).all(), f"path {p} bbox {gb} != oracle {ob}"
            assert gp[p]["tiles"] == oc.paths[p]["tiles"], f"path {p} tile offset"
    # ---- a8/a9/a10: tiles
    gt = ctx.debug_read(G.BUF_TILES, G.TILE)
    gs = ctx.debug_read(G.BUF_SEG_START, np.uint32)
    assert len(gt) == len(oc.tiles), (len(gt), len(oc.tiles))
    assert (gt["backdrop"] == oc.tiles["backdrop"]).all(), "backdrop mismatch"
    has = oc.tiles["seg_count_or_ix"] != 0
    assert ((gt["seg_count"] != 0) == has).all(), "tile occupancy mismatch"
    # oracle stores ~local_start per path; global start = path_seg_base + local
    base_per_tile = np.zeros(len(oc.tiles), dtype=np.uint32)
    for p in range(n_paths):
        ob = oc.paths[p]["bbox"]
        cnt = max(0, int(ob[2]) - int(ob[0])) * max(0, int(ob[3]) - int(ob[1]))
        t0 = int(oc.paths[p]["tiles"])
        base_per_tile[t0:t0 + cnt] = oc.path_seg_base[p]
    o_start = (~oc.tiles["seg_count_or_ix"]).astype(np.uint32) + base_per_tile
    assert (gs[has] == o_start[has]).all(), "segment start mismatch"
    rep["tiles"] = len(gt)
    # ---- a11: segments, per tile as multisets
    gseg = ctx.debug_read(G.BUF_SEGMENTS, G.SEGMENT)
    assert len(gseg) == len(oc.segments), (len(gseg), len(oc.segments))
    idx = np.nonzero(has)[0]
    for i in idx:
        s, n = int(gs[i]), int(gt["seg_count"][i])
        a = sorted_rows(gseg[s:s + n])
        b = sorted_rows(oc.segments[s:s + n].astype(G.SEGMENT))
        assert a.tobytes() == b.tobytes(), f"segments of path-tile {i} differ"
    rep["segments"] = len(gseg)
    # ---- a12: PTCL word for word
    if check_ptcl:
        poff = ctx.debug_read(G.BUF_PTCL_OFF, np.uint32)
        pw = ctx.debug_read(G.BUF_PTCL, np.uint32)
        ng = oc.wt * oc.ht
        assert len(poff) == ng
        words = 0
        for t in range(ng):
            o = oc.ptcl(t)
            g = pw[poff[t]:poff[t] + len(o)]
            assert (g == o).all(), f"PTCL of tile {t} differs:\n gpu {g}\n ref {o}"
            words += len(o)
        rep["ptcl_words"] = words
    return rep


def pixel_diff(a, b):
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    return int(d.max()) if d.size else 0, float(d.mean()) if d.size else 0.0, float((d.max(axis=-1) > 0).mean()) if d.size else 0.0


# ---------------------------------------------------------------- scene builders
MOVE, LINE, QUAD, CUBIC, CLOSE = 0, 1, 2, 3, 4


def circle_path(cx, cy, r):
    """scene/path.go:192-213 kappa circle: 4 cubics, closed."""
    k = r * 0.5522847498
    verbs = [MOVE, CUBIC, CUBIC, CUBIC, CUBIC, CLOSE]
    coords = [cx + r, cy,
              cx + r, cy + k, cx + k, cy + r, cx, cy + r,
              cx - k, cy + r, cx - r, cy + k, cx - r, cy,
              cx - r, cy - k, cx - k, cy - r, cx, cy - r,
              cx + k, cy - r, cx + r, cy - k, cx + r, cy]
    return np.array(verbs, dtype=np.uint8), np.array(coords, dtype=np.float64)


def polygon_path(pts):
    verbs = [MOVE] + [LINE] * (len(pts) - 1) + [CLOSE]
    return np.array(verbs, dtype=np.uint8), np.array(pts, dtype=np.float64).ravel()


def random_scene(seed, w, h, n, clips=False, quads=True, max_size=96.0):
    """Random closed blobs / circles / polygons with solid colours (config #1 distribution, small)."""
    rng = np.random.default_rng(seed)
    elems = []
    depth = 0
    for i in range(n):
        kind = rng.integers(0, 4 if quads else 3)
        cx, cy = rng.uniform(0, w - 1), rng.uniform(0, h - 1)   # every shape touches the canvas (the reference's bbox wraps otherwise)
        col = tuple(int(x) for x in rng.integers(0, 256, 3)) + (int(rng.integers(100, 256)),)
        if kind == 0:
            v, c = circle_path(np.float32(cx), np.float32(cy), np.float32(rng.uniform(2, max_size / 2)))
        elif kind == 1:
            m = int(rng.integers(3, 7))
            pts = np.stack([cx + rng.uniform(-max_size / 2, max_size / 2, m), cy + rng.uniform(-max_size / 2, max_size / 2, m)], axis=1)
            pts[0] = (cx, cy)
            v, c = polygon_path(pts.astype(np.float32))
        elif kind == 2:
            m = int(rng.integers(2, 5))
            p0 = np.array([cx, cy])
            verbs, coords = [MOVE], list(p0)
            for _ in range(m):
                pts = p0 + rng.uniform(-max_size / 2, max_size / 2, (3, 2))
                verbs.append(CUBIC)
                coords += list(pts.ravel())
            verbs.append(CLOSE)
            v, c = np.array(verbs, dtype=np.uint8), np.array(coords, dtype=np.float32).astype(np.float64)
        else:
            m = int(rng.integers(2, 5))
            p0 = np.array([cx, cy])
            verbs, coords = [MOVE], list(p0)
            for _ in range(m):
                pts = p0 + rng.uniform(-max_size / 2, max_size / 2, (2, 2))
                verbs.append(QUAD)
                coords += list(pts.ravel())
            verbs.append(CLOSE)
            v, c = np.array(verbs, dtype=np.uint8), np.array(coords, dtype=np.float32).astype(np.float64)
        if clips and rng.random() < 0.15 and depth < 6:
            elems.append(dict(type="begin_clip", verbs=v, coords=c))
            depth += 1
            continue
        if clips and depth > 0 and rng.random() < 0.15:
            elems.append(dict(type="end_clip"))
            depth -= 1
        elems.append(dict(type="draw", verbs=v, coords=c, color=col, even_odd=(not clips) and bool(rng.random() < 0.25)))
    while depth > 0:
        elems.append(dict(type="end_clip"))
        depth -= 1
    return elems


def compare_stages_fast(ctx, oc, w, h):
    """compare_stages for full-size scenes (10^4..10^6 paths): the same bit-exact checks on the packed-scene oracle,
    vectorised -- lines as a multiset, path bboxes, tile backdrops / occupancy / segment starts, segments as per-tile
    multisets, PTCL word for word."""
    lay = ctx.debug_read(G.BUF_LAYOUT, G.LAYOUT)[0]
    words = ctx.debug_read(G.BUF_SCENE, np.uint32)
    rep = {}
    rep.update(compare_scans(ctx))   # a3 / a4 / a5
    # a6
    gl = ctx.debug_read(G.BUF_LINES, G.LINE)
    ol = T.flatten_packed(words, lay).astype(G.LINE)

    def canon(a):
        return sorted_rows(a[(a["p0"] != a["p1"]).any(axis=1)])
    ga, oa = canon(gl), canon(ol)
    assert len(ga) == len(oa), f"{len(ga)} non-degenerate lines, oracle {len(oa)}"
    assert ga.tobytes() == oa.tobytes(), "flattened / stroked lines differ"
    rep["lines"] = len(ga)
    # a7
    gp = ctx.debug_read(G.BUF_PATHS, G.PATH)
    assert len(gp) == len(oc.paths)
    ob, gb = oc.paths["bbox"].astype(np.int64), gp["bbox"].astype(np.int64)
    live = ((ob[:, 2] - ob[:, 0]) > 0) & ((ob[:, 3] - ob[:, 1]) > 0)
    assert (ob[live] == gb[live]).all(), "path bboxes differ"
    assert (gp["tiles"][live] == oc.paths["tiles"][live]).all(), "path tile offsets differ"
    # a8-a10
    gt = ctx.debug_read(G.BUF_TILES, G.TILE)
    gs = ctx.debug_read(G.BUF_SEG_START, np.uint32)
    assert len(gt) == len(oc.tiles), (len(gt), len(oc.tiles))
    assert (gt["backdrop"] == oc.tiles["backdrop"]).all(), "backdrop mismatch"
    has = oc.tiles["seg_count_or_ix"] != 0
    assert ((gt["seg_count"] != 0) == has).all(), "tile occupancy mismatch"
    cnt = np.where(live, (ob[:, 2] - ob[:, 0]) * (ob[:, 3] - ob[:, 1]), 0)
    base_per_tile = np.repeat(oc.path_seg_base.astype(np.uint32), cnt)
    t0 = np.repeat(oc.paths["tiles"].astype(np.int64), cnt)
    assert (np.diff(t0) >= 0).all() and len(base_per_tile) == len(oc.tiles)
    o_start = (~oc.tiles["seg_count_or_ix"]).astype(np.uint32) + base_per_tile
    assert (gs[has] == o_start[has]).all(), "segment start mismatch"
    rep["tiles"] = len(gt)
    # a11: segments sorted inside their tile ranges
    gseg = ctx.debug_read(G.BUF_SEGMENTS, G.SEGMENT)
    assert len(gseg) == len(oc.segments), (len(gseg), len(oc.segments))
    tile_of = np.zeros(len(gseg), np.int64)
    idx = np.nonzero(has)[0]
    tile_of[gs[idx]] = 1
    tile_of = np.cumsum(tile_of)     # ranges are contiguous and ordered by start

    def by_tile(a):
        v = np.ascontiguousarray(a).view(np.uint8).reshape(len(a), -1)
        order = np.lexsort(tuple(v.T[::-1]) + (tile_of,))
        return a[order]
    assert by_tile(gseg).tobytes() == by_tile(oc.segments.astype(G.SEGMENT)).tobytes(), "segments differ"
    rep["segments"] = len(gseg)
    # a12
    poff = ctx.debug_read(G.BUF_PTCL_OFF, np.uint32).astype(np.int64)
    pw = ctx.debug_read(G.BUF_PTCL, np.uint32)
    ng = oc.wt * oc.ht
    olen = np.diff(oc.ptcl_offsets.astype(np.int64))
    take = np.repeat(poff[:ng], olen) + (np.arange(int(olen.sum())) - np.repeat(oc.ptcl_offsets[:ng].astype(np.int64), olen))
    assert (pw[take] == oc.ptcl_words).all(), "PTCL differs"
    rep["ptcl_words"] = int(olen.sum())
    return rep
