"""Extracts the outlines of the printable ASCII glyphs of the reference's test font
(gogpu/gg text/testdata/goregular.ttf, the Go Regular face) into tests/golden/fixtures/goregular_ascii.json:
per glyph the advance width and the contours as TrueType quadratic paths in font units (y up). Run in the build
container, where /root/reference exists; the fixture is committed because the GPU box has no reference tree.

    python tests/golden/make_glyph_outlines.py [/root/reference/text/testdata/goregular.ttf]

Minimal `glyf` reader (simple glyphs only; composite glyphs are skipped -- none among the ASCII range of this font)."""
import json
import os
import struct
import sys


def tables(data):
    n = struct.unpack(">H", data[4:6])[0]
    out = {}
    for i in range(n):
        tag, _, off, ln = struct.unpack(">4sIII", data[12 + 16 * i:28 + 16 * i])
        out[tag.decode()] = data[off:off + ln]
    return out


def cmap_format4(cmap):
    n = struct.unpack(">H", cmap[2:4])[0]
    for i in range(n):
        pid, eid, off = struct.unpack(">HHI", cmap[4 + 8 * i:12 + 8 * i])
        if (pid, eid) in ((3, 1), (0, 3), (0, 4)) and struct.unpack(">H", cmap[off:off + 2])[0] == 4:
            t = cmap[off:]
            segx2 = struct.unpack(">H", t[6:8])[0]
            seg = segx2 // 2
            end = struct.unpack(f">{seg}H", t[14:14 + segx2])
            start = struct.unpack(f">{seg}H", t[16 + segx2:16 + 2 * segx2])
            delta = struct.unpack(f">{seg}h", t[16 + 2 * segx2:16 + 3 * segx2])
            ro_off = 16 + 3 * segx2
            ro = struct.unpack(f">{seg}H", t[ro_off:ro_off + segx2])
            m = {}
            for s in range(seg):
                for c in range(start[s], min(end[s], 0xFFFE) + 1):
                    if ro[s] == 0:
                        g = (c + delta[s]) & 0xFFFF
                    else:
                        p = ro_off + 2 * s + ro[s] + 2 * (c - start[s])
                        g = struct.unpack(">H", t[p:p + 2])[0]
                        if g:
                            g = (g + delta[s]) & 0xFFFF
                    if g:
                        m[c] = g
            return m
    raise RuntimeError("no format-4 cmap")


def glyph_contours(glyf, off, end):
    if end <= off:
        return []
    g = glyf[off:end]
    nc = struct.unpack(">h", g[0:2])[0]
    if nc < 0:
        return None   # composite
    ends = struct.unpack(f">{nc}H", g[10:10 + 2 * nc])
    npts = ends[-1] + 1 if nc else 0
    ilen = struct.unpack(">H", g[10 + 2 * nc:12 + 2 * nc])[0]
    p = 12 + 2 * nc + ilen
    flags = []
    while len(flags) < npts:
        f = g[p]; p += 1
        flags.append(f)
        if f & 8:
            r = g[p]; p += 1
            flags += [f] * r
    def coords(short_bit, same_bit):
        nonlocal p
        out, v = [], 0
        for f in flags:
            if f & short_bit:
                d = g[p]; p += 1
                v += d if f & same_bit else -d
            elif not f & same_bit:
                v += struct.unpack(">h", g[p:p + 2])[0]; p += 2
            out.append(v)
        return out
    xs = coords(2, 16)
    ys = coords(4, 32)
    contours, s = [], 0
    for e in ends:
        pts = [(xs[i], ys[i], flags[i] & 1) for i in range(s, e + 1)]
        s = e + 1
        contours.append(pts)
    return contours


def to_quads(pts):
    """TrueType contour (on/off-curve points, implied on-curve midpoints) -> [("M", x, y), ("L", x, y) | ("Q", cx, cy, x, y)...]."""
    n = len(pts)
    if n == 0:
        return []
    # start at an on-curve point (or the midpoint of two off-curve points)
    st = next((i for i in range(n) if pts[i][2]), None)
    if st is None:
        x0, y0 = (pts[0][0] + pts[-1][0]) / 2, (pts[0][1] + pts[-1][1]) / 2
        seq = pts
    else:
        x0, y0 = pts[st][0], pts[st][1]
        seq = pts[st + 1:] + pts[:st + 1]
    out = [("M", x0, y0)]
    ctrl = None
    for x, y, on in seq:
        if on:
            out.append(("Q", ctrl[0], ctrl[1], x, y) if ctrl else ("L", x, y))
            ctrl = None
        else:
            if ctrl:
                mx, my = (ctrl[0] + x) / 2, (ctrl[1] + y) / 2
                out.append(("Q", ctrl[0], ctrl[1], mx, my))
            ctrl = (x, y)
    if ctrl:
        out.append(("Q", ctrl[0], ctrl[1], x0, y0))
    return out


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/text/testdata/goregular.ttf"
    t = tables(open(src, "rb").read())
    upem = struct.unpack(">H", t["head"][18:20])[0]
    long_loca = struct.unpack(">h", t["head"][50:52])[0]
    ng = struct.unpack(">H", t["maxp"][4:6])[0]
    loca = struct.unpack(f">{ng + 1}I", t["loca"][:4 * (ng + 1)]) if long_loca else \
        tuple(2 * v for v in struct.unpack(f">{ng + 1}H", t["loca"][:2 * (ng + 1)]))
    nhm = struct.unpack(">H", t["hhea"][34:36])[0]
    adv = [struct.unpack(">H", t["hmtx"][4 * min(i, nhm - 1):4 * min(i, nhm - 1) + 2])[0] for i in range(ng)]
    cm = cmap_format4(t["cmap"])
    glyphs = {}
    for c in range(33, 127):
        g = cm.get(c)
        if g is None:
            continue
        cs = glyph_contours(t["glyf"], loca[g], loca[g + 1])
        if cs is None:
            continue
        glyphs[chr(c)] = {"advance": adv[g], "contours": [to_quads(p) for p in cs]}
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fixtures", "goregular_ascii.json")
    json.dump({"font": "Go Regular (gogpu/gg text/testdata/goregular.ttf)", "units_per_em": upem, "glyphs": glyphs}, open(out, "w"),
              separators=(",", ":"))
    print(out, len(glyphs), "glyphs", os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
