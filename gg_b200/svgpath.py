"""SVG path-data parser (the `d` attribute) for the icon scenes: M L H V C S Q T Z, absolute and relative.
Mirrors what gg's svg package hands to scene.Scene (svg/path_parser.go produces the same verb sequence for these
commands); elliptical arcs are not needed by the in-tree icons and raise."""
import re

from .scene import CLOSE, CUBIC, LINE, MOVE, QUAD

_TOKEN = re.compile(r"([MmLlHhVvCcSsQqTtZzAa])|([-+]?(?:\d*\.\d+|\d+\.?)(?:[eE][-+]?\d+)?)")


def parse_path(d):
    """-> (verbs, coords) in user units."""
    toks = [(m.group(1), m.group(2)) for m in _TOKEN.finditer(d)]
    verbs, coords = [], []
    i, cmd = 0, None
    x = y = sx = sy = 0.0
    last_c = None   # reflected control point candidates for S / T
    last_q = None

    def num():
        nonlocal i
        v = float(toks[i][1])
        i += 1
        return v

    while i < len(toks):
        if toks[i][0]:
            cmd = toks[i][0]
            i += 1
            if cmd in "Zz":
                verbs.append(CLOSE)
                x, y = sx, sy
                last_c = last_q = None
                continue
        if cmd is None:
            raise ValueError("path data does not start with a command")
        rel = cmd.islower()
        c = cmd.upper()
        ox, oy = (x, y) if rel else (0.0, 0.0)
        if c == "M":
            x, y = ox + num(), oy + num()
            sx, sy = x, y
            verbs.append(MOVE); coords += [x, y]
            cmd = "l" if rel else "L"   # subsequent pairs are implicit LineTo
            last_c = last_q = None
        elif c == "L":
            x, y = ox + num(), oy + num()
            verbs.append(LINE); coords += [x, y]
            last_c = last_q = None
        elif c == "H":
            x = ox + num()
            verbs.append(LINE); coords += [x, y]
            last_c = last_q = None
        elif c == "V":
            y = oy + num()
            verbs.append(LINE); coords += [x, y]
            last_c = last_q = None
        elif c == "C":
            x1, y1, x2, y2 = ox + num(), oy + num(), ox + num(), oy + num()
            x, y = ox + num(), oy + num()
            verbs.append(CUBIC); coords += [x1, y1, x2, y2, x, y]
            last_c, last_q = (x2, y2), None
        elif c == "S":
            x1, y1 = (2 * x - last_c[0], 2 * y - last_c[1]) if last_c else (x, y)
            x2, y2 = ox + num(), oy + num()
            x, y = ox + num(), oy + num()
            verbs.append(CUBIC); coords += [x1, y1, x2, y2, x, y]
            last_c, last_q = (x2, y2), None
        elif c == "Q":
            x1, y1 = ox + num(), oy + num()
            x, y = ox + num(), oy + num()
            verbs.append(QUAD); coords += [x1, y1, x, y]
            last_q, last_c = (x1, y1), None
        elif c == "T":
            x1, y1 = (2 * x - last_q[0], 2 * y - last_q[1]) if last_q else (x, y)
            x, y = ox + num(), oy + num()
            verbs.append(QUAD); coords += [x1, y1, x, y]
            last_q, last_c = (x1, y1), None
        else:
            raise NotImplementedError("elliptical arcs")
    return verbs, coords
