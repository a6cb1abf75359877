"""Extract the known-answer vectors of the reference's own blend tests into blend_kats.json.

Run in the build container (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_blend_kats.py
Sources: internal/blend/porter_duff_test.go, advanced_test.go, hsl_test.go -- every table row of the
form {name, sr,sg,sb,sa, dr,dg,db,da, wr,wg,wb,wa} inside a `func TestBlend<Mode>` whose body calls
blend<Mode>(...). Rows are keyed by scene.BlendMode (scene/encoding.go:17-48).
"""
import json
import os
import re

REF = "/root/reference/internal/blend"
MODE = {"Multiply": 1, "Screen": 2, "Overlay": 3, "Darken": 4, "Lighten": 5, "ColorDodge": 6, "ColorBurn": 7, "HardLight": 8,
        "SoftLight": 9, "Difference": 10, "Exclusion": 11, "Hue": 12, "Saturation": 13, "Color": 14, "Luminosity": 15,
        "Clear": 16, "Source": 17, "Destination": 18, "SourceOver": 19, "DestinationOver": 20, "SourceIn": 21,
        "DestinationIn": 22, "SourceOut": 23, "DestinationOut": 24, "SourceAtop": 25, "DestinationAtop": 26, "Xor": 27, "Plus": 28}
out = []
for fn in ("porter_duff_test.go", "advanced_test.go", "hsl_test.go"):
    src = open(os.path.join(REF, fn)).read()
    funcs = list(re.finditer(r"^func Test(\w+)\(t \*testing\.T\) \{", src, re.M))
    for i, m in enumerate(funcs):
        body = src[m.end(): funcs[i + 1].start() if i + 1 < len(funcs) else len(src)]
        name = m.group(1)
        if not name.startswith("Blend") or name[5:] not in MODE or f"blend{name[5:]}(" not in body:
            continue
        body = re.sub(r"//[^\n]*", "", body)
        for row in re.finditer(r"\{\s*\"([^\"]*)\",((?:\s*\d+\s*,){12})\s*\}", body):
            nums = [int(x) for x in re.findall(r"\d+", row.group(2))]
            out.append({"mode": MODE[name[5:]], "func": "blend" + name[5:], "case": row.group(1), "file": fn,
                        "s": nums[0:4], "d": nums[4:8], "want": nums[8:12]})
json.dump(out, open(os.path.join(os.path.dirname(__file__), "blend_kats.json"), "w"), indent=0)
print(len(out), "vectors;", sorted({o["func"] for o in out}))
