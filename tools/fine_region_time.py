"""Fine-stage time of sub-rectangles of a resident frame (dirty-rect renders): where does the fine kernel's time go?
    python tools/fine_region_time.py [workload]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from gg_b200 import _lib  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "config3"
enc, w, h, bg, _ = bench.build_workload(name)
ctx = _lib.Context(0)
ctx.set_timing(True)
frame = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
key = enc.CacheKey()
rects = {"whole": None, "without columns 0-31": (32, 0, w, h), "columns 0-31 only": (0, 0, 32, h), "left half": (0, 0, w // 2, h), "top half": (0, 0, w, h // 2)}
for label, r in rects.items():
    best = 1e9
    for i in range(6):
        resident = ctx.begin_keyed(w, h, key)
        ctx.set_background(bg)
        if not resident:
            ctx.add_encoding(*enc.streams())
        if r:
            ctx.set_dirty_rect(*r)
        ctx.render_device(frame.data_ptr(), w * 4, _lib.KEEP_SCENE)
        s = ctx.stats()
        if i >= 2:
            best = min(best, s["ms_fine"])
    print(f"{name} {label:24s} fine {best:.3f} ms")
