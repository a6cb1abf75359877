"""Render one workload N times on one context (no e2e leg, no oracle): the command to wrap in ncu.
    python tools/profile_run.py [config1|config2|config3|config3_nowipe|config4|config5] [renders]
The first render grows the buffers (several passes); every later render is one steady-state pass."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402  (device memory for the frame only)

import bench  # noqa: E402
from gg_b200 import _lib  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "config3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
enc, w, h, bg, _ = bench.build_workload(name)
ctx = _lib.Context(0)
ctx.set_timing(True)
ctx.begin(w, h)
ctx.set_background(bg)
ctx.add_encoding(*enc.streams())
ctx.upload()
frame = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
for i in range(n):
    ctx.render_device(frame.data_ptr(), w * 4, _lib.KEEP_SCENE)
    s = ctx.stats()
    print(i, "passes", s["passes"], "launches", s["kernel_launches"], {k: round(s[k], 3) for k in ("ms_front", "ms_binning", "ms_coarse", "ms_fine")})
print({k: s[k] for k in ("n_lines", "n_path_tiles", "n_segments", "n_hits", "n_ptcl_words", "device_bytes")})
