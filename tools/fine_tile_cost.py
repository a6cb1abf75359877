"""Distribution of fine's per-tile work after restart points: executed PTCL words per tile and per tile pair.
    python tools/fine_tile_cost.py [workload]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402,F401

import bench  # noqa: E402
from gg_b200 import _lib  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "config3"
enc, w, h, bg, _ = bench.build_workload(name)
ctx = _lib.Context(0)
ctx.begin(w, h)
ctx.set_background(bg)
ctx.add_encoding(*enc.streams())
ctx.upload()
frame = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
for _ in range(8):
    ctx.render_device(frame.data_ptr(), w * 4, _lib.KEEP_SCENE)
poff = ctx.debug_read(_lib.BUF_PTCL_OFF, np.uint32).astype(np.int64)
ptcl = ctx.debug_read(_lib.BUF_PTCL, np.uint32)
rst = ctx.debug_read(_lib.BUF_RESTART, np.uint32).reshape(-1, 2)[:, 0].astype(np.int64)
wt, ht = (w + 15) // 16, (h + 15) // 16
n = wt * ht
pos = poff[:n] + np.maximum(rst[:n], 1)
ex = np.zeros(n, dtype=np.int64)
nclip = np.zeros(n, dtype=np.int64)
nseg = np.zeros(n, dtype=np.int64)
size = np.zeros(16, dtype=np.int64)
size[[0, 1, 3, 5, 6, 10, 11]] = [1, 4, 1, 2, 2, 1, 3]
active = np.ones(n, dtype=bool)
while active.any():
    idx = np.nonzero(active)[0]
    tags = ptcl[pos[idx]]
    f = tags == 1
    nseg[idx[f]] += ptcl[pos[idx[f]] + 1] >> 1
    nclip[idx[tags == 11]] += 1
    sz = size[np.minimum(tags, 15)]
    ex[idx] += sz
    pos[idx] += sz
    active[idx[(tags == 0) | (sz == 0)]] = False
print(f"EndClips executed per tile: mean {nclip.mean():.2f} max {nclip.max()}; segments per tile: mean {nseg.mean():.1f} max {nseg.max()}")
# count EndClip commands executed per tile (the expensive ones): walk the few heaviest tiles only
print(f"{name}: tiles {n}, executed words total {ex.sum()}, mean {ex.mean():.1f}, max {ex.max()}")
for q in (50, 90, 99, 99.9):
    print(f"  p{q}: {np.percentile(ex, q):.0f} words")
pairs = ex.reshape(ht, wt)
if wt % 2:
    pairs = np.pad(pairs, ((0, 0), (0, 1)))
pc = pairs.reshape(ht, -1, 2).sum(axis=2).ravel()
print(f"pairs {len(pc)}: mean {pc.mean():.1f} max {pc.max()}  top-10 {np.sort(pc)[-10:]}  share of work in the top 1 % of pairs {np.sort(pc)[-len(pc)//100:].sum() / pc.sum():.2f}")
heavy = np.argsort(pc)[-5:]
print("heaviest pairs at (row, col):", [(int(i // (pairs.shape[1] // 2)), int(i % (pairs.shape[1] // 2))) for i in heavy])
hist, edges = np.histogram(pc, bins=[0, 1, 8, 16, 32, 64, 128, 256, 512, 1024, 1 << 20])
print("pair cost histogram (words):", dict(zip([f"<{e}" for e in edges[1:]], hist.tolist())))
