"""Host-side scene ingest timing (no GPU needed): python tools/ingest_bench.py [bands] ; GGCUDA_INGEST_THREADS=n selects the pool size."""
import hashlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gg_b200 import _lib, scenes  # noqa: E402

bands = int(sys.argv[1]) if len(sys.argv) > 1 else 1
enc, w, h = scenes.config3(bands=bands)
s = enc.streams()
c = _lib.Context(-1)
if bands > 1:
    c.begin(w, h)
    c.set_band(0, 135)
best = 1e9
reps = int(os.environ.get('GG_INGEST_REPS', 5))
for rep in range(reps):
    t0 = time.perf_counter()
    n = 20 if reps > 1 else 1
    for _ in range(n):
        c.begin(w, h)
        if bands > 1:
            c.set_band(0, 135)
        c.add_encoding(*s)
    best = min(best, (time.perf_counter() - t0) / n * 1e3)
words, lay = c.pack_host()
print(f"threads={os.environ.get('GGCUDA_INGEST_THREADS', '1 (default)')} cpus={os.cpu_count()} bands={bands} ingest {best:.3f} ms  md5 {hashlib.md5(words.tobytes()).hexdigest()[:12]}")
