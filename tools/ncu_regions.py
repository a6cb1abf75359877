"""Where a kernel's instructions and stall samples go, by region of its source file.
    python tools/ncu_regions.py REPORT.ncu-rep [kernel-name-substring] [source-file]
Joins ncu's SASS page (per-instruction executed counts and stall samples, `--import-source on` captures) with the line table
of the SAME build of gg_b200/libggcuda.so (nvdisasm -g), and sums per region: a region starts at every top-level
`__device__` / `__global__` / `struct` definition of the source file and, inside fine_kernel, at every command handler."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
kname = sys.argv[2] if len(sys.argv) > 2 else "fine_kernel"
src = sys.argv[3] if len(sys.argv) > 3 else ("fine.cu" if "fine" in kname else "pipeline.cu")

# regions from the source text
anchors = []
for i, l in enumerate(open(os.path.join(ROOT, "gg_b200", "csrc", src)), 1):
    m = re.match(r"(?:template <[^>]*>\s*)?(?:__device__|__global__|static|struct)\b.*?\b([A-Za-z_][A-Za-z0-9_]*)\s*(?:\(|\{|$)", l)
    if m and not l.startswith(" "):
        name = m.group(1)
        if name == "__launch_bounds__":   # __global__ void __launch_bounds__(...) kernel(...)
            k = re.search(r"\)\s+([A-Za-z_][A-Za-z0-9_]*)\(", l)
            name = (k.group(1) if k else "kernel") + " (head)"
        anchors.append((i, name))
    m = re.search(r"(?:if|else if) \(tag == GG_CMD_([A-Z_]+)\)", l)
    if m:
        anchors.append((i, "cmd " + m.group(1).lower()))
    m = re.match(r"\s+// ---- ([A-Za-z][^:]{0,40})", l)
    if m:
        anchors.append((i, "phase: " + m.group(1).strip()))
    if "ss.drain();   // a slice fetched ahead" in l:
        anchors.append((i, "tile store"))
anchors.sort()


def region(line):
    name = "(head)"
    for a, n in anchors:
        if a <= line:
            name = n
        else:
            break
    return name


# line table of the shipped library
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "gg_b200", "libggcuda.so")], cwd=tmp, check=True, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.startswith(src.split(".")[0] + ".")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
off2line, cur, infn = {}, None, False
for l in dis.splitlines():
    if l.startswith(".text."):
        infn = kname in l
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+\S", l)
    if m and infn:
        off2line[int(m.group(1), 16)] = cur

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", "regex:" + kname], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[hdr_i], [r for r in rows[hdr_i + 1:] if len(r) == len(rows[hdr_i])]
A, N, IE, NI = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("stall_no_inst")
base = int(data[0][A], 16)
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
for r in data:
    fl = off2line.get(int(r[A], 16) - base)
    name = "?" if not fl else (region(fl[1]) if fl[0] == src else fl[0])
    e = agg[name]
    e[0] += int(r[IE]); e[1] += int(r[N]); e[2] += int(r[NI]); e[3] += 1 if int(r[IE]) > 0 else 0
tot = [sum(v[i] for v in agg.values()) for i in range(3)]
print(f"{kname}: {tot[0]} warp instructions executed, {tot[1]} stall samples ({tot[2]} of them no_instruction), {len(data)} SASS instructions ({len(data) * 16} bytes)")
print(f"{'region':28s} {'instr %':>8s} {'samples %':>10s} {'no_inst %':>10s} {'SASS executed':>14s}")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    if v[0] == 0 and v[1] == 0:
        continue
    print(f"{k:28s} {v[0] / tot[0] * 100:8.1f} {v[1] / tot[1] * 100:10.1f} {v[2] / max(1, tot[2]) * 100:10.1f} {v[3]:14d}")
