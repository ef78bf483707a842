#!/usr/bin/env python
"""Attribute an ncu report's per-instruction counters to CUDA source lines (no GUI needed).

    python tools/ncu_lines.py <report.ncu-rep> <kernel regex> <lib.so> [top N] [sort: samples|smem|inst|conflicts]

The regex matches the function's base name (no template arguments); NCU_LAUNCH_SKIP=k picks the (k+1)-th matching launch
(two instantiations of one template share the base name).

Joins `ncu --page source --csv` (per SASS address: instructions executed, stall samples, shared-memory wavefronts
and the excess over the ideal count = bank conflicts) with `nvdisasm -g` line info of the cubin inside the shared
library.  The .so must be the build that was profiled.
"""
import csv, os, re, subprocess, sys, tempfile, collections

rep, kre, so = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
sort = sys.argv[5] if len(sys.argv) > 5 else "samples"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}", "--launch-skip", os.environ.get("NCU_LAUNCH_SKIP", "0"),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
lines = out.split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
kname = lines[start - 1].split('","')[1].split("(")[0] if start else kre
rows = list(csv.reader(lines[start:]))
H = rows[0]
col = {n: H.index(n) for n in ("Address", "Instructions Executed", "# Samples")}
opt = {n: (H.index(n) if n in H else None) for n in ("L1 Wavefronts Shared", "L1 Wavefronts Shared Excessive", "L1 Wavefronts Shared Ideal")}


def num(r, i):
    if i is None or i >= len(r):
        return 0
    try:
        return int(float(r[i] or 0))
    except ValueError:
        return 0


recs = []
for r in rows[1:]:
    if len(r) <= col["# Samples"] or not r[0].startswith("0x"):
        if recs and r and (r[0].startswith("Kernel Name") or r[0] == "Address"):
            break                                           # next launch of the same kernel
        continue
    recs.append((int(r[0], 16), num(r, col["Instructions Executed"]), num(r, col["# Samples"]),
                 num(r, opt["L1 Wavefronts Shared"]), num(r, opt["L1 Wavefronts Shared Excessive"])))
base = recs[0][0]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin") and "plan" not in f][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.split("\n")
inside = False
cur = ("?", 0)
off2line = {}
for l in dis:
    if l.startswith("//---") and ".text." in l:
        inside = (kname in l) or (kre in l)              # templated kernels: the section carries the mangled name
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/", l)
    if m:
        off2line[int(m.group(1), 16)] = cur
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
tot = [0, 0, 0, 0]
for addr, n, smp, wf, ex in recs:
    key = off2line.get(addr - base, ("?", 0))
    for i, v in enumerate((n, smp, wf, ex)):
        agg[key][i] += v
        tot[i] += v
print(f"{kname}: {tot[0]} warp instructions, {tot[1]} samples, {tot[2]} shared wavefronts of which {tot[3]} excessive (bank conflicts)")
order = {"samples": 1, "smem": 2, "inst": 0, "conflicts": 3}[sort]
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][order])[:top]:
    print(f"{key[0]:16s}:{key[1]:<5d} inst {100 * v[0] / max(1, tot[0]):5.1f}%  samples {100 * v[1] / max(1, tot[1]):5.1f}%  "
          f"smem wavefronts {100 * v[2] / max(1, tot[2]):5.1f}%  excessive {100 * v[3] / max(1, tot[3]):5.1f}% ({v[3]})")
