#!/usr/bin/env python
"""Attribute an ncu report's per-instruction counters to CUDA source lines (no GUI needed).

    python tools/ncu_lines.py <report.ncu-rep> <kernel regex> <lib.so> [top N]

Joins `ncu --page source --csv` (per SASS address: instructions executed, stall samples) with
`nvdisasm -g` line info of the cubin inside the shared library.  The .so must be the build that was profiled.
"""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, kre, so = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout
lines = out.split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
kname = lines[start - 1].split('","')[1].split("(")[0] if start else kre
rows = list(csv.reader(lines[start:]))
H = rows[0]
ia, ii, isamp = H.index("Address"), H.index("Instructions Executed"), H.index("# Samples")
recs = []
for r in rows[1:]:
    if len(r) <= isamp or not r[ia].startswith("0x"):
        if recs and r and r[0].startswith('"Kernel Name"'):
            break
        continue
    recs.append((int(r[ia], 16), int(r[ii] or 0), int(r[isamp] or 0), r[1].strip()))
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
        inside = kname in l
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
agg = collections.defaultdict(lambda: [0, 0])
tot_i = tot_s = 0
for addr, n, smp, txt in recs:
    key = off2line.get(addr - base, ("?", 0))
    agg[key][0] += n
    agg[key][1] += smp
    tot_i += n
    tot_s += smp
print(f"{kname}: {tot_i} warp instructions, {tot_s} samples")
for key, (n, smp) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{key[0]:16s}:{key[1]:<5d} inst {100 * n / max(1, tot_i):5.1f}%  samples {100 * smp / max(1, tot_s):5.1f}%")
