#!/usr/bin/env python
"""Static SASS size of one kernel by source region (no GPU): where a kernel's code bytes come from.

    python tools/sass_lines.py <lib.so> <kernel substring> [bucket lines, default 10]
"""
import collections, os, re, subprocess, sys, tempfile
so, kname = sys.argv[1:3]
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 10
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin") and "plan" not in f][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.split("\n")
inside, cur, cnt = False, ("?", 0), collections.Counter()
for l in dis:
    if l.startswith("//---") and ".text." in l:
        inside = kname in l
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l):
        cnt[cur] += 1
tot = sum(cnt.values())
print(f"{kname}: {tot} instructions = {tot * 16 / 1024:.1f} KB")
reg = collections.Counter()
for (f, ln), c in cnt.items():
    reg[(f, ln // bucket * bucket)] += c
for k in sorted(reg):
    print(f"  {k[0]:16s}:{k[1]:<5d} {reg[k]:6d}")
