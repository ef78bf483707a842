#!/usr/bin/env python
"""Build tuning variants of libgoofer_b200.so (same sources, extra -D flags) into goofer_b200/_lib/variants/.

    python tools/build_variants.py name1:-DFOO=1,-DBAR=2 name2:-DFOO=3 ...

Select one at run time with GOOFER_B200_LIB=goofer_b200/_lib/variants/<name>.so (goofer_b200/_build.py).
The variants travel to the GPU box like the main library; they are not committed.
"""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from goofer_b200 import _build

out_dir = os.path.join(_build.LIB_DIR, "variants")
os.makedirs(out_dir, exist_ok=True)

def one(spec):
    name, _, flags = spec.partition(":")
    cmd = ["nvcc"] + _build.NVCC_FLAGS + [f for f in flags.split(",") if f] + ["-o", os.path.join(out_dir, name + ".so")] + _build.sources()
    r = subprocess.run(cmd, capture_output=True, text=True)
    return name, r.returncode, r.stderr[-2000:]

with ThreadPoolExecutor(4) as ex:
    for name, rc, err in ex.map(one, sys.argv[1:]):
        print(name, "ok" if rc == 0 else "FAILED\n" + err)
