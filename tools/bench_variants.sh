#!/usr/bin/env bash
# Bench every variant library in goofer_b200/_lib/variants/ (tools/build_variants.py) on this box: device-resident step,
# per-kernel CUDA-event times and the headline end-to-end variant; no CPU leg.  Run from the repo root on a B200.
mkdir -p gpurun_out
for f in goofer_b200/_lib/variants/*.so; do
  n=$(basename $f .so)
  GOOFER_B200_LIB=$PWD/$f python bench.py --steps 10 --warmup 3 --cpu-sample 0 --verify 0 --e2e-variants prod "$@" > gpurun_out/var_$n.log 2> gpurun_out/var_$n.err
  python - "$n" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/var_{n}.log"))
    k = d["roofline"]["kernels_ms_per_step"]
    e = (d.get("e2e") or {}).get("ms_per_step")
    print(f"{n:12s} {d['ms_per_step']:.3f} ms  e2e {e if e is None else round(e, 3)} ms  " + " ".join(f"{a}={b:.3f}" for a, b in k.items()))
except Exception as e:
    print(n, "FAILED", e)
PY
done
