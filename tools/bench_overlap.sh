#!/usr/bin/env bash
# Side-stream excitation chain (GOOFER_OVERLAP) against the single-stream order on every single-GPU workload.
# Run from the repo root on a B200; one line per (workload, setting).
mkdir -p gpurun_out
for w in c1 c2 c3 c4; do
  case $w in c3) n=256;; c4) n=96;; *) n=1024;; esac
  for o in 0 1; do
    GOOFER_OVERLAP=$o python bench.py --workload $w --notes $n --steps 10 --warmup 3 --cpu-sample 0 > gpurun_out/ov_${w}_$o.json 2> gpurun_out/ov_${w}_$o.err
    python - "$w" "$o" <<'PY'
import json, sys
w, o = sys.argv[1:3]
try:
    d = json.load(open(f"gpurun_out/ov_{w}_{o}.json"))
    print(f"{w} overlap={o} {d['ms_per_step']:.3f} ms/step  e2e {d['e2e']['ms_per_step']:.3f} ms")
except Exception as e:
    print(w, o, "FAILED", e)
PY
  done
done
