#!/usr/bin/env bash
# e2e: tapered download parts (smaller last part = shorter tail after the last kernel)
mkdir -p gpurun_out
for parts in default "0.28,0.56,0.84" "0.3,0.6,0.86" "0.32,0.62,0.88" "0.27,0.54,0.78,0.93" "0.3,0.58,0.8,0.94" "0.35,0.65,0.88" "0.22,0.44,0.66,0.84,0.95"; do
  if [ "$parts" = default ]; then unset GOOFER_HOST_PARTS; else export GOOFER_HOST_PARTS=$parts; fi
  python bench.py --steps 20 --warmup 5 --cpu-sample 0 --verify 0 --e2e-variants prod > gpurun_out/r2t.json 2> gpurun_out/r2t.err
  python - "$parts" <<'PY'
import json, sys
d = json.load(open("gpurun_out/r2t.json"))
print(f"{sys.argv[1]:32s} step {d['ms_per_step']:.3f} e2e {d['e2e']['ms_per_step']:.3f} calls {d['e2e']['rank0_call_ms']}")
PY
done
