#!/usr/bin/env bash
# tail parts alternating between two streams: A/B against one stream (GOOFER_PART_STREAMS=1), then the GPU tests
mkdir -p gpurun_out
for ps in 2 1 2 1; do
  GOOFER_PART_STREAMS=$ps python bench.py --steps 20 --warmup 5 --cpu-sample 0 --verify 4 --e2e-variants all > gpurun_out/r2y_$ps.json 2> gpurun_out/r2y_$ps.err
  python - "$ps" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/r2y_{sys.argv[1]}.json")); e = d["e2e"]
print(f"streams {sys.argv[1]}: step {d['ms_per_step']:.3f} e2e {e['ms_per_step']:.3f} {e['rank0_call_ms']} variants", {k: round(v['ms_per_step'], 2) for k, v in e['variants'].items()}, d['verify']['ok'], d['verify'].get('e2e_pcm16_worst_lsb_diff'))
PY
done
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r2y_pytest.log
