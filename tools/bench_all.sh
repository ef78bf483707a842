#!/usr/bin/env bash
# Every single-GPU bench line of the round (c1-c4, one c5 shard, the reference arm), one JSON file each under
# gpurun_out/.  Run from the repo root on a B200: `bash tools/bench_all.sh`; copy the files to profiles/ afterwards.
mkdir -p gpurun_out
python bench.py > gpurun_out/final_c2_1024notes.json 2> gpurun_out/final_c2.err
python bench.py --workload c1 > gpurun_out/final_c1_1024notes.json 2> gpurun_out/final_c1.err
python bench.py --workload c3 --notes 256 --steps 5 --cpu-sample 24 > gpurun_out/final_c3_256notes.json 2> gpurun_out/final_c3.err
python bench.py --workload c4 --notes 96 --steps 5 --cpu-sample 12 > gpurun_out/final_c4_96notes.json 2> gpurun_out/final_c4.err
python bench.py --workload c5 --notes 8192 --steps 3 --cpu-sample 0 > gpurun_out/final_c5_8192notes_1gpu.json 2> gpurun_out/final_c5.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_c2_reference_arm.json 2> gpurun_out/final_ref.err
for f in gpurun_out/final_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    e = d.get("e2e") or {}
    print(f"{sys.argv[1]:50s} {d['value']:10.1f} {d['unit']}  {d['ms_per_step']:8.3f} ms/step  e2e {e.get('value', 0):10.1f}")
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
done
