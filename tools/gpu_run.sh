python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --cpu-sample 0 > gpurun_out/bench62.log 2> gpurun_out/bench62.err; echo rc=$?
python - <<'PY'
import json
for f in ("bench62",):
    try:
        d=json.load(open(f"gpurun_out/{f}.log"))
        print(f, round(d["value"]), round(d["ms_per_step"],3), d.get("e2e"), d["gpu_launches"], d["roofline"]["kernels_ms_per_step"])
    except Exception as e: print(f, "FAILED", e)
PY
