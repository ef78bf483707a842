python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --cpu-sample 0 > gpurun_out/bench61.log 2> gpurun_out/bench61.err; echo rc=$?
python bench.py --steps 10 --warmup 3 --workload c1 --cpu-sample 0 --no-e2e > gpurun_out/bench61_c1.log 2> gpurun_out/bench61_c1.err; echo rc=$?
python - <<'PY'
import json
for f in ("bench61","bench61_c1"):
    try:
        d=json.load(open(f"gpurun_out/{f}.log"))
        print(f, round(d["value"]), round(d["ms_per_step"],3), round(d.get("e2e",{}).get("ms_per_step",0),3), d["gpu_launches"], d["roofline"]["kernels_ms_per_step"])
    except Exception as e: print(f, "FAILED", e)
PY
