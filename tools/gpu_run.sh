python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/bench60.log 2> gpurun_out/bench60.err; echo rc=$?
python bench.py --steps 10 --warmup 3 --workload c1 > gpurun_out/bench60_c1.log 2> gpurun_out/bench60_c1.err; echo rc=$?
python bench.py --steps 5 --warmup 3 --workload c3 --notes 256 --cpu-sample 24 > gpurun_out/bench60_c3.log 2> gpurun_out/bench60_c3.err; echo rc=$?
python bench.py --steps 5 --warmup 3 --workload c4 --notes 96 --cpu-sample 6 > gpurun_out/bench60_c4.log 2> gpurun_out/bench60_c4.err; echo rc=$?
python - <<'PY'
import json
for f in ("bench60","bench60_c1","bench60_c3","bench60_c4"):
    try:
        d=json.load(open(f"gpurun_out/{f}.log"))
        print(f, round(d["value"]), round(d["ms_per_step"],3), round(d.get("e2e",{}).get("ms_per_step",0),3), d["gpu_launches"], d["roofline"]["kernels_ms_per_step"], d.get("cpu_baseline",{}).get("value"))
    except Exception as e: print(f, "FAILED", e)
PY
