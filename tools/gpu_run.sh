set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu58.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu58.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench58.log 2> gpurun_out/bench58.err; echo rc=$?
python - <<'PY'
import json
for f in ("bench58",):
    d=json.load(open(f"gpurun_out/{f}.log"))
    print(f, d["value"], d["ms_per_step"], d.get("e2e",{}).get("ms_per_step"), d["gpu_launches"], d["roofline"]["kernels_ms_per_step"], d.get("cpu_baseline"))
PY
bash tools/bench_variants.sh
