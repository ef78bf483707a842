set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu55.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu55.log
python bench.py --steps 10 --warmup 3 --cpu-sample 0 > gpurun_out/bench55.log 2> gpurun_out/bench55.err; echo rc=$?
python - <<'PY'
import json
for f in ("bench55",):
    d=json.load(open(f"gpurun_out/{f}.log"))
    print(f, d["value"], d["ms_per_step"], d.get("e2e",{}).get("ms_per_step"), d["gpu_launches"], d["roofline"]["kernels_ms_per_step"])
PY
cat > /tmp/probe.py <<'PY'
import os, sys, time, argparse
sys.path.insert(0, os.getcwd())
import torch, bench
args = argparse.Namespace(workload="c2", notes=1024)
ab, _ = bench.build_batch(args, 0)
ab.pin()
for parts in ("", "0.25,0.5,0.75", "0.2,0.4,0.6,0.8", "0.167,0.333,0.5,0.667,0.833", "0.125,0.25,0.375,0.5,0.625,0.75,0.875", "0.3,0.5,0.7,0.85,0.95"):
    if parts: os.environ["GOOFER_HOST_PARTS"] = parts
    for _ in range(3): ab.render_host()
    t = time.perf_counter()
    for _ in range(6): ab.render_host()
    print("parts", parts or "default", "ms/step", 1e3 * (time.perf_counter() - t) / 6, flush=True)
os.environ["GOOFER_HOST_NO_PULL"] = "1"
os.environ["GOOFER_HOST_PARTS"] = "0.25,0.5,0.75"
for _ in range(3): ab.render_host()
t = time.perf_counter()
for _ in range(6): ab.render_host()
print("no pull, quarters", "ms/step", 1e3 * (time.perf_counter() - t) / 6, flush=True)
del os.environ["GOOFER_HOST_NO_PULL"]
os.environ["GOOFER_HOST_TRACE"] = "1"
ab.render_host(); ab.render_host()
PY
python /tmp/probe.py > gpurun_out/e2e55.log 2>&1
grep -v "^\[host" gpurun_out/e2e55.log; tail -5 gpurun_out/e2e55.log
