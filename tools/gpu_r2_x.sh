#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -q -x > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r2x_pytest.log
python bench.py --steps 20 --warmup 5 --cpu-sample 0 --e2e-variants prod > gpurun_out/r2x_bench_c2.json 2> gpurun_out/r2x_c2.err; echo "c2 rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2x_bench_c2.json")); e = d["e2e"]; v = d["verify"]
print(f"{d['value']:10.1f} {d['ms_per_step']:8.3f} ms/step  e2e {e['value']:10.1f} ({e['ms_per_step']:.2f} ms) verify {v['ok']} {v['worst_max_abs']}")
print(d["roofline"]["kernels_ms_per_step"])
PY
