#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -q -x -s > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "worst|passed|failed" gpurun_out/r2n_pytest.log | tail -12
python bench.py --steps 20 --warmup 5 --cpu-sample 0 > gpurun_out/r2n_bench_c2.json 2> gpurun_out/r2n_bench_c2.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2n_bench_c2.json"))
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 3))
print(d["roofline"]["kernels_ms_per_step"]); print(d["verify"])
PY
