#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -q -x > gpurun_out/r2f0b_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2f0b_pytest.log
for cfg in "c2 1024"; do
  set -- $cfg
  python bench.py --workload $1 --notes $2 --steps 10 --warmup 3 --cpu-sample 0 --e2e-variants prod > gpurun_out/r2f0b_bench_$1.json 2> gpurun_out/r2f0b_bench_$1.err; echo "bench $cfg rc=$?"
  python - gpurun_out/r2f0b_bench_$1.json <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 3), d["verify"]["ok"], d["verify"]["worst_max_abs"])
k = d["roofline"]["kernels_ms_per_step"]; print({a: k[a] for a in ("mask", "fir", "f0", "walk", "pulse") if a in k})
PY
done
