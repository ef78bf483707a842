#!/usr/bin/env bash
# overlap-save Gaussian smoothing + growl event scan: parity, then c3 at 256 / 1,024 notes and c2 for regression
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -q -x -s > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "worst|passed|failed|Error|error" gpurun_out/r2o_pytest.log | tail -14
for cfg in "c3 256" "c3 1024" "c2 1024"; do
  set -- $cfg
  python bench.py --workload $1 --notes $2 --steps 10 --warmup 3 --cpu-sample 0 --e2e-variants prod > gpurun_out/r2o_bench_$1_$2.json 2> gpurun_out/r2o_bench_$1_$2.err; echo "bench $cfg rc=$?"
  python - gpurun_out/r2o_bench_$1_$2.json <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 3))
print(d["roofline"]["kernels_ms_per_step"]); print(d["verify"])
PY
done
