#!/usr/bin/env bash
# Round 2, GPU call D (1 GPU): deferred uploads + faster transposer; c2 line; c5 at 65,536 notes on one GPU.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -q -x > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2d_pytest.log
GOOFER_HOST_TRACE=1 python tools/scratch/e2e_trace.py > gpurun_out/r2d_e2e_trace.log 2>&1; grep -E "host" gpurun_out/r2d_e2e_trace.log | sed -n 12,24p; grep "ms per call" gpurun_out/r2d_e2e_trace.log
bash tools/bench_c5.sh 1
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_c2_1gpu.json").read().strip().splitlines()[-1])
print({k: round(v["ms_per_step"], 3) for k, v in d["e2e"]["variants"].items()})
print(d["roofline"]["kernels_ms_per_step"])
PY
