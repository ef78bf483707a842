#!/usr/bin/env bash
# Round 2, GPU call H (1 GPU): sources uploaded before planning, envelope kernel code-size work.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -q -x -s > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "worst|passed|failed" gpurun_out/r2h_pytest.log | tail -12
bash tools/bench_variants.sh
GOOFER_HOST_TRACE=1 python tools/scratch/e2e_trace.py > gpurun_out/r2h_e2e_trace.log 2>&1; grep -E "host" gpurun_out/r2h_e2e_trace.log | sed -n 14,28p; grep "ms per call" gpurun_out/r2h_e2e_trace.log; grep "part" gpurun_out/r2h_e2e_trace.log | sed -n 5,8p
python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench_c2.json 2> gpurun_out/r2h_bench_c2.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2h_bench_c2.json"))
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 3))
print({k: (round(v["ms_per_step"], 3), v["rank0_call_ms"]) for k, v in d["e2e"]["variants"].items()})
print(d["roofline"]["kernels_ms_per_step"]); print(d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
PY
