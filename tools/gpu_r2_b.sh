#!/usr/bin/env bash
# Round 2, GPU call B: group-private frame kernel + frame-major phases: full GPU test suite, bench, variants, ncu of the frame kernel.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2b_smoke.log
timeout 1700 python -m pytest tests -m gpu -q -s > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2b_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench_c2.json 2> gpurun_out/r2b_bench_c2.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2b_bench_c2.json"))
    print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 3))
    print({k: round(v["ms_per_step"], 3) for k, v in d["e2e"]["variants"].items()})
    print(d["roofline"]["kernels_ms_per_step"]); print(d.get("verify"))
except Exception as e:
    print("bench parse failed", e)
PY
bash tools/bench_variants.sh
GOOFER_HOST_TRACE=1 python tools/scratch/e2e_trace.py > gpurun_out/r2b_e2e_trace.log 2>&1; grep -E "host:|ms per call" gpurun_out/r2b_e2e_trace.log | tail -6; grep "part" gpurun_out/r2b_e2e_trace.log | tail -4
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --cpu-sample 0 --verify 0"
$CMD > gpurun_out/r2b_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gf_frame_kernel\|gf_env_kernel -s 6 -c 2 -o gpurun_out/r2b_frame_env $CMD > gpurun_out/r2b_ncu_f.log 2>&1
echo "ncu rc=$?"
