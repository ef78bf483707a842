#!/usr/bin/env bash
# Round 2, GPU call A: parity at full sizes, the new bench line, e2e timeline, launch list + full ncu capture of one step.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2a_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_c2.json 2> gpurun_out/r2a_bench_c2.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2a_bench_c2.json"))
    print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 3))
    print({k: round(v["ms_per_step"], 3) for k, v in d["e2e"]["variants"].items()})
    print(d["roofline"]["kernels_ms_per_step"]); print(d.get("verify"))
except Exception as e:
    print("bench parse failed", e)
PY
GOOFER_HOST_TRACE=1 python tools/scratch/e2e_trace.py > gpurun_out/r2a_e2e_trace.log 2>&1; tail -25 gpurun_out/r2a_e2e_trace.log
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --cpu-sample 0 --verify 0"
$CMD > gpurun_out/r2a_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2a_launches.csv $CMD > gpurun_out/r2a_ncu_l.log 2>&1
echo "launch list rc=$?"
# one whole step under --set full: skip the 3 warm-up steps (16 launches each incl. meta copies), capture two steps
$CMD > gpurun_out/r2a_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gf_ -s 48 -c 32 -o gpurun_out/r2a_full $CMD > gpurun_out/r2a_ncu_f.log 2>&1
echo "full rc=$?"
ls -la gpurun_out | tail -12
