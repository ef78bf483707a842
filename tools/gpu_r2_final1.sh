#!/usr/bin/env bash
# Round 2, final single-GPU call: whole GPU test suite, every single-GPU bench line, launch list + full ncu capture of one c2 step.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_smoke.log
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "worst|passed|failed|skipped" gpurun_out/r2_pytest_gpu.log | tail -14
bash tools/bench_c5.sh 1
python bench.py --workload c1 --steps 10 --warmup 3 > gpurun_out/r2_bench_c1_1024notes.json 2> gpurun_out/r2_c1.err
python bench.py --workload c3 --notes 256 --steps 5 --cpu-sample 24 > gpurun_out/r2_bench_c3_256notes.json 2> gpurun_out/r2_c3.err
python bench.py --workload c3 --notes 1024 --steps 5 --cpu-sample 24 > gpurun_out/r2_bench_c3_1024notes.json 2> gpurun_out/r2_c3b.err
python bench.py --workload c4 --notes 96 --steps 5 --cpu-sample 12 > gpurun_out/r2_bench_c4_96notes.json 2> gpurun_out/r2_c4.err
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2_bench_c2_reference_arm.json 2> gpurun_out/r2_ref.err
for f in gpurun_out/r2_bench_c1_1024notes.json gpurun_out/r2_bench_c3_256notes.json gpurun_out/r2_bench_c3_1024notes.json gpurun_out/r2_bench_c4_96notes.json gpurun_out/r2_bench_c2_reference_arm.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    e = d.get("e2e") or {}
    v = d.get("verify") or {}
    print(f"{sys.argv[1]:52s} {d['value']:10.1f} {d['unit']}  {d['ms_per_step']:8.3f} ms/step  e2e {e.get('value', 0):10.1f} ({e.get('ms_per_step', 0):.2f} ms) verify {v.get('ok')} {v.get('worst_max_abs')} cpu {d.get('cpu_baseline', {}).get('value')}")
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
done
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --cpu-sample 0 --verify 0"
$CMD > gpurun_out/r2_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_c2_1024notes.csv $CMD > gpurun_out/r2_ncu_l.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/r2_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gf_ -s 54 -c 18 -o gpurun_out/r2_full_c2 $CMD > gpurun_out/r2_ncu_f.log 2>&1
echo "full rc=$?"
CMD2="python bench.py --steps 1 --warmup 3 --no-e2e --cpu-sample 0 --verify 0 --noise device"
$CMD2 > gpurun_out/r2_plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gf_phi_kernel -s 3 -c 1 -o gpurun_out/r2_full_phi $CMD2 > gpurun_out/r2_ncu_p.log 2>&1
echo "phi rc=$?"
