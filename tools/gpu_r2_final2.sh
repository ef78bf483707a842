#!/usr/bin/env bash
# Round 2, last single-GPU check of the committed build: smoke(), the whole GPU suite, the default bench line and the c3 lines
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_smoke.log
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "worst|passed|failed|skipped" gpurun_out/r2_pytest_gpu.log | tail -9
python bench.py > gpurun_out/r2_bench_c2_1024notes.json 2> gpurun_out/r2_c2_default.err; echo "default bench rc=$?"
python bench.py --workload c3 --notes 256 --steps 5 --cpu-sample 24 > gpurun_out/r2_bench_c3_256notes.json 2> gpurun_out/r2_c3.err
python bench.py --workload c3 --notes 1024 --steps 5 --cpu-sample 24 > gpurun_out/r2_bench_c3_1024notes.json 2> gpurun_out/r2_c3b.err
for f in gpurun_out/r2_bench_c2_1024notes.json gpurun_out/r2_bench_c3_256notes.json gpurun_out/r2_bench_c3_1024notes.json; do python - "$f" <<'PY'
import json, sys
d = json.load(open(sys.argv[1])); e = d["e2e"]; v = d["verify"]
print(f"{sys.argv[1]:46s} {d['value']:10.1f} {d['ms_per_step']:8.3f} ms/step  e2e {e['value']:10.1f} ({e['ms_per_step']:.2f} ms) verify {v['ok']} {v['worst_max_abs']} cpu {d['cpu_baseline']['value']:.1f} launches {d['gpu_launches']}")
PY
done
