#!/usr/bin/env bash
# after skipping the onset scan for passes longer than 4 s: parity + c4 / c2 lines
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -q -x > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r2u_pytest.log
python bench.py --workload c4 --notes 96 --steps 5 --cpu-sample 12 > gpurun_out/r2_bench_c4_96notes.json 2> gpurun_out/r2_c4.err; echo "c4 rc=$?"
python bench.py --steps 20 --warmup 5 --cpu-sample 0 > gpurun_out/r2u_bench_c2.json 2> gpurun_out/r2u_c2.err; echo "c2 rc=$?"
for f in gpurun_out/r2_bench_c4_96notes.json gpurun_out/r2u_bench_c2.json; do python - "$f" <<'PY'
import json, sys
d = json.load(open(sys.argv[1])); e = d["e2e"]; v = d["verify"]
print(f"{sys.argv[1]:44s} {d['value']:10.1f} {d['ms_per_step']:8.3f} ms/step  e2e {e['value']:10.1f} ({e['ms_per_step']:.2f} ms) verify {v['ok']} {v['worst_max_abs']}")
print(d["roofline"]["kernels_ms_per_step"])
PY
done
