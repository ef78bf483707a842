#!/usr/bin/env bash
# Round 2, GPU call G (1 GPU): tie-free prefix sums in the walk kernel, envelope kernel with several tiles per CTA + cp.async row prefetch.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -q -x -s > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "worst|passed|failed" gpurun_out/r2g_pytest.log | tail -12
bash tools/bench_variants.sh
for w in c1 c3; do
  python bench.py --workload $w --notes $([ $w = c3 ] && echo 256 || echo 1024) --steps 8 --warmup 3 --cpu-sample 0 --verify 4 > gpurun_out/r2g_bench_$w.json 2> gpurun_out/r2g_bench_$w.err
  python - "$w" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/r2g_bench_{sys.argv[1]}.json"))
print(sys.argv[1], round(d["ms_per_step"], 3), "ms", round(d["value"]), "notes/s; e2e", round(d["e2e"]["ms_per_step"], 3), d["roofline"]["kernels_ms_per_step"], d["verify"]["ok"], d["verify"]["worst_max_abs"])
PY
done
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --cpu-sample 0 --verify 0"
$CMD > gpurun_out/r2g_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gf_ -s 51 -c 17 -o gpurun_out/r2g_full $CMD > gpurun_out/r2g_ncu_f.log 2>&1
echo "ncu rc=$?"
