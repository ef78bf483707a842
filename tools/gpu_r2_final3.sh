#!/usr/bin/env bash
# Round 2, final numbers of the committed build on one GPU: default line, c1 / c4 / c5 lines, launch list + full ncu capture of one c2 step
mkdir -p gpurun_out
python bench.py > gpurun_out/r2_bench_c2_1024notes.json 2> gpurun_out/r2_c2_default.err; echo "default bench rc=$?"
python bench.py --workload c1 --steps 10 --warmup 3 > gpurun_out/r2_bench_c1_1024notes.json 2> gpurun_out/r2_c1.err
python bench.py --workload c4 --notes 96 --steps 5 --cpu-sample 12 > gpurun_out/r2_bench_c4_96notes.json 2> gpurun_out/r2_c4.err
python bench.py --workload c5 --notes 65536 --noise device --steps 3 --warmup 3 --cpu-sample 0 --verify 4 > gpurun_out/r2_bench_c5_65536notes_1gpu.json 2> gpurun_out/r2_c5.err
for f in gpurun_out/r2_bench_c2_1024notes.json gpurun_out/r2_bench_c1_1024notes.json gpurun_out/r2_bench_c4_96notes.json gpurun_out/r2_bench_c5_65536notes_1gpu.json; do python - "$f" <<'PY'
import json, sys
d = json.load(open(sys.argv[1])); e = d["e2e"]; v = d["verify"]
print(f"{sys.argv[1]:50s} {d['value']:10.1f} {d['ms_per_step']:8.3f} ms/step  e2e {e['value']:10.1f} ({e['ms_per_step']:.2f} ms) verify {v['ok']} {v['worst_max_abs']} launches {d['gpu_launches']}")
PY
done
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --cpu-sample 0 --verify 0"
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_c2_1024notes.csv $CMD > gpurun_out/r2_ncu_l.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gf_ -s 54 -c 18 -o gpurun_out/r2_full_c2 -f $CMD > gpurun_out/r2_ncu_f.log 2>&1; echo "full rc=$?"; du -sh gpurun_out
