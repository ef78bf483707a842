#!/usr/bin/env bash
# default bench line (with the CPU legs) + full ncu capture of one c2 step (18 kernels); keep gpurun_out below 64 MiB per call
mkdir -p gpurun_out
python bench.py > gpurun_out/r2_bench_c2_1024notes.json 2> gpurun_out/r2_c2_default.err; echo "default bench rc=$?"
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --cpu-sample 0 --verify 0"
$CMD > gpurun_out/r2_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gf_ -s 54 -c 18 -o gpurun_out/r2_full_c2 -f $CMD > gpurun_out/r2_ncu_f.log 2>&1
echo "full c2 rc=$?"; du -sh gpurun_out
