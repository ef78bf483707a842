#!/usr/bin/env bash
# full ncu capture of the kernels that are new in c3 (overlap-save convolution, event / onset scans) + launch list of one c3 step
mkdir -p gpurun_out
CMD3="python bench.py --workload c3 --notes 256 --steps 1 --warmup 1 --cpu-sample 0 --verify 0 --no-e2e"
$CMD3 > gpurun_out/r2_plain4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"fftconv|sg_scan|walk_scan|gf_walk_kernel" -s 5 -c 5 -o gpurun_out/r2_full_c3_new -f $CMD3 > gpurun_out/r2_ncu_c3.log 2>&1
echo "full c3 rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c3_256notes.csv $CMD3 > gpurun_out/r2_ncu_l3.log 2>&1; echo "c3 launch list rc=$?"; du -sh gpurun_out
