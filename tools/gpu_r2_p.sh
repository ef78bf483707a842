#!/usr/bin/env bash
# ncu: full capture of the overlap-save kernels and the growl scan on one c3 step (256 notes) + launch list of that step
mkdir -p gpurun_out
CMD="python bench.py --workload c3 --notes 256 --steps 1 --warmup 1 --cpu-sample 0 --verify 0 --no-e2e"
$CMD > gpurun_out/r2p_plain.log 2>&1; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2p_launches_c3_256notes.csv $CMD > gpurun_out/r2p_ncu_l.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"fftconv|sg_scan|gf_fir" -c 10 -o gpurun_out/r2p_full_conv $CMD > gpurun_out/r2p_ncu_f.log 2>&1; echo "full rc=$?"
ls -la gpurun_out/*.ncu-rep
