python tests/gpu_debug.py > gpurun_out/parity_table.txt 2>&1; tail -20 gpurun_out/parity_table.txt
python bench.py --steps 2 --warmup 3 --cpu-sample 0 --no-e2e > gpurun_out/b_pre.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_v4.csv python bench.py --steps 2 --warmup 3 --cpu-sample 0 --no-e2e > gpurun_out/ncu_l.log 2>&1; echo rc=$?
ncu --set full --import-source on --clock-control none --kernel-name 'regex:gf_(src_env|tracks|mask|fir32|f0|walk|onset|pulse|env|frame|peak|mix)_kernel' --launch-skip 36 --launch-count 12 -o gpurun_out/r1_full_v3 -f python bench.py --steps 1 --warmup 3 --cpu-sample 0 --no-e2e > gpurun_out/ncu63.log 2>&1; echo rc=$?
grep "PROF== Profiling" gpurun_out/ncu63.log | head -14
