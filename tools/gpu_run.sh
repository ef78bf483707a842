python bench.py --workload c5 --notes 8192 --steps 3 --warmup 3 --cpu-sample 32 > gpurun_out/bench64_c5.log 2> gpurun_out/bench64_c5.err; echo rc=$?
tail -c 2500 gpurun_out/bench64_c5.log; tail -3 gpurun_out/bench64_c5.err
