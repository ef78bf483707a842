python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --cpu-sample 0 > gpurun_out/bench_2gpu_v2.log 2> gpurun_out/bench_2gpu_v2.err; echo rc=$?
tail -c 1500 gpurun_out/bench_2gpu_v2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_2gpu_ref.log 2> gpurun_out/bench_2gpu_ref.err; echo rc=$?
tail -c 600 gpurun_out/bench_2gpu_ref.log
