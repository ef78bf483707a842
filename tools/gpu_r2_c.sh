#!/usr/bin/env bash
# Round 2, GPU call C: always-frame-major phases (transposer for host phases), early phase generator, TMA variant of the frame kernel.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -q -x > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2c_pytest.log
GOOFER_B200_LIB=$PWD/goofer_b200/_lib/variants/tma.so timeout 900 python -m pytest tests/test_gpu_full_size.py -m gpu -q -x -k "whole_config or device_drawn" > gpurun_out/r2c_pytest_tma.log 2>&1; echo "pytest tma rc=$?"; tail -3 gpurun_out/r2c_pytest_tma.log
bash tools/bench_variants.sh
echo "--- device noise value ---"
bash tools/bench_variants.sh --noise device --no-e2e 2>/dev/null | sed 's/^/dn /'
GOOFER_HOST_TRACE=1 python tools/scratch/e2e_trace.py > gpurun_out/r2c_e2e_trace.log 2>&1; grep -E "host" gpurun_out/r2c_e2e_trace.log | tail -12; grep "part" gpurun_out/r2c_e2e_trace.log | tail -4
for v in c2 tma; do
CMD="env GOOFER_B200_LIB=$PWD/goofer_b200/_lib/variants/$v.so python bench.py --steps 1 --warmup 3 --no-e2e --cpu-sample 0 --verify 0"
$CMD > gpurun_out/r2c_plain_$v.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gf_frame_kernel -s 3 -c 1 -o gpurun_out/r2c_frame_$v $CMD > gpurun_out/r2c_ncu_$v.log 2>&1
echo "ncu $v rc=$?"
done
