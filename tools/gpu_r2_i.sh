#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -q -x > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2i_pytest.log
GOOFER_HOST_TRACE=1 python tools/scratch/e2e_trace.py > gpurun_out/r2i_e2e_trace.log 2>&1; grep -E "host" gpurun_out/r2i_e2e_trace.log | sed -n 14,28p; grep "ms per call" gpurun_out/r2i_e2e_trace.log
bash tools/bench_c5.sh 1
