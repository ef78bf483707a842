#!/usr/bin/env bash
# Round 2, GPU call F (1 GPU): per-block tick tables in the f0 kernel; register-cap variants; parity.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -q -x -s > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "worst|passed|failed" gpurun_out/r2f_pytest.log | tail -12
bash tools/bench_variants.sh
